import os
import sys
import zipfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def toy_dir(tmp_path_factory):
    """assets/example of the reference (train.txt / test.txt), committed as tests/golden/toy_example.zip."""
    d = tmp_path_factory.mktemp("toy")
    with zipfile.ZipFile(os.path.join(ROOT, "tests", "golden", "toy_example.zip")) as z:
        z.extractall(d)
    return str(d)


@pytest.fixture(scope="session")
def oracle():
    from oracle import vmis_oracle
    vmis_oracle.lib()
    return vmis_oracle


@pytest.fixture(scope="session")
def sb():
    import serenade_b200
    serenade_b200.load_library()
    return serenade_b200
