"""CPU tests of the serving shell's session window (recommend_resource.rs:39-54 over sessions/mod.rs:37-71)."""
import hashlib

import numpy as np
import pytest

from util import random_index_data


@pytest.fixture()
def host_index(sb):
    items, off, ts = random_index_data(np.random.default_rng(1), 50, 10)
    return sb.VMISIndex.from_sessions(items, off, ts, 10, 8, 1.0, device=sb.DEVICE_NONE)


def reference_window(store, clock, sid, item, consent, max_items, idle=20 * 60):
    """recommend_resource.rs:39-54 + sessions/mod.rs:37-71, restated on a dict"""
    if not consent:
        return [item]
    got = store.get(sid)
    items = list(got[0]) if got and clock - got[1] <= idle else []
    if not items:
        items.append(item)
    elif items[-1] != item:
        items.append(item)
        if len(items) > max_items:
            del items[0:1]
    store[sid] = (items, clock)
    return items


def test_md5_matches_hashlib(sb):
    for msg in [b"", b"a", b"144", b"session-" * 9, bytes(range(200))]:
        assert sb.md5(msg) == hashlib.md5(msg).digest()


def test_session_window_follows_the_endpoint(sb, host_index):
    srv = sb.Server(host_index, 5, 10, 5, max_items_in_session=3)
    srv.set_clock(1_000_000)
    assert srv.session_window("s1", 7) == [7]
    assert srv.session_window("s1", 7) == [7]                       # repeat of the last item is not appended (:43)
    assert srv.session_window("s1", 8) == [7, 8]
    assert srv.session_window("s1", 7) == [7, 8, 7]                 # only the LAST item is compared
    assert srv.session_window("s1", 9) == [8, 7, 9]                 # drain(0..1) at max_items_in_session (:45-48)
    assert srv.session_window("s1", 1, user_consent=False) == [1]   # no consent: store untouched (:52-54)
    assert srv.stored_items("s1") == [8, 7, 9]
    assert srv.session_window("s2", 5) == [5]                       # another visitor
    # 20 minutes idle is still the same session, 20 min + 1 s starts over (sessions/mod.rs:45-52)
    srv.set_clock(1_000_000 + 20 * 60)
    assert srv.stored_items("s1") == [8, 7, 9]
    srv.set_clock(1_000_000 + 20 * 60 + 1)
    assert srv.stored_items("s1") == []
    assert srv.session_window("s1", 4) == [4]
    # TTL: entries older than 30 min are dropped from the store (serving.rs:55-56)
    assert srv.stats()["sessions"] == 2
    srv.set_clock(1_000_000 + 31 * 60)
    assert srv.stats()["sessions"] == 1
    srv.close()


def test_session_window_random_walk_vs_restatement(sb, host_index):
    rng = np.random.default_rng(3)
    srv = sb.Server(host_index, 5, 10, 5, max_items_in_session=4)
    store, clock = {}, 5_000_000
    for _ in range(5000):
        clock += int(rng.choice([0, 1, 30, 600, 1300], p=[0.3, 0.3, 0.2, 0.15, 0.05]))
        srv.set_clock(clock)
        sid = "visitor-%d" % rng.integers(0, 40)
        item = int(rng.integers(1, 6))
        consent = bool(rng.random() < 0.9)
        assert srv.session_window(sid, item, consent) == reference_window(store, clock, sid, item, consent, 4)
    srv.close()


def test_recommend_fails_loudly_without_gpu(sb, host_index):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    srv = sb.Server(host_index, 5, 10, 5, max_items_in_session=2)
    with pytest.raises(sb.VmisError):
        srv.recommend("s", 7)
    srv.close()
