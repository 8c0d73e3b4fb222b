"""On-device index build (prepare_hashmap as radix sorts + kernels) and on-device synthetic generation:
bit-identical to the host-built index and to the CPU oracle."""
import os

import numpy as np
import pytest

from util import csr, random_index_data

pytestmark = pytest.mark.gpu


def _dev(arr, torch):
    view = {np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32}[arr.dtype]
    return torch.from_numpy(np.ascontiguousarray(arr).view(view)).cuda()


@pytest.mark.parametrize("seed", range(4))
def test_device_build_matches_host_build_and_oracle(sb, oracle, seed):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(seed)
    items, off, ts = random_index_data(rng, int(rng.integers(200, 3000)), int(rng.integers(10, 300)),
                                       max_len=int(rng.integers(3, 12)), unique_ts=bool(seed % 2),
                                       id_scale=int(rng.integers(1, 1 << 30)))
    m, max_len = int(rng.integers(1, 60)), int(rng.integers(3, 12))
    d_items, d_off, d_ts = _dev(items, torch), _dev(off, torch), _dev(ts, torch)
    dix = sb.VMISIndex.from_device_sessions(d_items.data_ptr(), d_off.data_ptr(), d_ts.data_ptr(), len(ts), m, max_len, 1.3)
    os.environ["VMIS_BUILD"] = "host"                    # force the host CSR builder for the comparison
    try:
        hix = sb.VMISIndex.from_sessions(items, off, ts, m, max_len, 1.3, device=0)
    finally:
        del os.environ["VMIS_BUILD"]
    oix = oracle.OracleIndex.from_sessions(items, off, ts, m, max_len, 1.3)
    ds, hs = dix.stats(), hix.stats()
    for key in ("n_sessions_kept", "n_items", "n_pairs_kept", "n_postings", "max_len", "m_build"):
        assert ds[key] == hs[key], key
    known = np.unique(items)
    for it in known[::7]:
        try:
            want = hix.idf(int(it))
        except KeyError:
            with pytest.raises(KeyError):
                dix.idf(int(it))
            continue
        assert dix.idf(int(it)) == want
    queries = [[int(x) for x in rng.choice(known, size=int(rng.integers(1, 9)))] for _ in range(400)] + [[10 ** 15], []]
    q_items, q_off = csr(queries)
    for k, mq, n in [(20, m, 21), (500, 500, 40), (3, max(1, m // 2), 5)]:
        a = sb.predict_batch(dix, queries, k, mq, n)
        b = sb.predict_batch(hix, queries, k, mq, n)
        o = oix.predict_batch(q_items, q_off, k, mq, n, mode=1)
        for x, y, z in zip(a, b, o[:3]):
            assert np.array_equal(x, y) and np.array_equal(x, z)
        sa = dix.find_neighbors_batch(queries, k, mq)
        sh = hix.find_neighbors_batch(queries, k, mq)
        assert all(np.array_equal(x, y) for x, y in zip(sa, sh))
    with pytest.raises(IndexError):
        dix.items_for_session(0)            # no host mirror of the sessions when the data came from the device
    aix = sb.VMISIndex.from_sessions(items, off, ts, m, max_len, 1.3, device=0)   # default: host data, device build
    assert np.array_equal(aix.items_for_session(0), items[off[0]:off[1]])
    a = sb.predict_batch(aix, queries, 20, m, 21)
    b = sb.predict_batch(hix, queries, 20, m, 21)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_device_synth_matches_host_synth(sb, oracle):
    n_items, n_sessions = 5000, 40000
    six = sb.VMISIndex.synth(42, n_items, n_sessions, 1502, 34, 2.0)
    items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
    os.environ["VMIS_BUILD"] = "host"
    try:
        hix = sb.VMISIndex.from_sessions(items, off, ts, 1502, 34, 2.0, device=0)
    finally:
        del os.environ["VMIS_BUILD"]
    ss, hs = six.stats(), hix.stats()
    for key in ("n_sessions_kept", "n_items", "n_pairs_kept", "n_postings"):
        assert ss[key] == hs[key], key
    q = sb.synth_queries(43, n_items, 2000, 4)
    a = sb.predict_batch(six, q, 288, 1502, 21)
    b = sb.predict_batch(hix, q, 288, 1502, 21)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 1502, 34, 2.0)
    o = oix.predict_batch(q[0], q[1], 288, 1502, 21, mode=1)
    assert all(np.array_equal(x, y) for x, y in zip(a, o[:3]))


def test_device_built_shards(sb):
    n_items, n_sessions, n_shards = 4000, 30000, 4
    full = sb.VMISIndex.synth(42, n_items, n_sessions, 300, 34, 2.0)
    shards = [sb.VMISIndex.synth(42, n_items, n_sessions, 300, 34, 2.0, 0, s, n_shards) for s in range(n_shards)]
    for a in range(n_shards):
        for b in range(n_shards):
            if a != b:
                shards[a].attach_shard_ptr(b, shards[b].shard_ptr())
    q = sb.synth_queries(43, n_items, 1500, 4)
    want = sb.predict_batch(full, q, 100, 300, 21)
    for sh in shards:
        got = sb.predict_batch(sh, q, 100, 300, 21)
        assert all(np.array_equal(x, y) for x, y in zip(got, want))


def test_save_load_roundtrip(sb, tmp_path):
    """serialised index blob: loaded handle answers bit-identically; attributes travel; bad files are rejected"""
    items, off, ts = sb.synth_sessions(42, 3000, 20000)
    ix = sb.VMISIndex.from_sessions(items, off, ts, 400, 34, 2.0, device=0)
    known = np.unique(items)
    ix.set_attributes(known[:50], np.full(50, 6, dtype=np.uint8))       # adult + for sale
    q = sb.synth_queries(43, 3000, 3000, 4)
    want = [sb.predict_batch(ix, q, 100, 400, 21, biz) for biz in (False, True)]
    path = str(tmp_path / "index.vmis")
    ix.save(path)
    ld = sb.VMISIndex.load(path)
    for key in ("n_sessions_kept", "n_items", "n_pairs_kept", "n_postings", "max_len", "m_build", "idf_weighting"):
        assert ld.stats()[key] == ix.stats()[key], key
    for biz in (False, True):
        got = sb.predict_batch(ld, q, 100, 400, 21, biz)
        assert all(np.array_equal(x, y) for x, y in zip(got, want[biz]))
    assert ld.idf(int(known[3])) == ix.idf(int(known[3]))
    assert ld.find_attributes(int(known[3])) == {"is_for_sale": True, "is_adult": True}
    bad = tmp_path / "bad.vmis"
    bad.write_bytes(b"not an index")
    with pytest.raises(sb.VmisError) as e:
        sb.VMISIndex.load(str(bad))
    assert e.value.code == -2


def test_blob_loader_rejects_corrupt_files(sb, tmp_path):
    """nothing in a blob is trusted — absurd counts, truncation, flipped payload bytes and out-of-range
    references all fail with VMIS_ERR_IO instead of aborting the process or handing the kernel a bad pointer."""
    items, off, ts = sb.synth_sessions(42, 2000, 8000)
    ix = sb.VMISIndex.from_sessions(items, off, ts, 300, 34, 2.0, device=0)
    path = str(tmp_path / "index.vmis")
    ix.save(path)
    blob = bytearray(open(path, "rb").read())
    # header layout: magic[8], 6 x u32, 7 x u64, f64

    def try_load(data, what):
        p = tmp_path / "c.vmis"
        p.write_bytes(bytes(data))
        with pytest.raises(sb.VmisError) as e:
            sb.VMISIndex.load(str(p))
        assert e.value.code == -2, what

    huge = bytearray(blob); huge[32:40] = (1 << 60).to_bytes(8, "little")            # n_items = 2^60
    try_load(huge, "huge n_items")
    try_load(blob[:len(blob) // 2], "truncated")
    cap = bytearray(blob); cap[40:48] = (int.from_bytes(blob[40:48], "little") - 1).to_bytes(8, "little")
    try_load(cap, "hash capacity not a power of two")
    flip = bytearray(blob); flip[len(blob) // 2] ^= 0x40                                # payload bit flip → checksum
    try_load(flip, "bit flip")
    assert sb.VMISIndex.load(path).stats()["n_items"] == ix.stats()["n_items"]          # the intact file still loads
