"""GPU parity of the pre-computed (Avro) index path: VMISIndex::new (vmis_index.rs:85-313) loaded into HBM and
queried through the kernel vs the CPU oracle holding the very same posting lists / idf / attributes."""
import numpy as np
import pytest

import avro_util as au
from util import csr, random_index_data

pytestmark = pytest.mark.gpu


def _equal(sb, gix, oix, queries, k, m, n, biz=False):
    ids, sc, cnt = sb.predict_batch(gix, queries, k, m, n, biz)
    q_items, q_off = csr(queries)
    oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, k, m, n, biz, mode=1)
    assert np.array_equal(cnt, ocnt), f"counts differ at {np.nonzero(cnt != ocnt)[0][:10]}"
    for q in range(len(queries)):
        c = cnt[q]
        if not np.array_equal(ids[q, :c], oids[q, :c]) or not np.array_equal(sc[q, :c], osc[q, :c]):
            raise AssertionError(f"query {q} {queries[q]} k={k} m={m} n={n} biz={biz}:\n gpu {ids[q, :c]} {sc[q, :c]}\n"
                                 f" ora {oids[q, :c]} {osc[q, :c]}")


def _queries(rng, known, n=300):
    qs = []
    for _ in range(n):
        L = int(rng.integers(1, 12))
        ev = [int(x) for x in rng.choice(known, size=L, replace=True)]
        if rng.random() < 0.2:
            ev[int(rng.integers(0, L))] = 999_999_999_999
        qs.append(ev)
    return qs


def _oracle_from_parts(oracle, p):
    return oracle.OracleIndex.from_parts(p["item_ids"], p["post_off"], p["post_sessions"], p["idf"], p["attr"], p["items"],
                                         p["off"], p["ts"])


@pytest.mark.parametrize("seed,style,codec", [(0, "plain", "null"), (1, "spark", "snappy"), (2, "spark", "deflate")])
def test_avro_index_predict_parity(sb, oracle, tmp_path, seed, style, codec):
    rng = np.random.default_rng(100 + seed)
    items, off, ts = random_index_data(rng, 500, 50, max_len=8, unique_ts=bool(seed % 2), id_scale=977)
    m_build = 20
    src = oracle.OracleIndex.from_sessions(items, off, ts, m_build, 6, 2.0)
    parts = au.parts_from_oracle(src, items, off, ts)
    parts["attr"] = np.array([1 | (2 if i % 4 else 0) | (4 if i % 3 == 0 else 0) for i in range(len(parts["item_ids"]))],
                             dtype=np.uint8)
    au.write_index_dir(str(tmp_path), parts, style=style, codec=codec, files=2, records_per_block=23)
    gix = sb.VMISIndex.new(str(tmp_path), device=0)
    oix = _oracle_from_parts(oracle, parts)
    qs = _queries(rng, np.unique(items))
    # m <= m_carry: first-match positions carried through the merges; m > m_carry: item lists scanned (mod.rs:133-138)
    assert gix.prebuilt_info()["m_carry"] == m_build
    for k, m, n, biz in [(5, m_build, 21, False), (288, 1502, 21, False), (7, 10, 5, True), (50, 64, 33, True), (1, 1, 1, False)]:
        _equal(sb, gix, oix, qs, k, m, n, biz)
    # the index built here from the same sessions gives the same answers (same lists, same idf doubles)
    gix2 = sb.VMISIndex.from_sessions(items, off, ts, m_build, 6, 2.0, device=0)
    a = sb.predict_batch(gix, qs, 5, m_build, 21)
    b = sb.predict_batch(gix2, qs, 5, m_build, 21)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_prebuilt_lists_in_any_order_and_not_prefixes(sb, oracle):
    """posting lists shuffled (normalised at load) and thinned at random (not most-recent prefixes → m_carry 0 →
    the kernel scans the item lists for the first-match position like the reference)"""
    rng = np.random.default_rng(77)
    items, off, ts = random_index_data(rng, 400, 40, max_len=7, id_scale=31)
    src = oracle.OracleIndex.from_sessions(items, off, ts, 30, 7, 1.0)
    parts = au.parts_from_oracle(src, items, off, ts)
    po, ps = parts["post_off"], parts["post_sessions"]
    new_ps, new_po = [], [0]
    for i in range(len(parts["item_ids"])):
        lst = ps[int(po[i]):int(po[i + 1])].copy()
        keep = rng.random(len(lst)) < 0.7
        keep[rng.integers(0, len(lst))] = True
        lst = lst[keep]
        rng.shuffle(lst)
        new_ps.extend(int(x) for x in lst)
        new_po.append(len(new_ps))
    parts["post_sessions"] = np.array(new_ps, dtype=np.uint32)
    parts["post_off"] = np.array(new_po, dtype=np.uint64)
    gix = sb.VMISIndex.from_parts(parts["item_ids"], parts["post_off"], parts["post_sessions"], parts["idf"], parts["attr"],
                                  parts["items"], parts["off"], parts["ts"], device=0)
    info = gix.prebuilt_info()
    assert info["m_carry"] == 0 and info["lists_reordered"] > 0
    oix = _oracle_from_parts(oracle, parts)
    qs = _queries(rng, np.unique(items))
    for k, m, n in [(5, 30, 21), (288, 1502, 21), (3, 4, 5)]:
        _equal(sb, gix, oix, qs, k, m, n)


def test_prebuilt_save_load_keeps_m_carry(sb, oracle, tmp_path):
    rng = np.random.default_rng(78)
    items, off, ts = random_index_data(rng, 300, 30, max_len=6, id_scale=5)
    src = oracle.OracleIndex.from_sessions(items, off, ts, 16, 6, 1.0)
    parts = au.parts_from_oracle(src, items, off, ts)
    gix = sb.VMISIndex.from_parts(parts["item_ids"], parts["post_off"], parts["post_sessions"], parts["idf"], parts["attr"],
                                  parts["items"], parts["off"], parts["ts"], device=0)
    blob = str(tmp_path / "ix.blob")
    gix.save(blob)
    gl = sb.VMISIndex.load(blob, device=0)
    assert gl.prebuilt_info()["m_carry"] == 16
    qs = _queries(rng, np.unique(items), 100)
    a = sb.predict_batch(gix, qs, 9, 16, 21)
    b = sb.predict_batch(gl, qs, 9, 16, 21)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_avro_index_item_sharded(sb, oracle, tmp_path):
    """the Avro index with its posting lists split by item over 3 handles (peers attached by pointer): same answers"""
    rng = np.random.default_rng(79)
    items, off, ts = random_index_data(rng, 600, 60, max_len=7, id_scale=13)
    src = oracle.OracleIndex.from_sessions(items, off, ts, 25, 7, 2.0)
    parts = au.parts_from_oracle(src, items, off, ts)
    au.write_index_dir(str(tmp_path), parts, style="spark", codec="deflate", files=2)
    shards = [sb.VMISIndex.new(str(tmp_path), device=0, shard=s, n_shards=3) for s in range(3)]
    for a in range(3):
        for b in range(3):
            if a != b:
                shards[a].attach_shard_ptr(b, shards[b].shard_ptr())
    oix = _oracle_from_parts(oracle, parts)
    qs = _queries(rng, np.unique(items), 200)
    for sh in shards:
        _equal(sb, sh, oix, qs, 20, 25, 21)


def test_device_built_index_exports_and_reloads(sb, oracle, tmp_path):
    """index built on the device → vmis_index_to_avro → VMISIndex::new: the reloaded index answers identically"""
    items, off, ts = sb.synth_sessions(42, 3000, 30000)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 300, 34, 2.0, device=0)
    gix.to_avro(str(tmp_path), "deflate", 4)
    back = sb.VMISIndex.new(str(tmp_path), device=0)
    assert back.prebuilt_info()["m_carry"] == 300 and back.prebuilt_info()["lists_reordered"] == 0
    q = sb.synth_queries(43, 3000, 2000, 4)
    a = sb.predict_batch(gix, q, 100, 300, 21)
    b = sb.predict_batch(back, q, 100, 300, 21)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_long_sessions_in_a_prebuilt_index(sb, oracle):
    """the offline index keeps every session (the reference's production statistics name one of 9408
    events).  Without a bound, a query whose worst-case score table (k x longest session) exceeds the workspace limit
    is refused with a message that names the session length and the remedy; with max_session_len the long sessions
    leave the posting lists (as prepare_hashmap does, vmis_index.rs:452) and the index answers like the oracle holding
    the pruned lists."""
    rng = np.random.default_rng(7)
    items, off, ts = random_index_data(rng, 400, 60, max_len=8, id_scale=31)
    # one giant session: 300 000 distinct items (58 of them shared with the small sessions)
    giant = np.concatenate([np.unique(items)[:58], np.arange(10_000_000, 10_000_000 + 299_942, dtype=np.uint64)])
    items = np.concatenate([items, np.sort(giant)])
    off = np.concatenate([off, [off[-1] + len(giant)]]).astype(np.uint64)
    ts = np.concatenate([ts, [5_000_000]]).astype(np.uint32)
    src = oracle.OracleIndex.from_sessions(items, off, ts, 50, 400_000, 2.0)
    parts = au.parts_from_oracle(src, items, off, ts)
    args = (parts["item_ids"], parts["post_off"], parts["post_sessions"], parts["idf"], None, parts["items"], parts["off"],
            parts["ts"])
    full = sb.VMISIndex.from_parts(*args, device=0)
    qs = _queries(rng, np.unique(items)[:58], 64)
    with pytest.raises(sb.VmisError) as e:
        sb.predict_batch(full, qs, 2000, 50, 21)
    assert e.value.code == -4 and "300000 items" in str(e.value) and "max_session_len" in str(e.value)
    _equal(sb, full, _oracle_from_parts(oracle, parts), qs, 20, 50, 21)      # a small k still fits: the giant is scored
    pruned = sb.VMISIndex.from_parts(*args, device=0, max_session_len=8)
    assert pruned.prebuilt_info()["pruned_postings"] == 300_000
    keep = np.array([s != 400 for s in parts["post_sessions"]])
    p2 = dict(parts)
    cnt = np.add.reduceat(keep.astype(np.int64), parts["post_off"][:-1].astype(np.int64)) if len(keep) else np.zeros(0, np.int64)
    cnt[np.diff(parts["post_off"].astype(np.int64)) == 0] = 0
    p2["post_sessions"] = parts["post_sessions"][keep]
    p2["post_off"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    _equal(sb, pruned, _oracle_from_parts(oracle, p2), qs, 2000, 50, 21)
    _equal(sb, pruned, _oracle_from_parts(oracle, p2), qs, 20, 50, 21)
