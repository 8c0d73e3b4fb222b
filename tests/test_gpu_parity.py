"""Parity of the sm_100a kernel (through the C ABI) against the CPU oracle.  Needs a B200."""
import ctypes as C
import os

import numpy as np
import pytest

from util import csr, evaluator_queries, mrr_hitrate_at, random_index_data, read_test_sessions, same_modulo_ties

pytestmark = pytest.mark.gpu

README_13598 = [2835, 10, 12068, 4313, 3097, 8028, 3545, 7812, 17519, 1164, 17935, 1277, 13335, 8655, 14664, 14556,
                6868, 13509, 9248, 2498, 11724]


def _assert_batch_equal(sb, gix, oix, queries, k, m, n, biz=False):
    """GPU vs canonical oracle: ids bit-exact and in the same order, f64 scores bit-exact."""
    ids, sc, cnt = sb.predict_batch(gix, queries, k, m, n, biz)
    q_items, q_off = csr(queries)
    oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, k, m, n, biz, mode=1)
    assert np.array_equal(cnt, ocnt), f"counts differ at {np.nonzero(cnt != ocnt)[0][:10]}"
    for q in range(len(queries)):
        c = cnt[q]
        if not np.array_equal(ids[q, :c], oids[q, :c]) or not np.array_equal(sc[q, :c], osc[q, :c]):
            raise AssertionError(f"query {q} {queries[q]} k={k} m={m} n={n}:\n gpu {ids[q, :c]} {sc[q, :c]}\n"
                                 f" ora {oids[q, :c]} {osc[q, :c]}")
    return ids, sc, cnt


@pytest.fixture(scope="module")
def toy(sb, oracle, toy_dir):
    train = os.path.join(toy_dir, "train.txt")
    gix = sb.VMISIndex.new_from_csv(train, 1502, 1.0, max_len=15, device=0)
    oix = oracle.OracleIndex.new_from_csv(train, 1502, 1.0, 15)
    tests = read_test_sessions(os.path.join(toy_dir, "test.txt"))
    return gix, oix, tests


def test_kat_should_train_and_predict(sb, oracle):
    # mod.rs:229-310
    items = np.array([920006, 920005, 920004, 920005, 920004, 920003, 920002], dtype=np.uint64)
    off = np.array([0, 3, 7], dtype=np.uint64)
    ts = np.array([1, 1], dtype=np.uint32)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 5, 5, 1.0, device=0)
    recs = sb.predict(gix, [920005], 500, 500, 20, False)
    assert len(recs) == 4
    assert recs[0][0] == 920004
    assert recs[0][1] == pytest.approx(2 * 0.9 * np.log(3.5), rel=1e-12)
    assert [r[0] for r in recs[1:]] == [920002, 920003, 920006]
    for r in recs[1:]:
        assert r[1] == pytest.approx(0.9 * np.log(7.0), rel=1e-12)


def test_readme_golden_response(sb, toy):
    # README.md:131-155 — response of the shipped binary for item 13598 (m=1502,k=288,n=21)
    gix, oix, _ = toy
    recs = sb.predict(gix, [13598], 288, 1502, 21, False)
    ids = [r[0] for r in recs]
    assert sorted(ids) == sorted(README_13598)
    sc = {r[0]: r[1] for r in recs}
    # same order modulo exact-score tie classes
    readme_scores = [sc[i] for i in README_13598]
    assert all(readme_scores[i] >= readme_scores[i + 1] - 1e-12 for i in range(20))
    assert ids[:3] == README_13598[:3]


@pytest.mark.parametrize("k,m,n,L", [(50, 500, 21, 2), (288, 1502, 21, 4), (288, 1502, 21, 1), (1500, 2500, 21, 20),
                                       (5, 10, 3, 4), (100, 100, 40, 3)])
def test_toy_replay_matches_canonical(sb, toy, k, m, n, L):
    gix, oix, tests = toy
    queries, _ = evaluator_queries(tests, L)
    assert len(queries) == 931
    _assert_batch_equal(sb, gix, oix, queries, k, m, n)


def test_toy_faithful_modulo_ties(sb, toy):
    """vs the statement-by-statement restatement: same scores to 1e-12, same ids modulo tie classes
    (k >= m here, so the unpinned k-boundary heuristic of vmis_index.rs:400-410 is not exercised)."""
    gix, oix, tests = toy
    queries, _ = evaluator_queries(tests, 4)
    ids, sc, cnt = sb.predict_batch(gix, queries, 1502, 1502, 21)
    bad = 0
    for q, ev in enumerate(queries):
        fi, fs = oix.predict(ev, 1502, 1502, 21, mode=0)
        if not same_modulo_ties(ids[q, :cnt[q]], sc[q, :cnt[q]], fi, fs):
            bad += 1
    assert bad == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_small_indexes(sb, oracle, seed):
    """tiny m (eviction / truncation), duplicates, unknown items, timestamp ties, negative weights"""
    rng = np.random.default_rng(seed)
    n_sessions, n_items = int(rng.integers(50, 400)), int(rng.integers(8, 60))
    items, off, ts = random_index_data(rng, n_sessions, n_items, max_len=int(rng.integers(2, 9)),
                                       unique_ts=bool(seed % 2), id_scale=int(rng.integers(1, 1 << 20)))
    m_build = int(rng.integers(1, 40))
    max_len = int(rng.integers(3, 9))
    gix = sb.VMISIndex.from_sessions(items, off, ts, m_build, max_len, 1.5, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, m_build, max_len, 1.5)
    known = np.unique(items)
    queries = []
    for _ in range(300):
        L = int(rng.integers(1, 16))
        ev = [int(x) for x in rng.choice(known, size=L, replace=True)]
        if rng.random() < 0.3:
            ev[int(rng.integers(0, L))] = 999_999_999_999  # unknown item
        queries.append(ev)
    queries.append([999_999_999_999])
    queries.append([])
    for k, m, n in [(3, m_build, 5), (1, 1, 1), (500, 500, 50), (7, max(1, m_build // 2), 21), (20, 64, 33)]:
        _assert_batch_equal(sb, gix, oix, queries, k, m, n)


def test_find_neighbors_matches_canonical(sb, oracle):
    rng = np.random.default_rng(11)
    items, off, ts = random_index_data(rng, 600, 40, max_len=6, unique_ts=False)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 50, 6, 1.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 50, 6, 1.0)
    known = np.unique(items)
    queries = [[int(x) for x in rng.choice(known, size=int(rng.integers(1, 7)))] for _ in range(200)]
    for k, m in [(10, 50), (50, 50), (3, 7), (200, 30)]:
        sess, sim, cnt = gix.find_neighbors_batch(queries, k, m)
        for q, ev in enumerate(queries):
            os_, osim = oix.find_neighbors(ev, k, m, mode=1)
            assert cnt[q] == len(os_)
            assert np.array_equal(sess[q, :cnt[q]], os_), (q, ev, k, m)
            assert np.array_equal(sim[q, :cnt[q]], osim)


def test_business_rules(sb, oracle):
    # mod.rs:162-182
    rng = np.random.default_rng(5)
    items, off, ts = random_index_data(rng, 500, 30, max_len=6)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 100, 6, 1.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 100, 6, 1.0)
    known = np.unique(items)
    flags = rng.integers(0, 8, size=len(known)).astype(np.uint8)
    gix.set_attributes(known, flags)
    for it, f in zip(known, flags):
        oix.set_attributes(int(it), exists=bool(f), for_sale=bool(f & 2), adult=bool(f & 4))
        a = gix.find_attributes(int(it))
        assert a == oix.find_attributes(int(it))
    queries = [[int(x) for x in rng.choice(known, size=int(rng.integers(1, 5)))] for _ in range(300)]
    _assert_batch_equal(sb, gix, oix, queries, 50, 100, 21, biz=True)
    _assert_batch_equal(sb, gix, oix, queries, 50, 100, 21, biz=False)


def test_long_sessions_and_limits(sb, oracle):
    rng = np.random.default_rng(9)
    items, off, ts = random_index_data(rng, 800, 150, max_len=8)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 64, 8, 1.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 64, 8, 1.0)
    known = np.unique(items)
    queries = [[int(x) for x in rng.choice(known, size=L)] for L in (100, 128, 127, 64, 33, 101, 120)]
    _assert_batch_equal(sb, gix, oix, queries, 40, 64, 21)
    with pytest.raises(sb.VmisError) as e:
        sb.predict_batch(gix, [[int(known[0])] * 129], 40, 64, 21)
    assert e.value.code == -4
    with pytest.raises(sb.VmisError):
        sb.predict_batch(gix, queries, 4096, 64, 21)


def test_overflow_table_path(sb, oracle):
    """neighbour item lists larger than the shared score table → per-CTA global table"""
    rng = np.random.default_rng(21)
    n_items, n_sessions = 3000, 1500
    items, off = [], [0]
    for s in range(n_sessions):
        tail = rng.choice(np.arange(1, n_items), size=29, replace=False)
        items.extend(sorted([7] + [int(x) * 3 + 11 for x in tail]))   # item 7 is in every session
        off.append(len(items))
    items, off = np.array(items, dtype=np.uint64), np.array(off, dtype=np.uint64)
    ts = rng.permutation(n_sessions).astype(np.uint32)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 1502, 34, 2.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 1502, 34, 2.0)
    queries = [[7], [7, int(items[5])], [int(items[3]), 7, int(items[40])]] * 5
    _assert_batch_equal(sb, gix, oix, queries, 600, 1502, 21)
    _assert_batch_equal(sb, gix, oix, queries, 288, 1502, 70)


def test_global_table_fallback_with_many_distinct_items(sb, oracle):
    """more distinct neighbour items than the shared score table can hold → the CTA's global table (big k)"""
    rng = np.random.default_rng(33)
    n_items, n_sessions = 80_000, 3000
    items, off = [], [0]
    for s in range(n_sessions):
        tail = rng.choice(np.arange(1, n_items), size=29, replace=False)
        items.extend(sorted([5] + [int(x) * 7 + 13 for x in tail]))       # item 5 is in every session
        off.append(len(items))
    items, off = np.array(items, dtype=np.uint64), np.array(off, dtype=np.uint64)
    ts = rng.permutation(n_sessions).astype(np.uint32)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 2500, 34, 1.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 2500, 34, 1.0)
    queries = [[5], [int(items[7]), 5], [5, int(items[40]), int(items[41])]] * 3
    _assert_batch_equal(sb, gix, oix, queries, 2048, 2500, 21)          # ~45 k distinct items per query
    _assert_batch_equal(sb, gix, oix, queries, 2048, 2500, 100)         # exact path with several rounds
    _assert_batch_equal(sb, gix, oix, queries, 288, 1502, 21)           # and back to the shared table


def test_synthetic_config2_batch1024(sb, oracle):
    """BASELINE.json config 2: synthetic 1M interactions / 50k items, batch = 1024 query sessions"""
    items, off, ts = sb.synth_sessions(42, 50_000, 193_000)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 1502, 34, 2.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 1502, 34, 2.0)
    q_items, q_off = sb.synth_queries(43, 50_000, 1024, 4)
    queries = [list(map(int, q_items[q_off[i]:q_off[i + 1]])) for i in range(1024)]
    ids, sc, cnt = _assert_batch_equal(sb, gix, oix, queries, 288, 1502, 21)
    assert (cnt == 21).mean() > 0.95
    # The faithful restatement of the Rust code (heaps + hash maps, mode 0) against the kernel's canonical order.  At
    # k < m the reference keeps "the k best" through a heap whose tie handling depends on hashbrown's iteration order
    # (vmis_index.rs:394-412); single-item sessions are ALL ties there (every candidate has similarity 1), so only the
    # sessions where the boundary is tie-free can agree exactly; a different choice among tied neighbours shifts the
    # scores (a few percent) even where the ranked ids coincide.  Measured on this data (bench.py prints the same
    # figures for config 3): ~14 % identical ranked lists, ~26 % identical top-21 sets, mean set overlap@21 0.88.  The
    # score tolerance proper (1e-5) is asserted where ties cannot interfere: test_faithful_agreement_without_boundary_ties.
    same_list = same_set = 0
    overlap = []
    qs = list(range(0, 1024, 2))
    for q in qs:
        fi, fs = oix.predict(queries[q], 288, 1502, 21, mode=0)
        gi, gs = ids[q, :cnt[q]], sc[q, :cnt[q]]
        a, b = set(gi.tolist()), set(fi.tolist())
        overlap.append(len(a & b) / max(1, max(len(a), len(b))))
        same_set += a == b
        same_list += len(fi) == len(gi) and bool(np.array_equal(fi, gi))
    print(f"canonical vs faithful on {len(qs)} sessions: same ranked list {same_list}, same top-21 set {same_set}, "
          f"mean overlap@21 {np.mean(overlap):.3f}")
    assert same_list >= len(qs) // 20 and same_set >= len(qs) // 8 and np.mean(overlap) >= 0.8


def test_faithful_agreement_without_boundary_ties(sb, oracle):
    """Where the reference's result does not depend on unpinned tie order — k >= m, unique timestamps — the faithful
    restatement and the kernel must agree on the ranked ids (modulo exact-score ties) and on the scores to 1e-5."""
    items, off, ts = sb.synth_sessions(42, 50_000, 193_000)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 300, 34, 2.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 300, 34, 2.0)
    q_items, q_off = sb.synth_queries(47, 50_000, 512, 4)
    ids, sc, cnt = sb.predict_batch(gix, (q_items, q_off), 300, 300, 21)
    agree = 0
    for q in range(512):
        fi, fs = oix.predict(q_items[q_off[q]:q_off[q + 1]], 300, 300, 21, mode=0)
        assert len(fi) == cnt[q]
        gs = sc[q, :cnt[q]]
        # same scores position by position (ties may permute ids inside a tie class, never scores)
        assert np.allclose(np.sort(fs)[::-1], gs, rtol=1e-5, atol=0)
        agree += bool(np.array_equal(fi, ids[q, :cnt[q]]))
        # ids equal as sets except inside the last tie class, which the cut at 21 may split differently
        last = gs[-1] if cnt[q] else 0.0
        a = {int(i) for i, s_ in zip(ids[q, :cnt[q]], gs) if s_ > last * (1 + 1e-12)}
        b = {int(i) for i, s_ in zip(fi, fs) if s_ > last * (1 + 1e-12)}
        assert a == b
    assert agree >= 256


def test_toy_quality_faithful_vs_kernel(sb, oracle, toy, toy_dir):
    """README HPO optimum on the toy data (k=288, m=1502, last 4 items): MRR@20 / HitRate@20 of the kernel next to the
    faithful restatement — the distance a user switching from the reference would see (vmis_index.rs:394-412,
    mod.rs:185-212 leave the order of ties to hashbrown)."""
    gix, oix, tests = toy
    queries, rest = evaluator_queries(tests, 4)
    ids, sc, cnt = sb.predict_batch(gix, queries, 288, 1502, 21)
    g_recs = [ids[q, :cnt[q]].tolist() for q in range(len(queries))]
    f_recs = [oix.predict(q, 288, 1502, 21, mode=0)[0].tolist() for q in queries]
    g_mrr, g_hr = mrr_hitrate_at(g_recs, rest)
    f_mrr, f_hr = mrr_hitrate_at(f_recs, rest)
    overlap = np.mean([len(set(a) & set(b)) / max(1, max(len(a), len(b))) for a, b in zip(g_recs, f_recs)])
    print(f"toy k=288 m=1502: kernel MRR@20 {g_mrr:.4f} HR@20 {g_hr:.4f} | faithful {f_mrr:.4f} {f_hr:.4f} | overlap@21 {overlap:.3f}")
    assert abs(g_mrr - f_mrr) <= 0.006 and abs(g_hr - f_hr) <= 0.012 and overlap >= 0.85


def test_device_api_and_stats(sb, oracle):
    torch = pytest.importorskip("torch")
    items, off, ts = sb.synth_sessions(42, 20_000, 60_000)
    gix = sb.VMISIndex.from_sessions(items, off, ts, 1502, 34, 2.0, device=0)
    q_items, q_off = sb.synth_queries(43, 20_000, 4096, 4)
    n_q, n = 4096, 21
    dev = torch.device("cuda:0")
    d_items = torch.from_numpy(q_items.view(np.int64)).to(dev)
    d_off = torch.from_numpy(q_off.view(np.int32)).to(dev)
    d_ids = torch.zeros((n_q, n), dtype=torch.int64, device=dev)
    d_sc = torch.zeros((n_q, n), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(n_q, dtype=torch.int32, device=dev)
    d_st = torch.zeros((n_q, 4), dtype=torch.int32, device=dev)
    lib = sb.load_library()
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.vmis_predict_batch_device(gix.handle, d_items.data_ptr(), d_off.data_ptr(), n_q, 288, 1502, n, 0,
                                       d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), d_st.data_ptr(),
                                       C.c_void_p(stream))
    assert rc == 0, lib.vmis_last_error()
    torch.cuda.synchronize()
    ids, sc, cnt = sb.predict_batch(gix, (q_items, q_off), 288, 1502, n)
    assert np.array_equal(d_cnt.cpu().numpy().view(np.uint32), cnt)
    assert np.array_equal(d_ids.cpu().numpy().view(np.uint64), ids)
    assert np.array_equal(d_sc.cpu().numpy(), sc)
    st = d_st.cpu().numpy()
    assert np.array_equal(st[:, 3].astype(np.uint32), cnt)
    assert (st[:, 1] <= 288).all() and (st[:, 0] >= st[:, 1]).all()


def test_concurrent_host_threads(sb, toy):
    """many host threads on one index (actix workers over Arc<VMISIndex>, serving.rs:62-94)"""
    import threading
    gix, oix, tests = toy
    queries, _ = evaluator_queries(tests, 4)
    ref = sb.predict_batch(gix, queries, 288, 1502, 21)
    errs = []

    def work():
        try:
            for _ in range(5):
                got = sb.predict_batch(gix, queries, 288, 1502, 21)
                assert all(np.array_equal(a, b) for a, b in zip(got, ref))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_batched_replay_evaluator_matches_readme(sb, toy, toy_dir):
    """README.md:166-177 (evaluator on the toy data, example.toml: m=500, k=50, max_items=2, n=21):
    931 evaluations, HitRate@20 0.6402, Mrr@20 0.3277 (tie order unpinned in the reference → ±0.004)."""
    from serenade_b200.evaluate import evaluate, read_test_sessions as rts
    gix, _, _ = toy
    res = evaluate(gix, rts(os.path.join(toy_dir, "test.txt")), 50, 500, 21, 2, 20)
    assert res["qty_evaluations"] == 931
    assert res["hitrate"] == pytest.approx(0.6402, abs=0.0006)
    assert res["mrr"] == pytest.approx(0.3277, abs=0.004)
    res = evaluate(gix, rts(os.path.join(toy_dir, "test.txt")), 288, 1502, 21, 4, 20)   # README HPO optimum
    assert res["mrr"] == pytest.approx(0.3401, abs=0.004) and res["hitrate"] > 0.65


def test_full_size_properties_config3(sb, oracle):
    """BASELINE config 3 at full size (60 M interactions / 1.76 M items, index generated and built on the device):
    size-independent properties + oracle parity of a same-generator mid-size slice is covered elsewhere."""
    gix = sb.VMISIndex.synth(42, 1_760_000, 11_556_000, 1502, 34, 2.0)
    st = gix.stats()
    assert st["n_pairs_kept"] == 60_017_474 and st["n_items"] == 1_721_361 and st["n_postings"] == 31_637_718
    q_items, q_off = sb.synth_queries(43, 1_760_000, 1 << 15, 4)
    n_q = len(q_off) - 1
    ids, sc, cnt = sb.predict_batch(gix, (q_items, q_off), 288, 1502, 21)
    # idempotence
    ids2, sc2, cnt2 = sb.predict_batch(gix, (q_items, q_off), 288, 1502, 21)
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2) and np.array_equal(cnt, cnt2)
    # permutation of the batch permutes the rows
    perm = np.random.default_rng(0).permutation(n_q)
    sessions = [q_items[q_off[i]:q_off[i + 1]] for i in range(n_q)]
    p_ids, p_sc, p_cnt = sb.predict_batch(gix, [sessions[i] for i in perm], 288, 1502, 21)
    assert np.array_equal(p_ids, ids[perm]) and np.array_equal(p_sc, sc[perm]) and np.array_equal(p_cnt, cnt[perm])
    # order (score desc, id asc), no duplicates, current item never recommended, padding is zero
    for q in range(0, n_q, 97):
        c = cnt[q]
        s, i = sc[q, :c], ids[q, :c]
        assert (np.diff(s) <= 0).all()
        ties = np.diff(s) == 0
        assert (np.diff(i.astype(np.int64))[ties] > 0).all()
        assert len(np.unique(i)) == c and sessions[q][-1] not in set(i.tolist())
        assert (ids[q, c:] == 0).all() and (sc[q, c:] == 0).all()
    # how_many prefix property: top-5 is the prefix of top-21; k/m monotone sanity on counts
    ids5, sc5, cnt5 = sb.predict_batch(gix, (q_items, q_off), 288, 1502, 5)
    assert np.array_equal(ids5, ids[:, :5] * (np.arange(5)[None, :] < cnt5[:, None]))


def test_full_size_config3_vs_oracle(sb, oracle):
    """BASELINE config 3 at FULL size against the canonical oracle: 60 M interactions / 1.76 M items, 16 384 evolving
    sessions, k=288, m=1502 — ids, order, f64 score bits and counts identical.  (Host-generated sessions so that the
    oracle sees the same data; the device-generated index of the same seed must then answer identically too.)"""
    items, off, ts = sb.synth_sessions(42, 1_760_000, 11_556_000)
    assert len(items) == 60_017_474
    gix = sb.VMISIndex.from_sessions(items, off, ts, 1502, 34, 2.0, device=0)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 1502, 34, 2.0)
    q_items, q_off = sb.synth_queries(4242, 1_760_000, 1 << 14, 4)
    ids, sc, cnt = sb.predict_batch(gix, (q_items, q_off), 288, 1502, 21)
    oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, 288, 1502, 21, mode=1, threads=os.cpu_count() or 8)
    assert np.array_equal(cnt, ocnt) and np.array_equal(ids, oids) and np.array_equal(sc, osc)
    del oix
    dix = sb.VMISIndex.synth(42, 1_760_000, 11_556_000, 1502, 34, 2.0)
    ids2, sc2, cnt2 = sb.predict_batch(dix, (q_items, q_off), 288, 1502, 21)
    assert np.array_equal(cnt, cnt2) and np.array_equal(ids, ids2) and np.array_equal(sc, sc2)


def test_config4_size_sharded_equals_unsharded(sb):
    """BASELINE config 4 size (582 M interactions / 6.5 M items, device-generated): the index with its postings split
    into two item shards (cross-attached by pointer on one GPU — the layout of config 5) answers 65 536 evolving
    sessions exactly like the unsharded index."""
    n_items, n_sessions = 6_500_000, 112_100_000
    ref = sb.VMISIndex.synth(42, n_items, n_sessions, 1502, 34, 2.0)
    assert ref.stats()["n_pairs_kept"] > 580_000_000
    q_items, q_off = sb.synth_queries(99, n_items, 1 << 16, 4)
    ids, sc, cnt = sb.predict_batch(ref, (q_items, q_off), 288, 1502, 21)
    assert (cnt == 21).mean() > 0.9
    ref.close()
    shards = [sb.VMISIndex.synth(42, n_items, n_sessions, 1502, 34, 2.0, 0, s, 2) for s in range(2)]
    shards[0].attach_shard_ptr(1, shards[1].shard_ptr())
    shards[1].attach_shard_ptr(0, shards[0].shard_ptr())
    for sh in shards:
        ids2, sc2, cnt2 = sb.predict_batch(sh, (q_items, q_off), 288, 1502, 21)
        assert np.array_equal(cnt, cnt2) and np.array_equal(ids, ids2) and np.array_equal(sc, sc2)


def test_device_api_flags_over_long_sessions(sb):
    """vmis_predict_batch_device cannot reject a launch: a session beyond VMIS_MAX_SESSION_LEN comes back with the
    sentinel count; the host API refuses the same batch up front (VMIS_ERR_LIMIT)."""
    torch = pytest.importorskip("torch")
    gix = sb.VMISIndex.synth(42, 20_000, 60_000, 1502, 34, 2.0)
    good_items, good_off = sb.synth_queries(43, 20_000, 3, 4)
    long_session = np.arange(1, 131, dtype=np.uint64)
    q_items = np.concatenate([good_items[:good_off[1]], long_session, good_items[good_off[1]:]])
    q_off = np.concatenate([[0, good_off[1]], good_off[1:] + 130]).astype(np.uint32)
    with pytest.raises(sb.VmisError) as e:
        sb.predict_batch(gix, (q_items, q_off), 288, 1502, 21)
    assert e.value.code == -4
    dev = torch.device("cuda:0")
    n_q, n = 4, 21
    d_items = torch.from_numpy(q_items.view(np.int64)).to(dev)
    d_off = torch.from_numpy(q_off.view(np.int32)).to(dev)
    d_ids = torch.zeros((n_q, n), dtype=torch.int64, device=dev)
    d_sc = torch.zeros((n_q, n), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(n_q, dtype=torch.int32, device=dev)
    lib = sb.load_library()
    rc = lib.vmis_predict_batch_device(gix.handle, d_items.data_ptr(), d_off.data_ptr(), n_q, 288, 1502, n, 0,
                                       d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), None,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy().view(np.uint32)
    assert cnt[1] == 0xFFFFFFFF and (d_ids[1] == 0).all()
    ids, sc, c = sb.predict_batch(gix, (good_items, good_off), 288, 1502, n)
    assert np.array_equal(cnt[[0, 2, 3]], c) and np.array_equal(d_ids.cpu().numpy().view(np.uint64)[[0, 2, 3]], ids)


def test_chunked_host_api_matches_device_api(sb):
    """batches larger than one pipeline chunk (2^17) go through a ring of streams; rows must not move"""
    torch = pytest.importorskip("torch")
    gix = sb.VMISIndex.synth(42, 20_000, 60_000, 1502, 34, 2.0)
    n_q, n = 300_001, 21
    q_items, q_off = sb.synth_queries(43, 20_000, n_q, 4)
    ids, sc, cnt = sb.predict_batch(gix, (q_items, q_off), 288, 1502, n)
    dev = torch.device("cuda:0")
    d_items = torch.from_numpy(q_items.view(np.int64)).to(dev)
    d_off = torch.from_numpy(q_off.view(np.int32)).to(dev)
    d_ids = torch.zeros((n_q, n), dtype=torch.int64, device=dev)
    d_sc = torch.zeros((n_q, n), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(n_q, dtype=torch.int32, device=dev)
    lib = sb.load_library()
    rc = lib.vmis_predict_batch_device(gix.handle, d_items.data_ptr(), d_off.data_ptr(), n_q, 288, 1502, n, 0,
                                       d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), None, None)
    assert rc == 0, lib.vmis_last_error()
    torch.cuda.synchronize()
    assert np.array_equal(d_cnt.cpu().numpy().view(np.uint32), cnt)
    assert np.array_equal(d_ids.cpu().numpy().view(np.uint64), ids)
    assert np.array_equal(d_sc.cpu().numpy(), sc)
    sess, sim, ncnt = gix.find_neighbors_batch((q_items[:q_off[140_000]], q_off[:140_001]), 50, 1502)
    s2, m2, c2 = gix.find_neighbors_batch((q_items[q_off[131_000]:q_off[140_000]], q_off[131_000:140_001] - q_off[131_000]), 50, 1502)
    assert np.array_equal(sess[131_000:], s2) and np.array_equal(sim[131_000:], m2) and np.array_equal(ncnt[131_000:], c2)


def test_small_batch_flag_path_matches_device_api(sb):
    """batches small enough for the mapped pinned buffer: rows are handed over through per-row completion flags while
    the kernel still runs (capi.cu latency path); sizes around the 16-session zero-copy-input limit and the 128-row
    hand-over step, then lone callers from several threads at once (pooled call contexts)"""
    import threading
    torch = pytest.importorskip("torch")
    gix = sb.VMISIndex.synth(42, 20_000, 60_000, 1502, 34, 2.0)
    n_all, n = 2600, 21
    q_items, q_off = sb.synth_queries(47, 20_000, n_all, 4)
    dev = torch.device("cuda:0")
    d_items = torch.from_numpy(q_items.view(np.int64)).to(dev)
    d_off = torch.from_numpy(q_off.view(np.int32)).to(dev)
    d_ids = torch.zeros((n_all, n), dtype=torch.int64, device=dev)
    d_sc = torch.zeros((n_all, n), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(n_all, dtype=torch.int32, device=dev)
    lib = sb.load_library()
    rc = lib.vmis_predict_batch_device(gix.handle, d_items.data_ptr(), d_off.data_ptr(), n_all, 288, 1502, n, 0,
                                       d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), None, None)
    assert rc == 0, lib.vmis_last_error()
    torch.cuda.synchronize()
    r_ids, r_sc, r_cnt = d_ids.cpu().numpy().view(np.uint64), d_sc.cpu().numpy(), d_cnt.cpu().numpy().view(np.uint32)
    for rep in range(3):                                   # pooled contexts and their buffers are reused
        for n_q in (1, 2, 16, 17, 127, 128, 129, 1024, 2600):
            lo = (rep * 37) % (n_all - n_q + 1)
            sub = (q_items[q_off[lo]:q_off[lo + n_q]], q_off[lo:lo + n_q + 1] - q_off[lo])
            ids, sc, cnt = sb.predict_batch(gix, sub, 288, 1502, n)
            assert np.array_equal(cnt, r_cnt[lo:lo + n_q]), (n_q, rep)
            assert np.array_equal(ids, r_ids[lo:lo + n_q]) and np.array_equal(sc, r_sc[lo:lo + n_q]), (n_q, rep)
    errs = []

    def lone_caller(t):
        try:
            for q in range(t, 400, 8):
                recs = sb.predict(gix, q_items[q_off[q]:q_off[q + 1]], 288, 1502, n, False)
                c = int(r_cnt[q])
                assert [r[0] for r in recs] == list(r_ids[q, :c]) and [r[1] for r in recs] == list(r_sc[q, :c]), q
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=lone_caller, args=(t,)) for t in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_cpp_evaluator_tool(toy_dir, tmp_path):
    """tools/evaluator.cpp (the reference's evaluator binary over include/vmis.hpp): KAT + README run"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "evaluator"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(root, "tools", "evaluator.cpp"),
                           "-L" + os.path.join(root, "serenade_b200"), "-lvmis_b200",
                           "-Wl,-rpath," + os.path.join(root, "serenade_b200")])
    out = subprocess.run([str(exe), "--kat"], capture_output=True, text=True)
    assert out.returncode == 0 and "KAT ok: 4 recommendations, first 920004" in out.stdout, out.stdout + out.stderr
    out = subprocess.run([str(exe), os.path.join(toy_dir, "train.txt"), os.path.join(toy_dir, "test.txt"), "500", "50", "21", "2", "1"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "Qty test evaluations: 931" in out.stdout
    header = "Mrr@20,Ndcg@20,HitRate@20,Popularity@20,Precision@20,Coverage@20,Recall@20,F1score@20\n"   # README.md:170
    got = [float(x) for x in out.stdout.split(header)[1].split("\n")[0].split(",")]
    for g, w in zip(got, [0.3277, 0.3553, 0.6402, 0.0499, 0.0680, 0.2765, 0.4456, 0.1180]):          # README.md:171
        assert g == pytest.approx(w, abs=0.004)
    assert got[2] == pytest.approx(0.6402, abs=0.0006)


def test_micro_batcher_online_shape(sb, toy):
    """one evolving session per caller thread (actix workers, recommend_resource.rs:56) through the micro-batcher"""
    import threading
    gix, oix, tests = toy
    queries, _ = evaluator_queries(tests, 4)
    want = sb.predict_batch(gix, queries, 288, 1502, 21)
    b = sb.Batcher(gix, 288, 1502, 21, False, max_batch=256, max_wait_us=500)
    errs = []

    def work(t):
        try:
            for q in range(t, len(queries), 16):
                got = b.predict(queries[q])
                c = want[2][q]
                assert [g[0] for g in got] == want[0][q, :c].tolist() and [g[1] for g in got] == want[1][q, :c].tolist()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(16)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs[:2]
    st = b.stats()
    assert st["requests"] == len(queries) and st["batches"] < len(queries)      # requests really were coalesced
    assert b.predict([]) == []
    b.close()


def test_serving_shell_recommend(sb, toy):
    """GET /v1/recommend minus HTTP: the README request (README.md:131-155) and concurrent visitors whose evolving
    sessions grow request by request (recommend_resource.rs:39-64) — every answer equals predict() on the window"""
    import threading
    gix, oix, tests = toy
    srv = sb.Server(gix, 288, 1502, 21, max_items_in_session=4, max_batch=64, max_wait_us=300)
    ids = srv.recommend("144", 13598, user_consent=True)
    assert sorted(ids) == sorted(README_13598) and ids[:3] == README_13598[:3]
    sessions = [tests[s] for s in sorted(tests)][:64]
    errs = []

    def visitor(v):
        try:
            window = []
            for item in sessions[v]:
                if not window:
                    window = [item]
                elif window[-1] != item:
                    window = (window + [item])[-4:]
                got = srv.recommend("visitor-%d" % v, item)
                want = [r[0] for r in sb.predict(gix, window, 288, 1502, 21)]
                assert got == want, (v, window)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=visitor, args=(v,)) for v in range(len(sessions))]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs[:2]
    assert srv.recommend("nobody", 13598, user_consent=False) == ids
    srv.close()


def test_evaluator_line_all_eight_metrics(sb, toy, toy_dir):
    """the line the reference's evaluator prints (README.md:170-171), from ONE batched predict through the kernel"""
    from serenade_b200.evaluate import evaluate_all, read_test_sessions as rts, read_training_items
    train = os.path.join(toy_dir, "train.txt")
    gix = sb.VMISIndex.new_from_csv(train, 500, 1.0, max_len=15, device=0)
    res = evaluate_all(gix, read_training_items(train), rts(os.path.join(toy_dir, "test.txt")), 50, 500, 21, 2, 20)
    assert res["qty_evaluations"] == 931
    want = {"Mrr@20": 0.3277, "Ndcg@20": 0.3553, "HitRate@20": 0.6402, "Popularity@20": 0.0499, "Precision@20": 0.0680,
            "Coverage@20": 0.2765, "Recall@20": 0.4456, "F1score@20": 0.1180}
    for name, v in want.items():
        assert res[name] == pytest.approx(v, abs=0.004), (name, res[name], v)
