"""CPU tests pinning the oracle (oracle/vmis_oracle.cpp) against every known answer the reference holds for the
predict_next path.  No GPU needed."""
import os

import numpy as np
import pytest

from util import evaluator_queries, mrr_hitrate_at, random_index_data, read_test_sessions, same_modulo_ties

README_13598 = [2835, 10, 12068, 4313, 3097, 8028, 3545, 7812, 17519, 1164, 17935, 1277, 13335, 8655, 14664, 14556,
                6868, 13509, 9248, 2498, 11724]
# canonical scores of the README response, tie classes in braces
README_SCORES = {2835: 10.133981736, 10: 9.510149274, 12068: 9.145230676, 3097: 8.886316811, 4313: 8.886316811,
                 8028: 8.886316811, 3545: 8.685487615, 7812: 8.685487615, 1164: 8.521398214, 17519: 8.521398214,
                 17935: 8.382662602, 1277: 8.262484349, 13335: 8.262484349, 8655: 8.156479617, 14556: 7.975875991,
                 14664: 7.975875991, 6868: 7.897565751, 13509: 7.897565751, 9248: 7.825527314, 2498: 7.758830139,
                 11724: 7.696736555}


@pytest.fixture(scope="module")
def toy(oracle, toy_dir):
    ix = oracle.OracleIndex.new_from_csv(os.path.join(toy_dir, "train.txt"), 1502, 1.0, 15)
    tests = read_test_sessions(os.path.join(toy_dir, "test.txt"))
    return ix, tests


def test_heap_ordering_kats(oracle):
    # mod.rs:313-336, :339-356, :360-383, :386-411
    assert oracle.heap_kat(0) == [234, 123]
    assert oracle.heap_kat(1) == [123, 234, 543]
    assert oracle.heap_kat(2) == [234, 123]
    assert oracle.heap_kat(3) == [234, 123]


@pytest.mark.parametrize("mode", [0, 1])
def test_should_train_and_predict(oracle, mode):
    # mod.rs:229-310 — the reference's only known-answer test through prepare_hashmap → find_neighbors → predict
    items = np.array([920006, 920005, 920004, 920005, 920004, 920003, 920002], dtype=np.uint64)
    off = np.array([0, 3, 7], dtype=np.uint64)
    ts = np.array([1, 1], dtype=np.uint32)
    ix = oracle.OracleIndex.from_sessions(items, off, ts, 5, 5, 1.0)
    assert list(ix.postings(920005)) == [1, 0]          # tie on ts → higher session idx first (vmis_index.rs:497-503)
    ids, sc = ix.predict([920005], 500, 500, 20, mode=mode)
    assert len(ids) == 4                                # mod.rs:300
    assert ids[0] == 920004                             # mod.rs:309
    assert sc[0] == pytest.approx(2 * 0.9 * np.log(3.5), rel=1e-14)
    assert sorted(ids[1:]) == [920002, 920003, 920006]
    assert np.allclose(sc[1:], 0.9 * np.log(7.0), rtol=1e-14)


def test_toy_index_statistics(toy):
    ix, _ = toy
    assert ix.num_sessions == 23753
    assert ix.kept_pairs == 77651
    assert ix.num_items == 17919
    # postings(13598): 4 sessions, timestamps strictly decreasing
    p = ix.postings(13598)
    assert len(p) == 4
    ts = [ix.session_ts(int(s)) for s in p]
    assert ts == [1592741644, 1592144620, 1591972635, 1591712974]
    assert list(ix.items_for_session(int(p[0]))) == [8655, 12068, 13509, 13598, 14556]
    assert list(ix.items_for_session(int(p[1]))) == [10, 2835, 8028, 13598]


@pytest.mark.parametrize("mode", [0, 1])
def test_readme_golden_response(toy, mode):
    """README.md:131-155: the JSON the shipped binary returns for item 13598 (m=1502, k=288, n=21)."""
    ix, _ = toy
    ids, sc = ix.predict([13598], 288, 1502, 21, mode=mode)
    assert sorted(map(int, ids)) == sorted(README_13598)
    for i, s in zip(ids, sc):
        assert s == pytest.approx(README_SCORES[int(i)], rel=1e-9)
    # README order == ours modulo exact-score tie classes
    assert same_modulo_ties(np.array(README_13598), np.array([README_SCORES[i] for i in README_13598]), ids, sc,
                            rtol=1e-9)
    # the same list results for m=500 / k=50
    ids2, _ = ix.predict([13598], 50, 500, 21, mode=mode)
    assert sorted(map(int, ids2)) == sorted(README_13598)


@pytest.mark.parametrize("mode", [0, 1])
def test_readme_evaluator_run(toy, mode):
    """README.md:166-177: evaluator on the toy data with the shipped example.toml (m=500, k=50, max_items=2, n=21):
    931 evaluations, HitRate@20 0.6402, Mrr@20 0.3277 (approximate: tie order is unpinned in the reference)."""
    ix, tests = toy
    queries, rest = evaluator_queries(tests, 2)
    assert len(queries) == 931
    recs = [ix.predict(q, 50, 500, 21, mode=mode)[0] for q in queries]
    mrr, hr = mrr_hitrate_at(recs, rest, 20)
    assert hr == pytest.approx(0.6402, abs=0.0025)
    assert mrr == pytest.approx(0.3277, abs=0.004)


def test_faithful_equals_canonical_without_boundary_ties(oracle):
    """closed form == sequential heap procedure when timestamps are unique and k >= m"""
    bad = 0
    for seed in range(4):
        rng = np.random.default_rng(seed)
        items, off, ts = random_index_data(rng, 300, 40, max_len=6, unique_ts=True)
        m_build = int(rng.integers(1, 40))
        ix = oracle.OracleIndex.from_sessions(items, off, ts, m_build, 6, 1.5)
        known = np.unique(items)
        for _ in range(300):
            ev = [int(x) for x in rng.choice(known, size=int(rng.integers(1, 8)))]
            for k, m in [(1000, m_build), (1000, max(1, m_build // 2)), (1000, 3)]:
                a = ix.predict(ev, k, m, 21, mode=0)
                b = ix.predict(ev, k, m, 21, mode=1)
                bad += not same_modulo_ties(a[0], a[1], b[0], b[1])
                na, nb = ix.find_neighbors(ev, k, m, mode=0), ix.find_neighbors(ev, k, m, mode=1)
                bad += sorted(map(int, na[0])) != sorted(map(int, nb[0]))
    assert bad == 0


def test_neighbor_similarity_multiset_is_mode_independent(toy):
    """with k < m the reference's boundary heuristic is order dependent, but the multiset of the k similarities is not"""
    ix, tests = toy
    queries, _ = evaluator_queries(tests, 4)
    for q in queries[:200]:
        a = ix.find_neighbors(q, 50, 500, mode=0)[1]
        b = ix.find_neighbors(q, 50, 500, mode=1)[1]
        assert np.allclose(np.sort(a), np.sort(b), rtol=1e-12)


def test_edge_cases(oracle):
    rng = np.random.default_rng(3)
    items, off, ts = random_index_data(rng, 200, 30, max_len=5)
    ix = oracle.OracleIndex.from_sessions(items, off, ts, 20, 5, 1.0)
    known = np.unique(items)
    for mode in (0, 1):
        assert len(ix.predict([10 ** 15], 10, 20, 5, mode=mode)[0]) == 0      # unknown item → no neighbours
        assert len(ix.predict([], 10, 20, 5, mode=mode)[0]) == 0              # documented deviation from the panic
        it = int(known[0])
        a = ix.predict([it, it, it], 100, 20, 50, mode=mode)                 # duplicates: only the last occurrence counts
        assert it not in set(map(int, a[0]))                                  # current item removed (mod.rs:157-160)
    with pytest.raises(KeyError):
        ix.idf(10 ** 15)                                                      # reference panics (vmis_index.rs:322)
    assert ix.find_attributes(10 ** 15) is None
    assert ix.find_attributes(int(known[0])) == {"is_for_sale": True, "is_adult": False}   # vmis_index.rs:514-517


def test_negative_session_weights(oracle):
    """linear_score is negative for positions 11..99 and 0 from 100 (mod.rs:110-116)"""
    items = np.array([1, 2, 3, 1, 4], dtype=np.uint64)
    off = np.array([0, 3, 5], dtype=np.uint64)
    ts = np.array([10, 20], dtype=np.uint32)
    ix = oracle.OracleIndex.from_sessions(items, off, ts, 10, 10, 1.0)
    ev = [1] + [999] * 14            # item 1 sits at position 15 from the end → weight 1 - 1.5 = -0.5
    for mode in (0, 1):
        ids, sc = ix.predict(ev, 10, 10, 10, mode=mode)
        assert len(ids) == 4 and (sc < 0).all()
    ev = [1] + [999] * 120
    for mode in (0, 1):
        ids, sc = ix.predict(ev, 10, 10, 10, mode=mode)
        assert len(ids) == 4 and (sc == 0).all()
