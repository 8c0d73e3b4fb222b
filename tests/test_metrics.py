"""The eight evaluator metrics (metrics/evaluation_reporter.rs) restated in serenade_b200/evaluate.py: pinned against
every known-answer test the reference holds for them, and against the README evaluator line on the toy data using
the CPU oracle's predictions (no GPU needed)."""
import os

import numpy as np
import pytest

from serenade_b200.evaluate import EvaluationReporter, read_test_sessions, read_training_items
from util import evaluator_queries

RECS24 = list(range(1, 25))


def _one(recs, nxt, length=20):
    r = EvaluationReporter(np.array([1, 2, 3], dtype=np.uint64), length)
    r.add(recs, nxt)
    return r.result()


def test_reference_known_answers():
    assert _one(RECS24, [3, 55, 3, 4])["Mrr@20"] == 0.3333333333333333                  # metrics/mrr.rs:53-62
    assert _one(RECS24, [3, 55, 88, 4])["Ndcg@20"] == 0.36121211352040195                # metrics/ndcg.rs:76-87
    assert _one(RECS24, [3, 55, 3, 4])["Precision@20"] == 2.0 / 20                       # metrics/precision.rs:64-75
    assert abs(_one(RECS24, [3, 55, 3, 4])["Recall@20"] - 0.5) < 2.3e-16                 # metrics/recall.rs:65-76
    assert abs(_one([1, 2], [2, 3])["HitRate@20"] - 1.0) < 2.3e-16                       # metrics/hitrate.rs:55-63
    assert abs(_one([1, 2], [2, 3])["F1score@20"] - 0.09090909090909091) < 2.3e-16       # metrics/f1score.rs:51-59
    empty = EvaluationReporter(np.zeros(0, dtype=np.uint64), 20).result()                # divide-by-zero tests
    assert empty["HitRate@20"] == 0.0 and empty["F1score@20"] == 0.0 and empty["Coverage@20"] == 0.0


def test_popularity_and_coverage():
    train = np.array([5, 5, 5, 5, 6, 6, 7, 9], dtype=np.uint64)                          # max frequency 4, 4 distinct
    r = EvaluationReporter(train, 2)
    r.add([5, 6, 7], [1])                  # top-2 = {5, 6}: (4/4 + 2/4) / 2
    r.add([42, 7], [1])                    # unknown item counts in the denominator only: (0 + 1/4) / 2
    res = r.result()
    assert res["Popularity@2"] == pytest.approx((0.75 + 0.125) / 2)
    assert res["Coverage@2"] == pytest.approx(4 / 4)                                     # {5, 6, 42, 7} / 4 training items


def test_readme_evaluator_line_with_oracle_predictions(oracle, toy_dir):
    """README.md:166-172: Mrr 0.3277, Ndcg 0.3553, HitRate 0.6402, Popularity 0.0499, Precision 0.0680, Coverage 0.2765,
    Recall 0.4456, F1score 0.1180 over 931 evaluations (example.toml: m=500, k=50, n=21, max_items=2, idf=1).  The
    tie order of the shipped binary is unpinned, hence the tolerances."""
    train = os.path.join(toy_dir, "train.txt")
    oix = oracle.OracleIndex.new_from_csv(train, 500, 1.0, 15)
    tests = read_test_sessions(os.path.join(toy_dir, "test.txt"))
    queries, rest = evaluator_queries({k: v for k, v in tests.items()}, 2)
    rep = EvaluationReporter(read_training_items(train), 20)
    for ev, nxt in zip(queries, rest):
        ids, _ = oix.predict(ev, 50, 500, 21, mode=1)
        rep.add(ids, nxt)
    res = rep.result()
    assert rep.n == 931
    want = {"Mrr@20": 0.3277, "Ndcg@20": 0.3553, "HitRate@20": 0.6402, "Popularity@20": 0.0499, "Precision@20": 0.0680,
            "Coverage@20": 0.2765, "Recall@20": 0.4456, "F1score@20": 0.1180}
    for name, v in want.items():
        assert res[name] == pytest.approx(v, abs=0.004), (name, res[name], v)
    assert res["HitRate@20"] == pytest.approx(0.6402, abs=0.0006)
