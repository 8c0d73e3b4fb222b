"""Batched replay evaluator: the reference's `evaluator` binary (src/bin/evaluator.rs:46-76) and HPO objective
(src/objective.rs:20-47) call `predict` once per prefix of every test session; here all prefixes of the whole
test set go through ONE `predict_batch` call and the ranking metrics are computed on the returned id matrix.

Metrics follow the reference definitions: Mrr (metrics/mrr.rs:24-33) and HitRate (metrics/hitrate.rs:25-33) score
the first of the remaining items against the top-`length` recommendations; `EvaluationReporter` restates all eight
metrics of metrics/evaluation_reporter.rs (the line the reference's evaluator prints, README.md:170-171)."""
import math

import numpy as np

from .vmis import predict_batch


def read_test_sessions(path):
    """io.rs:40-59 `read_test_data_evolving`: rows grouped by session id, items ordered by time."""
    sess = {}
    with open(path) as f:
        next(f)                                        # header
        for line in f:
            p = line.split()
            if len(p) >= 3:
                sess.setdefault(int(p[0]), []).append((int(p[1]), int(round(float(p[2])))))
    return {sid: [i for i, _ in sorted(ev, key=lambda x: x[1])] for sid, ev in sess.items()}


def evolving_queries(test_sessions, max_items_in_session):
    """evaluator.rs:46-57: for session_state in 1..len the last `max_items_in_session` items of the prefix.
    Returns CSR (q_items u64, q_off u32) and the next item of every prefix."""
    q_items, q_off, nxt = [], [0], []
    for sid in sorted(test_sessions):
        items = test_sessions[sid]
        for state in range(1, len(items)):
            start = state - max_items_in_session if state > max_items_in_session else 0
            q_items.extend(items[start:state])
            q_off.append(len(q_items))
            nxt.append(items[state])
    return (np.asarray(q_items, dtype=np.uint64), np.asarray(q_off, dtype=np.uint32)), np.asarray(nxt, dtype=np.uint64)


def evaluate(index, test_sessions, k, m, how_many=21, max_items_in_session=2, length=20, enable_business_logic=False):
    """→ dict(qty_evaluations, mrr, hitrate) at cut-off `length` (the reference reports @20 with how_many = 21)."""
    queries, nxt = evolving_queries(test_sessions, max_items_in_session)
    ids, _, cnt = predict_batch(index, queries, k, m, how_many, enable_business_logic)
    n = len(nxt)
    cols = np.arange(ids.shape[1])[None, :]
    valid = (cols < np.minimum(cnt, length)[:, None])
    hit = (ids == nxt[:, None]) & valid
    has = hit.any(axis=1)
    rank = hit.argmax(axis=1) + 1
    return {"qty_evaluations": int(n), "mrr": float(np.where(has, 1.0 / rank, 0.0).sum() / max(n, 1)),
            "hitrate": float(has.sum() / max(n, 1))}


def read_training_items(path):
    """io.rs:13-30 `read_training_data`: the item column of every training row (duplicates included) — what
    Popularity (metrics/popularity.rs:20-37) and Coverage (metrics/coverage.rs:17-27) are built from."""
    items = []
    with open(path) as f:
        next(f)
        for line in f:
            p = line.split()
            if len(p) >= 3:
                items.append(int(p[1]))
    return np.asarray(items, dtype=np.uint64)


class EvaluationReporter:
    """metrics/evaluation_reporter.rs:12-117 — Mrr, Ndcg, HitRate, Popularity, Precision, Coverage, Recall, F1score
    at cut-off `length`; `add(recommendations, next_items)` once per evaluated prefix (evaluator.rs:66-74)."""

    NAMES = ("Mrr", "Ndcg", "HitRate", "Popularity", "Precision", "Coverage", "Recall", "F1score")

    def __init__(self, training_items, length):
        self.length = length
        ids, freq = np.unique(np.asarray(training_items, dtype=np.uint64), return_counts=True)
        self.popularity_scores = dict(zip(ids.tolist(), freq.tolist()))      # popularity.rs:21-27
        self.max_frequency = int(freq.max()) if len(freq) else 0
        self.unique_training_items = len(ids)                                # coverage.rs:18-21
        self.test_items = set()
        self.n = 0
        self.sums = dict.fromkeys(("mrr", "ndcg", "hitrate", "popularity", "precision", "recall"), 0.0)

    @staticmethod
    def _dcg(top, next_set):                                                 # ndcg.rs:13-26
        r = 0.0
        for index, item in enumerate(top):
            if item in next_set:
                r += 1.0 if index == 0 else 1.0 / math.log2(index + 1.0)
        return r

    def add(self, recommendations, next_items):
        recommendations = [int(x) for x in recommendations]
        next_items = [int(x) for x in next_items]
        top = recommendations[:self.length]
        self.n += 1
        nxt = next_items[0]
        if nxt in top:                                                       # mrr.rs:24-33, hitrate.rs:25-33
            self.sums["mrr"] += 1.0 / (top.index(nxt) + 1.0)
            self.sums["hitrate"] += 1.0
        next_set = set(next_items)
        dcg_max = self._dcg(next_items[:self.length], next_set)              # ndcg.rs:44-58
        self.sums["ndcg"] += self._dcg(top, next_set) / dcg_max
        top_set = set(top)
        inter = len(top_set & next_set)
        self.sums["precision"] += inter / self.length                        # precision.rs:31-40
        self.sums["recall"] += inter / len(next_items)                       # recall.rs:33-45
        if top_set:                                                          # popularity.rs:41-58
            self.sums["popularity"] += sum(self.popularity_scores.get(i, 0) / self.max_frequency
                                           for i in top_set) / len(top_set)
        self.test_items.update(top)                                          # coverage.rs:31-39

    def result(self):
        """dict keyed like the reference's header line: 'Mrr@20', 'Ndcg@20', ..."""
        n = self.n
        avg = {k: (v / n if n else 0.0) for k, v in self.sums.items()}
        p, r = avg["precision"], avg["recall"]
        f1 = 2.0 * (p * r) / (p + r) if (p + r) > 0 else 0.0                 # f1score.rs:29-39 (NaN → 0)
        cov = len(self.test_items) / self.unique_training_items if self.unique_training_items else 0.0
        vals = (avg["mrr"], avg["ndcg"], avg["hitrate"], avg["popularity"], p, cov, r, f1)
        return {f"{name}@{self.length}": v for name, v in zip(self.NAMES, vals)}


def evaluate_all(index, training_items, test_sessions, k, m, how_many=21, max_items_in_session=2, length=20,
                 enable_business_logic=False):
    """The reference's `evaluator` run (evaluator.rs:40-90) with ONE batched predict: all eight metrics + count."""
    queries, _ = evolving_queries(test_sessions, max_items_in_session)
    ids, _, cnt = predict_batch(index, queries, k, m, how_many, enable_business_logic)
    rep = EvaluationReporter(training_items, length)
    q = 0
    for sid in sorted(test_sessions):
        items = test_sessions[sid]
        for state in range(1, len(items)):
            rep.add(ids[q, :cnt[q]], items[state:])
            q += 1
    out = rep.result()
    out["qty_evaluations"] = rep.n
    return out
