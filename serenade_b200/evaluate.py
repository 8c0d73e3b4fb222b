"""Batched replay evaluator: the reference's `evaluator` binary (src/bin/evaluator.rs:46-76) and HPO objective
(src/objective.rs:20-47) call `predict` once per prefix of every test session; here all prefixes of the whole
test set go through ONE `predict_batch` call and the ranking metrics are computed on the returned id matrix.

Metrics follow the reference definitions: Mrr (metrics/mrr.rs:24-33) and HitRate (metrics/hitrate.rs:25-33) score
the first of the remaining items against the top-`length` recommendations."""
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
