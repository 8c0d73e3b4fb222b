"""Shared helpers for the tests: evaluator-style query construction and comparators."""
import numpy as np


def read_test_sessions(path):
    """io.rs:40-59 read_test_data_evolving: sessions grouped by id, items ordered by time.
    (The reference uses an unstable sort on time; ties are kept in file order here.)"""
    sess = {}
    with open(path) as f:
        next(f)
        for line in f:
            p = line.split()
            if len(p) < 3:
                continue
            sess.setdefault(int(p[0]), []).append((int(p[1]), int(round(float(p[2])))))
    out = {}
    for sid, ev in sess.items():
        ev = sorted(ev, key=lambda x: x[1])
        out[sid] = [i for i, _ in ev]
    return out


def evaluator_queries(test_sessions, max_items_in_session):
    """evaluator.rs:46-57: for session_state in 1..len, the last max_items of the prefix;
    returns (queries, remaining_items)."""
    qs, rest = [], []
    for sid in sorted(test_sessions):
        items = test_sessions[sid]
        for state in range(1, len(items)):
            start = state - max_items_in_session if state > max_items_in_session else 0
            qs.append(items[start:state])
            rest.append(items[state:])
    return qs, rest


def mrr_hitrate_at(recs, rest, n=20):
    """metrics/mrr.rs:24-33 and metrics/hitrate.rs:25-33 (next item = first of the remaining items)."""
    rr, hit = 0.0, 0
    for r, nxt in zip(recs, rest):
        r = list(r)[:n]
        if nxt[0] in r:
            rr += 1.0 / (r.index(nxt[0]) + 1)
            hit += 1
    return rr / len(recs), hit / len(recs)


def same_modulo_ties(ids_a, sc_a, ids_b, sc_b, rtol=1e-12):
    """ids equal as sequences up to permutation inside exact-score tie classes; the last tie class may be
    cut differently by the how_many boundary, so it is compared as 'subset of the same score'."""
    if len(ids_a) != len(ids_b):
        return False
    if not np.allclose(sc_a, sc_b, rtol=rtol, atol=0):
        return False
    i, n = 0, len(ids_a)
    while i < n:
        j = i
        while j < n and abs(sc_a[j] - sc_a[i]) <= rtol * max(abs(sc_a[i]), 1e-300):
            j += 1
        if j < n and set(map(int, ids_a[i:j])) != set(map(int, ids_b[i:j])):
            return False
        i = j
    return True


def csr(sessions):
    lens = np.array([len(s) for s in sessions], dtype=np.int64)
    off = np.zeros(len(sessions) + 1, dtype=np.uint32)
    off[1:] = np.cumsum(lens)
    items = np.array([i for s in sessions for i in s], dtype=np.uint64)
    return items, off


def random_index_data(rng, n_sessions, n_items, max_len=8, unique_ts=True, id_scale=1):
    """Random training sessions: (items u64, off u64, ts u32)."""
    lens = rng.integers(1, max_len + 1, size=n_sessions)
    items, off = [], [0]
    for ln in lens:
        ln = min(int(ln), n_items)
        # skewed popularity so that posting lists overlap
        p = 1.0 / np.arange(1, n_items + 1)
        p /= p.sum()
        s = rng.choice(n_items, size=ln, replace=False, p=p)
        items.extend(sorted(int(x) * id_scale + 7 for x in s))
        off.append(len(items))
    if unique_ts:
        ts = rng.permutation(n_sessions).astype(np.uint32) + 1000
    else:
        ts = rng.integers(1000, 1000 + max(2, n_sessions // 4), size=n_sessions).astype(np.uint32)
    return np.array(items, dtype=np.uint64), np.array(off, dtype=np.uint64), ts
