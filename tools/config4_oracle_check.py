#!/usr/bin/env python
"""BASELINE config 4 (synthetic 582 M interactions / 6.5 M items) against the CPU oracle: the host generator produces the
same sessions the device generator does (tests/test_gpu_build.py), the oracle builds its index from them on the host
(minutes, ~30 GB), the B200 builds its own on the device, and the rows of a query sample must agree bit for bit.
One-off evidence run (GPU box): python tools/config4_oracle_check.py [n_queries]  → one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import serenade_b200 as sb  # noqa: E402
from oracle import vmis_oracle as vo  # noqa: E402

K, M, N, MAX_LEN, IDF_W = 288, 1502, 21, 34, 2.0
n_items, n_sessions = (6_500_000, 112_100_000) if not os.environ.get("VMIS_CHECK_SMALL") else (1_760_000, 11_556_000)
n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
t0 = time.time()
items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
t1 = time.time()
print(f"host sessions: {len(items)} interactions in {t1 - t0:.1f}s", file=sys.stderr, flush=True)
oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
t2 = time.time()
print(f"oracle index in {t2 - t1:.1f}s", file=sys.stderr, flush=True)
interactions = int(len(items))
del items, off, ts
gix = sb.VMISIndex.synth(42, n_items, n_sessions, M, MAX_LEN, IDF_W, 0, 0, 1)
t3 = time.time()
q_items, q_off = sb.synth_queries(43, n_items, n_q, 4)
ids, sc, cnt = sb.predict_batch(gix, (q_items, q_off), K, M, N)
oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, K, M, N, False, mode=1, threads=os.cpu_count() or 1)
same_cnt = bool(np.array_equal(cnt, ocnt))
col = np.arange(N)[None, :]
valid = col < cnt[:, None]
same_ids = bool(np.array_equal(np.where(valid, ids, 0), np.where(valid, oids, 0)))
same_sc = bool(np.array_equal(np.where(valid, sc, 0.0), np.where(valid, osc, 0.0)))
print(json.dumps({"workload": "synthetic-582M-6.5M" if n_items == 6_500_000 else "synthetic-60M-1.76M", "interactions": interactions,
                  "items": n_items, "sessions": n_sessions, "k": K, "m": M, "how_many": N, "queries": n_q,
                  "recommendations": int(cnt.sum()), "counts_equal": same_cnt, "ids_bit_exact": same_ids,
                  "scores_bit_exact": same_sc, "bit_exact_vs_canonical_oracle": same_cnt and same_ids and same_sc,
                  "oracle_index_build_s": round(t2 - t1, 1), "device_index_build_s": round(t3 - t2, 2)}))
