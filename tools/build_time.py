#!/usr/bin/env python
"""Wall time of the on-device index build (prepare_hashmap equivalent) at BASELINE config-3 size, and — under
`ncu --metrics gpu__time_duration.sum` — its kernel list.  GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import serenade_b200 as sb
n_items, n_sessions = 1_760_000, 11_556_000
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ix = sb.VMISIndex.synth(42, n_items, n_sessions, 1502, 34, 2.0)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    st = ix.stats()
    print(f"synth + build {1e3 * (t1 - t0):.1f} ms: {st['n_pairs_kept']} interactions, {st['n_items']} items, {st['n_postings']} postings")
    ix.close()
