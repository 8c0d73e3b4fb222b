"""Small mixed batch for compute-sanitizer (memcheck / racecheck / synccheck): evolving sessions of 1..40 items (both
phase-0 paths), business rules on and off, a k large enough to leave the granule map (per-neighbour walk + HBM table)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import serenade_b200 as sb
items, off, ts = sb.synth_sessions(42, 3000, 12000)
gix = sb.VMISIndex.from_sessions(items, off, ts, 600, 34, 2.0, device=0)
rng = np.random.default_rng(0)
known = np.unique(items)
qs = [[int(x) for x in rng.choice(known, size=int(rng.integers(1, 41)))] for _ in range(96)]
for k, m, n, biz in ((288, 600, 21, False), (50, 100, 21, True), (2048, 600, 40, False)):
    ids, sc, cnt = sb.predict_batch(gix, qs, k, m, n, biz)
    print(k, m, n, biz, int(cnt.sum()))
sess, sim, c = gix.find_neighbors_batch(qs[:16], 100, 600)
print("neighbors", int(c.sum()))
