#!/usr/bin/env python
"""Per-query work of the bench workload (postings visited, neighbours, neighbour items) and kernel time per
evolving-session length class.  GPU box only: python tools/workload_stats.py [batch]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import serenade_b200 as sb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
K, M, N = 288, 1502, 21
lib = sb.load_library()
gix = sb.VMISIndex.synth(42, 1_760_000, 11_556_000, M, 34, 2.0, 0, 0, 1)
qi, qo = sb.synth_queries(43, 1_760_000, B, 4)
dev = torch.device("cuda", 0)


def run(qi, qo, reps=5):
    n = len(qo) - 1
    di = torch.from_numpy(qi.view(np.int64)).to(dev)
    do = torch.from_numpy(qo.view(np.int32)).to(dev)
    ids = torch.zeros((n, N), dtype=torch.int64, device=dev)
    sc = torch.zeros((n, N), dtype=torch.float64, device=dev)
    cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.zeros((n, 4), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream()
    sp = C.c_void_p(s.cuda_stream)
    lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), n, K, M, N, 0, ids.data_ptr(), sc.data_ptr(),
                                  cnt.data_ptr(), st.data_ptr(), sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), n, K, M, N, 0, ids.data_ptr(),
                                      sc.data_ptr(), cnt.data_ptr(), None, sp)
    e1.record()
    torch.cuda.synchronize()
    if os.environ.get("VMIS_CLOCKS"):     # tuning build -DVMIS_PHASE_CLOCKS: the stats pass leaves cycle counts in the ids
        lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), n, K, M, N, 0, ids.data_ptr(), sc.data_ptr(),
                                      cnt.data_ptr(), st.data_ptr(), sp)
        torch.cuda.synchronize()
        return ids.cpu().numpy()[:, :14], e0.elapsed_time(e1) / reps
    return st.cpu().numpy(), e0.elapsed_time(e1) / reps


st, ms = run(qi, qo)
L = np.diff(qo.astype(np.int64))
print(f"all: {B} queries {ms:.3f} ms  {B / ms / 1e3:.2f} M qps")
names = ["postings_visited", "n_neighbors", "neighbor_items", "n_out"]
if os.environ.get("VMIS_CLOCKS"):
    names = ["cyc phase 0", "cyc phase 1+1b", "cyc phase 2a", "cyc 2b first gather", "cyc 2b first round", "cyc 2b other rounds",
             "cyc 2b barrier", "cyc 3 score+top4", "cyc 3 barrier", "cyc 3 sort+push", "cyc 3 barrier", "cyc 3 tail+sync", "-",
             "top-n queue entries"]
for j, nm in enumerate(names):
    v = st[:, j]
    print(f"  {nm:18s} mean {v.mean():9.1f}  p50 {np.percentile(v, 50):8.0f}  p90 {np.percentile(v, 90):8.0f}  p99 {np.percentile(v, 99):8.0f}  max {v.max()}")
for l in range(1, 5):
    sel = np.nonzero(L == l)[0]
    items = np.concatenate([qi[qo[q]:qo[q + 1]] for q in sel])
    off = np.zeros(len(sel) + 1, dtype=np.uint32)
    off[1:] = np.cumsum(L[sel])
    s2, ms2 = run(np.ascontiguousarray(items), off)
    print(f"L={l}: {len(sel)} queries {ms2:.3f} ms  {len(sel) / ms2 / 1e3:.2f} M qps  us/query/SM-slot {ms2 * 1e3 * 740 / len(sel):.1f}"
          f"  postings {s2[:, 0].mean():.0f} nn {s2[:, 1].mean():.0f} items {s2[:, 2].mean():.0f}"
          + (" | " + " ".join(f"{s2[:, j].mean():.0f}" for j in range(3, s2.shape[1])) if s2.shape[1] > 4 else ""))
