#!/usr/bin/env python
"""Small-batch makespan: launches of 1024 sessions on the config-2 index, natural order against longest-first order
(evolving-session length descending), back to back so the clocks stay up.  GPU box: python tools/lpt_probe.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import serenade_b200 as sb  # noqa: E402

K, M, N = 288, 1502, 21
lib = sb.load_library()
n_items = 50_000
gix = sb.VMISIndex.synth(42, n_items, 193_000, M, 34, 2.0, 0, 0, 1)
dev = torch.device("cuda", 0)
clk = "clk" in os.environ.get("VMIS_LIB", "")


def reorder(qi, qo, order):
    L = np.diff(qo.astype(np.int64))
    items = np.concatenate([qi[qo[q]:qo[q + 1]] for q in order])
    off = np.zeros(len(order) + 1, dtype=np.uint32)
    off[1:] = np.cumsum(L[order])
    return np.ascontiguousarray(items), off


for B in (256, 740, 1024, 2048, 4096):
    rows = []
    for mode in ("natural", "longest-first", "shortest-first"):
        sets = []
        for s in range(16):
            qi, qo = sb.synth_queries(100 + s, n_items, B, 4)
            L = np.diff(qo.astype(np.int64))
            if mode == "longest-first":
                qi, qo = reorder(qi, qo, np.argsort(-L, kind="stable"))
            elif mode == "shortest-first":
                qi, qo = reorder(qi, qo, np.argsort(L, kind="stable"))
            sets.append((torch.from_numpy(qi.view(np.int64)).to(dev), torch.from_numpy(qo.view(np.int32)).to(dev)))
        ids = torch.zeros((B, N), dtype=torch.int64, device=dev)
        sc = torch.zeros((B, N), dtype=torch.float64, device=dev)
        cnt = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.zeros((B, 4), dtype=torch.int32, device=dev)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def call(i, stats=False):
            di, do = sets[i % len(sets)]
            rc = lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), B, K, M, N, 0, ids.data_ptr(),
                                               sc.data_ptr(), cnt.data_ptr(), st.data_ptr() if stats else None, sp)
            assert rc == 0
        for i in range(64):
            call(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 256
        e0.record()
        for i in range(reps):
            call(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        line = f"B={B:5d} {mode:15s}: {us:7.1f} us per launch  {B / us:6.2f} M qps"
        if clk:
            call(0, True)
            torch.cuda.synchronize()
            c = ids.cpu().numpy()[:, :9].astype(np.float64)
            c = c[(c < 1e7).all(1)].sum(1)
            line += f" | cycles/query mean {c.mean():.0f} p50 {np.percentile(c, 50):.0f} p90 {np.percentile(c, 90):.0f} p99 {np.percentile(c, 99):.0f} max {c.max():.0f}"
        print(line, flush=True)
