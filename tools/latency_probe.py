#!/usr/bin/env python
"""Per-phase cycle counts of the predict kernel as a function of the batch size (1 session alone on the GPU ... a full
launch): where the latency of ONE query goes.  GPU box, tuning build only:
  make -C serenade_b200/csrc clocks; VMIS_LIB=$PWD/build/variants/libvmis_clk.so python tools/latency_probe.py"""
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
small = os.environ.get("VMIS_PROBE_SMALL")
if small:
    gix = sb.VMISIndex.synth(42, 50_000, 193_000, M, 34, 2.0, 0, 0, 1)
    n_items = 50_000
else:
    gix = sb.VMISIndex.synth(42, 1_760_000, 11_556_000, M, 34, 2.0, 0, 0, 1)
    n_items = 1_760_000
dev = torch.device("cuda", 0)
names = ["phase 0", "phase 1+1b", "phase 2a", "2b first gather", "2b first round", "2b other rounds", "2b barrier", "3 score+top4", "3 barrier", "3 sort+push",
         "3 barrier", "3 tail+sync"]
clk = "clk" in os.environ.get("VMIS_LIB", "")
for B in (1, 8, 148, 740, 1024):
    qi, qo = sb.synth_queries(43 + B, n_items, B, 4)
    di = torch.from_numpy(qi.view(np.int64)).to(dev)
    do = torch.from_numpy(qo.view(np.int32)).to(dev)
    ids = torch.zeros((B, N), dtype=torch.int64, device=dev)
    sc = torch.zeros((B, N), dtype=torch.float64, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.zeros((B, 4), dtype=torch.int32, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call(stats):
        rc = lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), B, K, M, N, 0, ids.data_ptr(), sc.data_ptr(),
                                           cnt.data_ptr(), st.data_ptr() if stats else None, sp)
        assert rc == 0
    for _ in range(3):
        call(False)
    torch.cuda.synchronize()
    reps = 20 if B <= 4096 else 3
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(False); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    line = f"B={B:7d}: kernel p50 {np.median(ts):9.1f} us  min {min(ts):9.1f} us"
    if clk:
        call(True)
        torch.cuda.synchronize()
        c = ids.cpu().numpy()[:, :len(names)].astype(np.float64)
        c = c[(c < 1e7).all(1) & (c[:, 4] > 0)]           # rows of queries whose phase 0 ran ahead carry foreign clocks
        line += "  | cycles: " + "  ".join(f"{n} {c[:, j].mean():.0f}" for j, n in enumerate(names)) + f"  | sum {c.sum(1).mean():.0f}"
    print(line, flush=True)
