#!/usr/bin/env python
"""NVLink traffic of the item-sharded kernel, one process driving two GPUs (so that it can run under ncu, which must
not wrap a multi-rank command): shard 0 on cuda:0, shard 1 on cuda:1, peer access enabled, shards cross-attached by
pointer; the batch runs on cuda:0, whose kernel reads every odd item's posting list from cuda:1's HBM.

  ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k vmis_predict_kernel --launch-skip 2 -c 1 \
      python tools/nvlink_probe.py            # per-launch NVLink bytes of device 0
Prints the algorithmic remote bytes of the same batch (4-byte postings of the remote items visited) for comparison."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import serenade_b200 as sb  # noqa: E402

assert torch.cuda.device_count() >= 2, "needs two GPUs (gpurun --gpus 2)"
n_items, n_sessions, B = 1_760_000, 11_556_000, 1 << 18
K, M, N = 288, 1502, 21
lib = sb.load_library()
shards = [sb.VMISIndex.synth(42, n_items, n_sessions, M, 34, 2.0, d, d, 2) for d in (0, 1)]
rt = C.CDLL("libcudart.so.12")
for a, b in ((0, 1), (1, 0)):
    assert torch.cuda.can_device_access_peer(a, b)
    rt.cudaSetDevice(a)
    rc = rt.cudaDeviceEnablePeerAccess(b, 0)
    assert rc in (0, 704), rc                     # 704 = already enabled
shards[0].attach_shard_ptr(1, shards[1].shard_ptr())
shards[1].attach_shard_ptr(0, shards[0].shard_ptr())
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
qi, qo = sb.synth_queries(43, n_items, B, 4)
di = torch.from_numpy(qi.view(np.int64)).to(dev)
do = torch.from_numpy(qo.view(np.int32)).to(dev)
ids = torch.zeros((B, N), dtype=torch.int64, device=dev)
sc = torch.zeros((B, N), dtype=torch.float64, device=dev)
cnt = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.zeros((B, 4), dtype=torch.int32, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):                              # launch 0: statistics pass; 1-3: plain
    if rep == 3:
        e0.record()
    rc = lib.vmis_predict_batch_device(shards[0].handle, di.data_ptr(), do.data_ptr(), B, K, M, N, 0, ids.data_ptr(), sc.data_ptr(),
                                       cnt.data_ptr(), st.data_ptr() if rep == 0 else None, sp)
    assert rc == 0, lib.vmis_last_error()
e1.record()
torch.cuda.synchronize()
postings = int(st[:, 0].to(torch.int64).sum().item())
ms = e0.elapsed_time(e1)
print(f"{B} evolving sessions on cuda:0, postings of odd items on cuda:1: {ms:.3f} ms = {B / ms / 1e3:.2f} M qps; "
      f"postings visited {postings} → algorithmic remote bytes (half of them, 4 B each) {postings * 2} "
      f"= {postings * 2 / B:.0f} B/query = {postings * 2 / ms / 1e6:.1f} GB/s over NVLink")
