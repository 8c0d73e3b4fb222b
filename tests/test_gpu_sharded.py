"""Item-sharded postings (BASELINE.json config 5): bit-identical to the single-GPU index.
Single-process variant runs on one GPU (shards cross-attached by pointer); the multi-process variant needs 2 GPUs
and exchanges CUDA IPC handles through torch.distributed."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from util import csr, random_index_data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_shards_on_one_gpu_match_oracle(sb, oracle, n_shards):
    items, off, ts = sb.synth_sessions(42, 3000, 30000)
    shards = [sb.VMISIndex.from_sessions_sharded(items, off, ts, 300, 34, 2.0, 0, s, n_shards) for s in range(n_shards)]
    with pytest.raises(sb.VmisError):                       # peers not attached yet
        sb.predict(shards[0], [int(items[0])], 50, 300, 21)
    for a in range(n_shards):
        for b in range(n_shards):
            if a != b:
                shards[a].attach_shard_ptr(b, shards[b].shard_ptr())
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 300, 34, 2.0)
    q_items, q_off = sb.synth_queries(43, 3000, 600, 4)
    oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, 100, 300, 21, mode=1)
    for sh in shards:
        ids, sc, cnt = sb.predict_batch(sh, (q_items, q_off), 100, 300, 21)
        assert np.array_equal(cnt, ocnt) and np.array_equal(ids, oids) and np.array_equal(sc, osc)
    sess, sim, ncnt = shards[-1].find_neighbors_batch((q_items, q_off), 100, 300)
    for q in range(0, 600, 37):
        ev = q_items[q_off[q]:q_off[q + 1]]
        os_, osim = oix.find_neighbors(ev, 100, 300, mode=1)
        assert np.array_equal(sess[q, :ncnt[q]], os_) and np.array_equal(sim[q, :ncnt[q]], osim)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import serenade_b200 as sb
    from serenade_b200.shard import shard_queries, gather_results
    from oracle import vmis_oracle as vo
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    items, off, ts = sb.synth_sessions(42, 20000, 150000)
    ix = sb.VMISIndex.from_sessions_sharded(items, off, ts, 1502, 34, 2.0, rank, rank, world)
    ix.connect_shards(rank, world)                                          # CUDA IPC handles over torch.distributed
    q_items, q_off = sb.synth_queries(43, 20000, 4001, 4)
    li, lo_, lo, hi = shard_queries(q_items, q_off, rank, world)
    ids, sc, cnt = sb.predict_batch(ix, (li, lo_), 288, 1502, 21)
    g_ids, g_sc, g_cnt = gather_results(ids, sc, cnt, 4001)
    if rank == 0:
        oix = vo.OracleIndex.from_sessions(items, off, ts, 1502, 34, 2.0)
        oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, 288, 1502, 21, mode=1, threads=8)
        assert np.array_equal(g_cnt, ocnt) and np.array_equal(g_ids, oids) and np.array_equal(g_sc, osc)
    dist.barrier(); ix.close(); dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_process_ipc_shards():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "worker.py")
        open(script, "w").write(WORKER.format(root=ROOT, port=port))
        procs = [subprocess.Popen([sys.executable, script, str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                  text=True) for r in range(2)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
