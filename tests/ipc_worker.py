"""Worker of the 2-process item-sharded check (tests/test_gpu_sharded.py and __graft_entry__.smoke()): one process per
GPU, CUDA IPC handles exchanged through torch.distributed (gloo), results gathered and compared with the CPU oracle.
usage: python tests/ipc_worker.py <rank> <world> <port> [n_items n_sessions n_queries]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import serenade_b200 as sb  # noqa: E402
from oracle import vmis_oracle as vo  # noqa: E402
from serenade_b200.shard import gather_results, shard_queries  # noqa: E402

rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n_items, n_sessions, n_q = (int(x) for x in sys.argv[4:7]) if len(sys.argv) >= 7 else (20000, 150000, 4001)
torch.cuda.set_device(rank)
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
ix = sb.VMISIndex.from_sessions_sharded(items, off, ts, 1502, 34, 2.0, rank, rank, world)
ix.connect_shards(rank, world)                                          # CUDA IPC handles over torch.distributed
q_items, q_off = sb.synth_queries(43, n_items, n_q, 4)
li, lo_, lo, hi = shard_queries(q_items, q_off, rank, world)
ids, sc, cnt = sb.predict_batch(ix, (li, lo_), 288, 1502, 21)
g_ids, g_sc, g_cnt = gather_results(ids, sc, cnt, n_q)
if rank == 0:
    oix = vo.OracleIndex.from_sessions(items, off, ts, 1502, 34, 2.0)
    oids, osc, ocnt, _, _ = oix.predict_batch(q_items, q_off, 288, 1502, 21, mode=1, threads=8)
    assert np.array_equal(g_cnt, ocnt) and np.array_equal(g_ids, oids) and np.array_equal(g_sc, osc)
dist.barrier()
ix.close()
dist.destroy_process_group()
print("rank", rank, "ok")
