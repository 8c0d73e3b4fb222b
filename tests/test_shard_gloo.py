"""world_size-2 gloo test of the query-sharded multi-GPU host logic (no GPU: the per-rank compute is the oracle)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_the_batch():
    from serenade_b200.shard import shard_bounds, shard_queries
    for n_q in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n_q, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n_q
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    q_off = np.array([0, 2, 3, 7, 8, 10], dtype=np.uint32)
    q_items = np.arange(10, dtype=np.uint64)
    it, off, lo, hi = shard_queries(q_items, q_off, 1, 2)
    assert (lo, hi) == (3, 5) and list(off) == [0, 1, 3] and list(it) == [7, 8, 9]


WORKER = textwrap.dedent("""
    import os, sys, time
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch.distributed as dist
    from serenade_b200.shard import shard_queries, gather_results, max_over_ranks
    from oracle import vmis_oracle as vo
    from util import random_index_data, csr
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    rng = np.random.default_rng(7)
    items, off, ts = random_index_data(rng, 400, 40, max_len=6)
    ix = vo.OracleIndex.from_sessions(items, off, ts, 50, 6, 1.0)          # replica on every rank
    known = np.unique(items)
    queries = [[int(x) for x in rng.choice(known, size=int(rng.integers(1, 5)))] for _ in range(101)]
    q_items, q_off = csr(queries)
    li, lo_, lo, hi = shard_queries(q_items, q_off, rank, 2)
    t0 = time.time()
    ids, sc, cnt, _, _ = ix.predict_batch(li, lo_, 20, 50, 21, mode=1)
    t = max_over_ranks(time.time() - t0 + rank)                            # rank 1 is "slower" by 1 s
    assert t >= 1.0
    g_ids, g_sc, g_cnt = gather_results(ids, sc, cnt, 101)
    f_ids, f_sc, f_cnt, _, _ = ix.predict_batch(q_items, q_off, 20, 50, 21, mode=1)
    assert np.array_equal(g_ids, f_ids) and np.array_equal(g_sc, f_sc) and np.array_equal(g_cnt, f_cnt)
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_gloo_sharding(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
