"""Item-sharded postings (BASELINE.json config 5): bit-identical to the single-GPU index.
Single-process variant runs on one GPU (shards cross-attached by pointer); the multi-process variant needs 2 GPUs
and exchanges CUDA IPC handles through torch.distributed."""
import os
import socket
import subprocess
import sys

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


def test_two_process_ipc_shards():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = os.path.join(ROOT, "tests", "ipc_worker.py")
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
