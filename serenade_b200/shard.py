"""Host logic of the multi-GPU mode: evolving sessions are independent, so the path shards by QUERY — every rank
holds a replica of the index and predicts a contiguous slice of the batch; there is no data-path collective
(the reference scales the same way, replica pods behind session-affinity routing,
recommend_resource.rs:16-19).  The only communication is the optional gather of the per-query results and the
max-over-ranks reduction of the timings, both through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_bounds(n_q, rank, world):
    """contiguous, balanced slice [lo, hi) of n_q queries for `rank` of `world`"""
    base, rem = divmod(n_q, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_queries(q_items, q_off, rank, world):
    """CSR slice of the batch for this rank → (q_items_local, q_off_local, lo, hi)"""
    q_off = np.asarray(q_off)
    lo, hi = shard_bounds(len(q_off) - 1, rank, world)
    a, b = int(q_off[lo]), int(q_off[hi])
    return np.ascontiguousarray(q_items[a:b]), (q_off[lo:hi + 1] - q_off[lo]).astype(np.uint32), lo, hi


def gather_results(ids, scores, counts, n_q, group=None):
    """all-gather the per-rank result rows into full (n_q, how_many) arrays on every rank"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    how_many = ids.shape[1]
    sizes = [shard_bounds(n_q, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)

    def padded(a, dtype):
        t = torch.zeros((pad,) + a.shape[1:], dtype=dtype)
        t[:a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a).view(np.dtype(str(dtype).replace("torch.", ""))))
        return t

    outs = []
    for a, dt in ((ids, torch.int64), (scores, torch.float64), (counts, torch.int32)):
        mine = padded(a, dt)
        bufs = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(bufs, mine, group=group)
        outs.append(torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)]).numpy())
    return outs[0].view(np.uint64).reshape(n_q, how_many), outs[1].reshape(n_q, how_many), outs[2].view(np.uint32)


def max_over_ranks(seconds, device=None, group=None):
    """a multi-GPU step time is the slowest rank's"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
