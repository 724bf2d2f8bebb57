"""Multi-GPU plumbing for the embarrassingly parallel batch: one process per GPU, problems
sharded across ranks, NO collective on the solve path -- only the final metric reduction
(max of the per-rank device time, sums of counts) goes through torch.distributed."""


def shard_range(total, rank, world):
    """Contiguous [lo, hi) slice of `total` problems owned by `rank` (remainder spread over
    the first ranks, so sizes differ by at most one and every problem has exactly one owner)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def reduce_metrics(times_ms, counts, dist=None, device="cpu"):
    """Job-level numbers from per-rank ones: element-wise MAX of `times_ms` (the job is as slow
    as its slowest rank), element-wise SUM of `counts`.  `dist` is torch.distributed when the
    process group is initialised, else None (single rank)."""
    import torch
    t = torch.tensor([float(v) for v in times_ms], dtype=torch.float64, device=device)
    c = torch.tensor([float(v) for v in counts], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()
