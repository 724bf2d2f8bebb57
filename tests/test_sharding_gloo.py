"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py (shard ownership, max-over-ranks
time, summed counts).  The solve path itself has no collective to test."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_partitions_exactly():
    from fsae_mpc_b200.sharding import shard_range
    for total in (0, 1, 7, 65536, 262144, 1000003):
        for world in (1, 2, 4, 8):
            seen = 0
            prev = 0
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                assert lo == prev and hi >= lo
                prev = hi
                seen += hi - lo
            assert prev == total and seen == total
            sizes = [shard_range(total, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1


def test_c_abi_shard_range_equals_python_rule():
    """fsae_shard_range (the split fsae_ltvmpc_host_pool applies inside ONE process) is the same partition as
    sharding.shard_range (the split of the one-process-per-GPU path).  Loads the library; no GPU call."""
    import ctypes as C
    from fsae_mpc_b200 import _lib
    from fsae_mpc_b200.sharding import shard_range
    lib = _lib.load()
    lo, hi = C.c_int64(), C.c_int64()
    for total in (0, 1, 7, 65536, 262144, 1000003):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                lib.fsae_shard_range(total, r, world, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == shard_range(total, r, world)


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    from fsae_mpc_b200.sharding import shard_range, reduce_metrics
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(1001, rank, world)
    # per-rank "device time" and counts as bench.py produces them
    times, counts = reduce_metrics([100.0 + 10 * rank, 120.0 - 5 * rank], [hi - lo, 3 * (rank + 1)], dist)
    dist.barrier()
    out.put((rank, lo, hi, times, counts))
    dist.destroy_process_group()


def test_reduce_metrics_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, c0), (r1, lo1, hi1, t1, c1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    assert t0 == t1 == [110.0, 120.0]          # max over ranks
    assert c0 == c1 == [1001.0, 9.0]           # sums
