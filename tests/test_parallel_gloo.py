"""Host-side logic of the multi-GPU paths on CPU: world_size-2 gloo processes run the same MIN-LOC reduction
code (stereomatch_b200/parallel.py) that runs over NCCL on the GPUs, plus the partitioning helpers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereomatch_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, D, N, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    cost = rng.integers(0, 6, size=(D, N)).astype(np.float64) * 0.125  # few distinct values => many cross-shard ties
    d0, d1 = parallel.label_range(D, world, rank)
    if d1 > d0:
        best = torch.from_numpy(cost[d0:d1].min(axis=0).copy())
        disp = torch.from_numpy((cost[d0:d1].argmin(axis=0) + d0).astype(np.int32))
    else:
        best = torch.full((N,), float(np.finfo(np.float64).max), dtype=torch.float64)
        disp = torch.full((N,), parallel.INT32_MAX, dtype=torch.int32)
    gmin, disp = parallel.minloc_reduce(best, disp)
    ok = np.array_equal(disp.numpy(), cost.argmin(axis=0).astype(np.int32)) and np.array_equal(gmin.numpy(), cost.min(axis=0))
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("D", [512, 10, 3])
def test_minloc_reduce_world2_gloo(D):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), D, 4096, 7, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_label_range_partitions():
    for D in (1, 3, 4, 100, 128, 512, 513):
        for world in (1, 2, 3, 8):
            rs = [parallel.label_range(D, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == D
            for (a0, a1), (b0, b1) in zip(rs, rs[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(r[0] % 4 == 0 for r in rs if r[1] > r[0])
            assert max(r[1] - r[0] for r in rs) - min(r[1] - r[0] for r in rs) <= 4 + 3


def test_frames_for_rank_partitions():
    for n in (0, 1, 5, 64, 69):
        for world in (1, 2, 4, 8):
            allf = sorted(f for r in range(world) for f in parallel.frames_for_rank(n, world, r))
            assert allf == list(range(n))
            sizes = [len(parallel.frames_for_rank(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
