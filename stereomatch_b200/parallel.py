"""Multi-GPU partitioning of the Stereo3DMST path (one process per GPU, torch.distributed for the plumbing).

Two ways the path shards:

* frames  — stereo pairs are independent (every stereo3dmst() call shares nothing): frame i goes to rank
            i mod world, no collective on the data path.
* labels  — one very large pair: every rank builds the (deterministic) forests itself and aggregates only its
            label range [d0, d1); the only exchange is a per-pixel MIN-LOC of (aggregated cost, disparity).
            NCCL has no MINLOC, and the exact-mode cost is fp64, so it is two all-reduces:
              1. all_reduce(MIN) of the best cost                       -> global minimum per pixel
              2. disparity := INT32_MAX where local cost != global min  (s3dmst_minloc_mask on the GPU)
                 all_reduce(MIN) of the disparity                       -> lowest d attaining the minimum
            which is exactly the reference tie rule (strict '<' over ascending d, Stereo3DMST.cpp:177).

`minloc_reduce` is written on torch tensors so the same code runs over NCCL on GPUs and over gloo on CPU
(tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import numpy as np

INT32_MAX = 2**31 - 1


def frames_for_rank(n_frames: int, world: int, rank: int):
    """Frame-sharding: indices of the stereo pairs rank `rank` processes (round robin)."""
    return list(range(rank, n_frames, world))


def label_range(D: int, world: int, rank: int, align: int = 4):
    """Contiguous label shard [d0, d1) of rank `rank`; shard boundaries are multiples of `align` (the pipelined
    aggregation kernel needs d0 % 4 == 0).  Ranks beyond the number of aligned blocks get an empty range."""
    blocks = (D + align - 1) // align
    per, extra = divmod(blocks, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(D, b0 * align), min(D, b1 * align)


def minloc_reduce(best, disp, group=None, mask_fn=None):
    """In-place MIN-LOC all-reduce of (best cost [N] float64, disparity [N] int32) tensors over `group`.

    mask_fn(global_min) must set disp to INT32_MAX wherever the local cost differs from the global minimum;
    the default does it with torch ops (CPU / any device), the GPU path passes Stereo3DMST.minloc_mask.
    Returns (global_min, disp)."""
    import torch
    import torch.distributed as dist

    gmin = best.clone()
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN, group=group)
    if mask_fn is None:
        disp.masked_fill_(best != gmin, INT32_MAX)
    else:
        mask_fn(gmin)
    dist.all_reduce(disp, op=dist.ReduceOp.MIN, group=group)
    return gmin, disp


class _DevArray:
    """Zero-copy view of device memory owned by the C library, for torch.as_tensor()."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def dense_result_tensors(eng, view):
    """torch views (no copy) of the context's dense result: (best cost f64 [N], disparity i32 [N])."""
    import torch

    pb, pd = eng.dense_result_dev(view)
    best = torch.as_tensor(_DevArray(pb, eng.N, "<f8"), device="cuda")
    disp = torch.as_tensor(_DevArray(pd, eng.N, "<i4"), device="cuda")
    return best, disp


def aggregate_dense_label_sharded(eng, view, D, group=None):
    """Label-sharded dense aggregation of one view: this rank's labels on its GPU, then the MIN-LOC reduction.
    On return every rank's context holds the global disparity / best cost (device side)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    d0, d1 = label_range(D, world, rank)
    best, disp = dense_result_tensors(eng, view)
    if d1 > d0:
        eng.aggregate_dense(view, d0, d1, fetch=False)
    else:  # more ranks than label blocks: contribute the identity of MIN-LOC
        best.fill_(float(np.finfo(np.float64).max))
        disp.fill_(INT32_MAX)
    eng.sync()
    torch.cuda.synchronize()
    gmin, _ = minloc_reduce(best, disp, group,
                            mask_fn=lambda gmin: (torch.cuda.synchronize(), eng.minloc_mask(view, gmin.data_ptr()), eng.sync()))
    best.copy_(gmin)  # the context now holds the global minimum next to the global disparity
    torch.cuda.synchronize()
    return best, disp
