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

On the GPU the reduction lives in the C library (s3dmst_comm_init / s3dmst_aggregate_dense_sharded, csrc/comm.cu:
NCCL bound at run time, no host synchronisation, the left view's reduction overlapping the right view's aggregation);
`comm_init_from_torch` only carries the 128-byte ncclUniqueId from rank 0 to the other ranks.  `minloc_reduce` is the
same reduction written on torch tensors so that it also runs over gloo on CPU (tests/test_parallel_gloo.py).
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
    the default does it with torch ops (the GPU path does not come through here: csrc/comm.cu).
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


def comm_init_from_torch(eng, group=None):
    """Give `eng` (one context per rank) its own NCCL communicator spanning `group`: rank 0 makes the ncclUniqueId with the
    C library, torch.distributed only broadcasts those 128 bytes."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    eng.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)
    return rank, world
