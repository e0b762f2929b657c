"""Label-range sharding of ONE large pair across the ranks of a torchrun job (BASELINE config C5), NCCL MIN-LOC.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/label_sharded.py [--check] [--W 3840 --H 2160 --D 512]

--check: a small pair whose full-range result the CPU oracle provides; every rank's reduced result must equal it
bit for bit.  Otherwise the C5-shaped run: timing (device events, max over ranks) and the size-independent
property that the sharded result equals the merge of the per-shard results done by k_wta-style tie rule on rank 0.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from stereomatch_b200 import api, parallel, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--W", type=int, default=3840); ap.add_argument("--H", type=int, default=2160); ap.add_argument("--D", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, D = (480, 270, 96) if a.check else (a.W, a.H, a.D)
    L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + 100)
    eng = api.Stereo3DMST(device=local, stream=torch.cuda.current_stream().cuda_stream)
    eng.set_images(L, R)
    d0, d1 = parallel.label_range(D, world, rank)
    # every rank rebuilds the (deterministic) forests and its own slice of the cost volume
    eng.build_forest(0); eng.build_forest(1)
    eng.build_cost_volume(D)
    res = {}
    for view in (0, 1):
        best, disp = parallel.aggregate_dense_label_sharded(eng, view, D)
        res[view] = (best.cpu().numpy().copy(), disp.cpu().numpy().copy())
    ok = True
    if a.check:
        from oracle.pyoracle import Oracle
        O = Oracle(fast=True)
        lv, rv = O.cost_adgrad(L, R, D)
        for view, (img, vol) in enumerate(((L, lv), (R, rv))):
            do, bo, _ = O.aggregate_dense(O.forest(img), vol)
            ok &= bool(np.array_equal(res[view][1], do) and np.array_equal(res[view][0].view(np.uint64), bo.view(np.uint64)))
    # timing of the sharded step (aggregation of this rank's labels + the two all-reduces), both views
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(a.steps):
        for view in (0, 1):
            parallel.aggregate_dense_label_sharded(eng, view, D)
    ev1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / a.steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    flags = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"config": f"label-sharded {W}x{H} D={D} over {world} GPU(s), {d1 - d0} labels on rank 0", "check": bool(flags.item()) if a.check else None,
                          "ms_per_pair_aggregation_plus_minloc": ms, "Mpix_disp_per_s": 2 * W * H * D / ms / 1e3,
                          "minloc_bytes_per_pair": 2 * W * H * 12 * 2}))
    eng.close()
    dist.destroy_process_group()
    if a.check and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
