"""Label-range sharding of ONE pair across the ranks of a torchrun job (BASELINE config C5) through the C ABI
(s3dmst_comm_init / s3dmst_aggregate_dense_sharded: the library's own NCCL communicator, csrc/comm.cu).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/label_sharded.py [--W 480 --H 270 --D 96]

Every rank's reduced (disparity, best cost) must equal the CPU oracle's full-range result bit for bit; exits non-zero
otherwise.  (bench.py --gpus N runs the same check plus the full C5 shape after its headline.)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from stereomatch_b200 import api, parallel, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--W", type=int, default=480); ap.add_argument("--H", type=int, default=270); ap.add_argument("--D", type=int, default=96)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, R, _ = synth.make_pair(a.W, a.H, a.D, seed=synth.BASE_SEED + 100)
    eng = api.Stereo3DMST(device=local)
    parallel.comm_init_from_torch(eng)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)      # every rank rebuilds the (deterministic) forests
    eng.build_cost_volume(a.D)
    eng.aggregate_dense_sharded(a.D)
    eng.sync()
    from oracle.pyoracle import Oracle
    O = Oracle(fast=True)
    lv, rv = O.cost_adgrad(L, R, a.D)
    ok = True
    for view, (img, vol) in enumerate(((L, lv), (R, rv))):
        do, bo, _ = O.aggregate_dense(O.forest(img), vol)
        disp, best = eng.get_dense_result(view)
        ok &= bool(np.array_equal(disp, do) and np.array_equal(best.view(np.uint64), bo.view(np.uint64)))
    flags = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"config": f"label-sharded {a.W}x{a.H} D={a.D} over {world} GPU(s), labels {eng.comm_label_range(a.D)} on rank 0",
                          "check": bool(flags.item()), "minloc_ms": eng.comm_minloc_ms()}))
    eng.close()
    dist.destroy_process_group()
    if not bool(flags.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()
