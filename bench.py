#!/usr/bin/env python
"""bench.py — throughput of the Stereo3DMST hot path (dense-label pipeline) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores (oracle port)

Workload = the configuration BASELINE.json's metric is quoted on, configs[3] ("C4"): a batch of 64 synthetic 1920x1080
pairs (256 disparities) plus the 5 bundled FLIR pairs, frame-sharded over the GPUs of one box.
One step = one full pass of the hot path over one batch of --batch (8) synthetic C4 frames per GPU: median + edge
weights + FH forest + min-size merge + BFS re-indexing (both views), truncated colour+gradient cost volume (both views),
two-pass tree-filter aggregation + WTA (both views), left-right check + scan-line fill.  Frames are independent, so N
GPUs process N x 8 different pairs per step with no collective ("weak" scaling; at N = 8 a step is exactly the 64
synthetic pairs of C4).  Metric: Mpix*disparities/s = N*batch*W*H*D / t_step.  The five FLIR pairs (2048x1536, D = 100,
rectified fixtures under tests/golden/) are reported as the `c1_flir` object; C2 (1280x720, D = 128) as `c2`.
At N >= 2 the label-sharded path of config C5 (one pair, label range split over the ranks, NCCL MIN-LOC through the
library's own communicator) runs after the headline on the same ranks: a small pair checked bit for bit against the
oracle on every rank, then the C5 shape (3840x2160, D = 512) with the size-independent check "reduced result == tie-rule
merge of the ranks' partial results"; a mismatch exits non-zero.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, D = 1920, 1080, 256
WORKLOAD = ("C4 frames: synthetic random-dot slanted-plane pairs 1920x1080, D=256, dense 3DMST pipeline "
            "(forest+cost+tree-filter+WTA+LR/fill, both views)")
METRIC = "Mpix*disparities/s"
ALG_BYTES_PER_PXLABEL = 12.0  # read cost 4 + write A_up 4 + read A_up 4 (fp32 model; exact mode moves 20)
ALG_BYTES_COST_BUILD = 4.0    # the cost-volume build writes 4 B per pixel*label per view
FUSE = True                   # --no-fuse: build the cost volume first (params.fuse_cost = -1)


def alg_bytes_per_pxlabel():
    """Algorithmic bytes of the dominant kernel.  With the matching cost computed inside k_agg_flow (the default) that
    kernel does the work of the cost-volume build and of the aggregation: 4 + 12 B per pixel*label of SURVEY §8d."""
    return ALG_BYTES_PER_PXLABEL + (ALG_BYTES_COST_BUILD if FUSE else 0.0)


def eng_kw():
    return {} if FUSE else {"fuse_cost": -1}
FLIR_TAGS = ("000020", "000040", "000060", "000061", "000080")
FLIR_D = 100                   # src/stereo_Yin.cpp:207
CPU_LABEL_PARTS = 4            # CPU arm: label shards per view of a frame (each shard rebuilds its view's forest, like a rank of the GPU's C5 path)


_JSON_OUT = None


def guard_stdout():
    """The ONE JSON line is the only thing that may reach stdout: libraries loaded into this process (NCCL prints its
    version banner there at WARN level) get stderr instead."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(s):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(s + "\n")
    out.flush()


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def read_traffic():
    """DRAM bytes per pixel*label of the batched aggregation launch, from this round's committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    key = "k_agg_flow_fused_c4" if FUSE else "k_agg_flow_c4"
    try:
        tj = json.load(open(p))
        return float(tj[key + "_bytes_per_pxlabel"]), tj.get(key + "_source", "profiles/traffic.json")
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU legs
_FRAME_CACHE = {}


def _cpu_shard(args):
    """One shard of one frame on one host core: ONE view's forest, that view's cost volume for labels [d0, d1), tree
    filter + WTA over them.  (size = (w, h, D): the C4 frame, or a small crop for the untimed warm-up steps.)"""
    seed, view, part, parts, size = args
    from oracle.pyoracle import Oracle
    from stereomatch_b200 import parallel, synth
    w, h, dd = size
    O = Oracle(fast=True)
    key = (seed, size)
    if key not in _FRAME_CACHE:        # synthetic input generation is not part of the path: once per worker (_cpu_init)
        _FRAME_CACHE[key] = synth.make_pair(w, h, dd, seed=seed)[:2]
    L, R = _FRAME_CACHE[key]
    d0, d1 = parallel.label_range(dd, parts, part)
    vol = O.cost_adgrad_range(L, R, d0, d1, views=(view,))[view]
    F = O.forest(R if view else L)
    disp, best, _ = O.aggregate_dense(F, vol)
    return disp + np.int32(d0), best


def _cpu_init(seeds, sizes):
    """Pool initialiser: every worker generates every frame of the step once (synthetic input generation stays outside the
    timed region)."""
    from stereomatch_b200 import synth
    for size in sizes:
        for seed in seeds:
            _FRAME_CACHE[(seed, size)] = synth.make_pair(size[0], size[1], size[2], seed=seed)[:2]


def _merge_and_check(shards, size):
    """MIN-LOC merge of a frame's label shards per view (lowest d wins ties) + left-right check with fill: the rest of the
    step.  shards[view] = that view's shards in ascending label order."""
    from oracle.pyoracle import Oracle
    O = Oracle(fast=True)
    w, h, dd = size
    views = []
    for v in (0, 1):
        disp, best = shards[v][0]
        disp, best = disp.copy(), best.copy()
        for d2, b2 in shards[v][1:]:
            take = b2 < best            # ascending label ranges: strict '<' keeps the lowest d on ties
            disp[take] = d2[take]; best[take] = b2[take]
        views.append(disp.astype(np.float32))
    O.lr_check(views[0], views[1], w, h, dd, True)
    return views


def cpu_baseline_serial():
    """The whole pipeline on ONE full C4 frame, serial, one core (the configuration the reference binary shipped in)."""
    from stereomatch_b200 import synth
    size = (W, H, D)
    _cpu_shard((synth.BASE_SEED + 10, 0, 0, 1, (64, 48, 16)))      # imports, library load
    t0 = time.perf_counter()
    sh = [[_cpu_shard((synth.BASE_SEED + 10, v, 0, 1, size))] for v in (0, 1)]
    t_gen = 0.0
    _merge_and_check(sh, size)
    dt = time.perf_counter() - t0 - t_gen
    return {"value": W * H * D / dt / 1e6, "unit": METRIC, "cores": 1, "kind": "port",
            "sample": f"one whole C4 frame ({W}x{H}, all {D} labels), full dense pipeline both views, oracle -O3 serial, {dt:.1f} s "
                      "(includes ~0.5 s of synthetic input generation)"}


def gpu_config(world, B):
    """The `config` object of the JSON line — the same for both arms (the reference arm runs a bounded sample of it)."""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "parallelism": f"frame-sharded x{world}, no collective",
            "batching": "frames of a step run on their own contexts/streams; one tree-aggregation launch covers the whole batch",
            "l2": f"working set per step ({(8.5 if FUSE else 12.7) * B:.0f} GB of {'' if FUSE else 'cost + '}running sums) exceeds the 126 MB L2; no explicit flush",
            "mode": "exact (fp64, reference association order)" + ("; matching cost computed inside the aggregation kernel (no cost volume)" if FUSE else "")}


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port; the reference itself needs OpenCV/Boost/mc-cnn and cannot run
    on the GPU box) on all host cores, on the SAME full-size C4 frames.  Bounded sample per step: cores // 8 whole frames,
    each as 8 shards (view x label quarter) on 8 cores — a shard builds its view's forest, its quarter of that view's cost
    volume, and aggregates it; the parent merges the quarters by MIN-LOC and runs the left-right check.  (The port is
    memory bound: on the 8-core build box 8 busy cores deliver ~2.5x one core, so coarser shards would only lengthen the
    steps.)  The untimed warm-up steps run the same code on 480x270 crops (a CPU has nothing to warm up beyond imports)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from stereomatch_b200 import synth
    cores = os.cpu_count() or 1
    P = CPU_LABEL_PARTS
    frames = max(1, cores // (2 * P))
    size, small = (W, H, D), (480, 270, 64)
    ctx = mp.get_context("spawn")

    def step(pool, sz):
        jobs = [(synth.BASE_SEED + 10 + f, v, part, P, sz) for f in range(frames) for v in (0, 1) for part in range(P)]
        res = pool.map(_cpu_shard, jobs, chunksize=1)
        for f in range(frames):
            _merge_and_check([[res[2 * P * f + P * v + part] for part in range(P)] for v in (0, 1)], sz)

    seeds = [synth.BASE_SEED + 10 + f for f in range(frames)]
    with ctx.Pool(min(cores, 2 * P * frames), initializer=_cpu_init, initargs=(seeds, (size, small))) as pool:
        for _ in range(max(1, args.warmup)):
            step(pool, small)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(pool, size)
        dt = time.perf_counter() - t0
    value = frames * W * H * D * args.steps / dt / 1e6
    busy = min(cores, 2 * P * frames)
    sample = (f"{frames} whole C4 frame(s) per step ({W}x{H}, all {D} labels, full dense pipeline both views), each frame as {2 * P} shards "
              f"(view x label quarter: forest of the view + its cost labels + tree filter/WTA) on {2 * P} cores, MIN-LOC merge + LR check/fill in the "
              f"parent; {busy} of {cores} cores busy; the {args.warmup} warm-up steps run on 480x270 crops")
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": gpu_config(args.gpus, args.batch),
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": busy, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------- GPU leg
def _load_flir():
    import cv2
    g = os.path.join(ROOT, "tests", "golden")
    pairs = []
    for tag in FLIR_TAGS:
        l = cv2.imread(os.path.join(g, f"flir_{tag}_left.jpg")); r = cv2.imread(os.path.join(g, f"flir_{tag}_right.jpg"))
        if l is None or r is None:
            return None
        pairs.append((l, r))
    return pairs


def bench_flir(api, local, peak):
    """The five bundled FLIR pairs (rectified fixtures): each alone on the GPU (latency), then the five as one batch."""
    pairs = _load_flir()
    if not pairs:
        return {"unavailable": "tests/golden/flir_*.jpg missing"}
    fh, fw = pairs[0][0].shape[:2]
    alg = alg_bytes_per_pxlabel() * fw * fh * FLIR_D * 2
    per = []
    eng = api.Stereo3DMST(device=local, **eng_kw())
    for (l, r) in pairs:
        eng.set_images(l, r)
        for _ in range(2):
            eng.run_dense(FLIR_D, fill=True, fetch=False)
        eng.sync()
        st = np.zeros(4); reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.run_dense(FLIR_D, fill=True, fetch=False); eng.sync()
            st += [eng.stage_ms(i) for i in range(4)]
        per.append(((time.perf_counter() - t0) / reps * 1e3, st / reps))
    eng.close()
    ms = float(np.mean([p[0] for p in per])); st = np.mean([p[1] for p in per], axis=0)
    # the five pairs as one batch (one aggregation launch over every pair's trees)
    engs = [api.Stereo3DMST(device=local, fh_ctas=28, **eng_kw()) for _ in pairs]
    for e, (l, r) in zip(engs, pairs):
        e.set_images(l, r)
    for _ in range(2):
        api.run_dense_batch(engs, FLIR_D, fill=True, fetch=False)
    for e in engs:
        e.sync()
    reps, agg = 3, 0.0
    t0 = time.perf_counter()
    for _ in range(reps):
        api.run_dense_batch(engs, FLIR_D, fill=True, fetch=False)
        for e in engs:
            e.sync()
        agg += engs[0].stage_ms(api.T_AGG)
    bms = (time.perf_counter() - t0) / reps * 1e3
    for e in engs:
        e.close()
    # the reference's OWN pipeline on the bundled pair (BASELINE config C1: plane init, 100 x MST_PMS per view, LabelToDisp,
    # LR check = s3dmst_run, what the drop-in stereo3dmst() symbol runs): ~500 s on one CPU core
    eng = api.Stereo3DMST(device=local, cost_scale=1 / 6.0)
    eng.set_images(*pairs[0])
    eng.run(FLIR_D, seed=1, fetch=False); eng.sync()
    t0 = time.perf_counter()
    eng.run(FLIR_D, seed=2)
    ref_mode_ms = (time.perf_counter() - t0) * 1e3
    eng.close()
    ach1 = alg / (st[2] * 1e-3) / 1e9
    achb = alg * len(pairs) / (agg / reps * 1e-3) / 1e9
    return {"workload": f"the 5 bundled FLIR pairs (rectified as src/stereo_Yin.cpp:135-147), {fw}x{fh}, D={FLIR_D} (stereo_Yin.cpp:207)",
            "ms_per_pair": ms, "ms_per_pair_each": [round(p[0], 3) for p in per],
            "stage_ms": {k: float(st[i]) for i, k in enumerate(("forest", "cost", "aggregate", "post"))},
            "Mpix_disp_per_s": fw * fh * FLIR_D / ms / 1e3,
            "roofline": {"bound": "hbm", "kernel": "k_agg_flow (cluster walk for the giant trees)", "achieved": ach1, "peak": peak, "unit": "GB/s",
                         "frac": ach1 / peak, "launch_ms": float(st[2]), "algorithmic_bytes_per_launch": alg},
            "batch_of_5": {"ms_per_pair": bms / len(pairs), "aggregate_ms": agg / reps, "roofline_frac": achb / peak},
            "reference_mode": {"ms_per_pair": ref_mode_ms, "ms_per_round_per_view": ref_mode_ms / 200.0,
                               "what": "s3dmst_run on pair 000020: random plane init, 100 rounds of MST_PMS per view with the library's device-side "
                                       "proposal generator, LabelToDisp, left-right check; host buffers in, both maps out"}}


def bench_c2(api, synth, local, peak, B, fh_ctas, steps):
    """Round 1's headline workload, kept for continuity: C2 (1280x720, D = 128) in batches of 8."""
    w2, h2, d2 = 1280, 720, 128
    engs = [api.Stereo3DMST(device=local, fh_ctas=fh_ctas, **eng_kw()) for _ in range(B)]
    for i, e in enumerate(engs):
        l, r, _ = synth.make_pair(w2, h2, d2, seed=synth.BASE_SEED + i)
        e.set_images(l, r)
    for _ in range(3):
        api.run_dense_batch(engs, d2, fill=True, fetch=False)
    for e in engs:
        e.sync()
    agg = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        api.run_dense_batch(engs, d2, fill=True, fetch=False)
        for e in engs:
            e.sync()
        agg += engs[0].stage_ms(api.T_AGG)
    ms = (time.perf_counter() - t0) / steps * 1e3
    for e in engs:
        e.close()
    ach = alg_bytes_per_pxlabel() * w2 * h2 * d2 * 2 * B / (agg / steps * 1e-3) / 1e9
    return {"workload": f"C2 synthetic pair {w2}x{h2}, D={d2}, batches of {B}", "value": B * w2 * h2 * d2 / ms / 1e3, "unit": METRIC,
            "ms_per_step": ms, "ms_per_frame": ms / B, "aggregate_ms": agg / steps, "roofline_frac": ach / peak}


def bench_label_sharded(api, parallel, synth, dist, torch, local, rank, world, steps=3):
    """Config C5 on the ranks of this job, through the library's own NCCL communicator (s3dmst_comm_init): a small pair
    checked against the oracle on every rank, then the C5 shape with the merge property.  Returns the JSON object."""
    def allmin_flag(ok):
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def maxms(ms):
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = api.Stereo3DMST(device=local, **eng_kw())
    parallel.comm_init_from_torch(eng)
    out = {"ranks": world, "collective": "per-pixel MIN-LOC of (best cost f64, disparity i32), lowest disparity on ties (csrc/comm.cu)"}
    # ---- (a) small pair, bit-exact against the CPU oracle's full-range result, on every rank
    cw, ch, cd = 480, 270, 96
    L, R, _ = synth.make_pair(cw, ch, cd, seed=synth.BASE_SEED + 100)
    eng.set_images(L, R)
    eng.build_forest(0); eng.build_forest(1)
    if not FUSE:
        eng.build_cost_volume(cd)
    eng.aggregate_dense_sharded(cd)   # (default: no cost volume anywhere — the aggregation kernel computes the matching cost)
    eng.sync()
    from oracle.pyoracle import Oracle
    O = Oracle(fast=True)
    lv, rv = O.cost_adgrad(L, R, cd)
    ok = True
    for view, (img, vol) in enumerate(((L, lv), (R, rv))):
        do, bo, _ = O.aggregate_dense(O.forest(img), vol)
        disp, best = eng.get_dense_result(view)
        ok &= bool(np.array_equal(disp, do) and np.array_equal(best.view(np.uint64), bo.view(np.uint64)))
    out["check"] = allmin_flag(ok)
    out["transport"] = {1: "one kernel over peer memory (CUDA IPC mappings of every rank's result buffers, NVLink loads/stores)",
                        0: "two ncclAllReduce(MIN) + mask kernel per view"}.get(eng.comm_transport(), "none")
    out["check_case"] = f"{cw}x{ch} D={cd}: every rank's reduced (disparity, best cost) == the oracle's full-range result, bit for bit"
    # ---- (b) the C5 shape: 3840x2160, D = 512
    w5, h5, d5 = 3840, 2160, 512
    L, R, _ = synth.make_pair(w5, h5, d5, seed=synth.BASE_SEED + 100)
    eng.set_images(L, R)
    d0, d1 = eng.comm_label_range(d5)
    eng.build_forest(0); eng.build_forest(1)
    if not FUSE:
        eng.build_cost_volume(d5)

    def timed_sharded():
        for _ in range(2):
            eng.aggregate_dense_sharded(d5)
        eng.sync()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        mm = 0.0
        for _ in range(steps):
            eng.aggregate_dense_sharded(d5)
            eng.sync()
            mm += eng.comm_minloc_ms()
        dist.barrier(); torch.cuda.synchronize()
        return maxms((time.perf_counter() - t0) * 1e3 / steps), maxms(mm / steps)

    # timing: aggregation of this rank's labels (default: matching cost included, no volume exists) + the reductions, both views
    ms, mlms = timed_sharded()
    reduced = [eng.get_dense_result(view) for view in (0, 1)]
    # merge property: reduced result == tie-rule merge of the ranks' partial results (gathered to every rank over torch's NCCL).
    # The partial results come from s3dmst_aggregate_dense on a cost volume (built on demand here): the other code path.
    merged_ok = True
    partial = []
    for view in (0, 1):
        if d1 > d0:
            eng.aggregate_dense(view, d0, d1, fetch=False)
            pd, pb = eng.get_dense_result(view)
        else:
            pd = np.full(eng.N, 2**31 - 1, np.int32); pb = np.full(eng.N, np.finfo(np.float64).max)
        partial.append((pd, pb))
    for view in (0, 1):
        pd, pb = partial[view]
        tb = torch.from_numpy(pb).cuda(); td = torch.from_numpy(pd).cuda()
        gb = [torch.empty_like(tb) for _ in range(world)]; gd = [torch.empty_like(td) for _ in range(world)]
        dist.all_gather(gb, tb); dist.all_gather(gd, td)
        mb, md = gb[0].clone(), gd[0].clone()
        for r in range(1, world):          # ranks hold ascending label ranges: strict '<' keeps the lowest d on ties
            take = gb[r] < mb
            mb = torch.where(take, gb[r], mb); md = torch.where(take, gd[r], md)
        disp, best = reduced[view]
        merged_ok &= bool(np.array_equal(md.cpu().numpy(), disp) and np.array_equal(mb.cpu().numpy().view(np.uint64), best.view(np.uint64)))
        del gb, gd, tb, td, mb, md
    out["c5_merge_check"] = allmin_flag(merged_ok)
    out.update({"c5_case": f"{w5}x{h5} D={d5}, {d1 - d0} labels on rank 0", "ms_per_pair": ms, "minloc_ms": mlms,
                "bytes": 2 * w5 * h5 * 12, "Mpix_disp_per_s": w5 * h5 * d5 / ms / 1e3,
                "note": "ms_per_pair = " + ("matching cost + " if FUSE else "") + "aggregation of the rank's label range + MIN-LOC, both views, max over ranks"
                        + (" (no cost volume exists on any rank)" if FUSE else " (the full cost volume was built on every rank beforehand, not timed)")
                        + "; minloc_ms = device time of the reductions alone (the left view's overlaps the right view's aggregation); "
                          "bytes = (8 + 4) B x pixels x 2 views all-reduced"})
    if FUSE:  # for the record: the same call once a volume exists (built by the merge check above): aggregation only
        ms_v, _ = timed_sharded()
        out["ms_per_pair_from_prebuilt_volume"] = ms_v
    eng.close()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from stereomatch_b200 import api, parallel, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = max(1, args.batch)
    # one context (and stream) per frame of the batch; in a batch the cooperative forest kernel takes a share of the SMs
    engs = [api.Stereo3DMST(device=local, fh_ctas=(args.fh_ctas if B > 1 else 0), fh_threads=(args.fh_threads if B > 1 else 0), fh_cluster=args.fh_cluster, **eng_kw()) for _ in range(B)]
    frames = [synth.make_pair(W, H, D, seed=synth.BASE_SEED + 10 + rank * B + i) for i in range(B)]   # different frames per rank (C4: seed+10 ...)
    # pinned host staging for the e2e leg
    pin = [(torch.from_numpy(L.copy()).pin_memory(), torch.from_numpy(R.copy()).pin_memory()) for L, R, _ in frames]
    outs = [(torch.empty(W * H, dtype=torch.float32).pin_memory(), torch.empty(W * H, dtype=torch.float32).pin_memory()) for _ in range(B)]
    del frames
    for e, (hl, hr) in zip(engs, pin):
        e.set_images(hl.numpy(), hr.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        api.run_dense_batch(engs, D, fill=True, fetch=False)

    # ---- device-resident throughput ("value"): images already in HBM, results left in HBM
    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sum(e.launch_count() for e in engs)
    stage_tot = np.zeros(4)
    for e in engs:       # the library accumulates its per-stage CUDA-event times over the calls of the timed region
        for st in (api.T_FOREST, api.T_COST, api.T_AGG, api.T_POST):
            e.stage_total_ms(st, reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        step()
        for e in engs:   # (queueing the steps back to back is 2 % slower: the next forests then start inside this step's aggregation tail)
            e.sync()
    ev1.record()
    barrier()
    t_wall = (time.perf_counter() - t_wall) * 1e3
    # aggregation: ONE launch set per step (engine 0's stream); forest/cost/post: mean over the frames' own streams (they overlap)
    agg_total, agg_samples = engs[0].stage_total_ms(api.T_AGG)
    if agg_samples != args.steps:
        raise SystemExit(f"bench.py: {agg_samples} aggregation timings for {args.steps} steps")
    stage_tot[api.T_AGG] = agg_total
    for st in (api.T_FOREST, api.T_COST, api.T_POST):
        stage_tot[st] = float(np.mean([e.stage_total_ms(st)[0] for e in engs]))
    # the frames run on their own streams: the default-stream events bracket host-synchronised steps, so take the larger
    ms_max = maxms(max(ev0.elapsed_time(ev1), t_wall))
    launches = sum(e.launch_count() for e in engs) - launches0
    clocks = sampler.stop() if rank == 0 else None
    agg_ms = stage_tot[api.T_AGG]
    value = world * B * W * H * D * args.steps / (ms_max * 1e-3) / 1e6

    # ---- end to end through the public call with host buffers (H2D + pipeline + D2H inside the timed region)
    # A stream of batches through s3dmst_set_images_async + s3dmst_run_dense_batch_async: every step uploads its frames from
    # pinned host buffers and copies both maps of every frame back into pinned host buffers (two sets, used alternately);
    # the copies of step k drain while step k+1's uploads and forests start, and the last ones are waited for inside the
    # timed region.
    outs2 = [(torch.empty(W * H, dtype=torch.float32).pin_memory(), torch.empty(W * H, dtype=torch.float32).pin_memory()) for _ in range(B)]

    def e2e_step(k):
        for e, (hl, hr) in zip(engs, pin):
            e.set_images(hl.numpy(), hr.numpy(), sync=False)   # pinned buffers: the uploads overlap the other frames' first kernels
        api.run_dense_batch(engs, D, fill=True, out=(outs if k % 2 == 0 else outs2), wait=False)

    for k in range(2):
        e2e_step(k)
    for e in engs:
        e.sync()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    for e in engs:
        e.sync()            # the D2H copies of the last step
    e1.record()
    barrier()
    e2e_ms = maxms(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_value = world * B * W * H * D * args.steps / (e2e_ms * 1e-3) / 1e6
    for e in engs:
        e.close()
    del engs

    # ---- single-frame latency (one C4 pair alone on the GPU, forest kernel on every SM)
    lat = api.Stereo3DMST(device=local, **eng_kw())
    lat.set_images(pin[0][0].numpy(), pin[0][1].numpy())
    for _ in range(3):
        lat.run_dense(D, fill=True, fetch=False)
    lat.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        lat.run_dense(D, fill=True, fetch=False)
    lat.sync()
    single_ms = (time.perf_counter() - t0) * 1e3 / 5
    single_stages = {k: float(lat.stage_ms(i)) for i, k in enumerate(("forest", "cost", "aggregate", "post"))}
    lat.close()
    del pin, outs, outs2

    peak, peak_src = read_peaks()
    extra = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extra["c1_flir"] = bench_flir(api, local, peak)
        extra["c2"] = bench_c2(api, synth, local, peak, 8, args.fh_ctas, 5)
    label_sharded = None
    if world > 1 and not args.no_label_sharded:
        label_sharded = bench_label_sharded(api, parallel, synth, dist, torch, local, rank, world)

    rc = 0
    if rank == 0:
        per_launch_ms = agg_ms / args.steps                # one aggregation launch set per step covers every frame's trees
        pxl = float(W) * H * D * 2 * B                     # pixel*labels of one launch: both views of B frames
        alg_bytes = alg_bytes_per_pxlabel() * pxl
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        tpp, tsrc = read_traffic()
        cpu = cpu_baseline_serial() if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_max / args.steps, "ms_per_frame": ms_max / args.steps / B, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": gpu_config(world, B),
            "stage_ms_per_step": {k: float(stage_tot[i] / args.steps) for i, k in enumerate(("forest", "cost", "aggregate", "post"))},
            "single_frame": {"ms_per_frame": single_ms, "stage_ms": single_stages},
            "roofline": {"bound": "hbm", "kernel": "k_agg_flow", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (tpp * pxl if tpp else None), "traffic_source": tsrc, "peak_source": peak_src,
                         "launch_ms": per_launch_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_per_pxlabel": alg_bytes_per_pxlabel(),
                         "algorithmic_model": ("k_agg_flow computes the matching cost itself: 4 B (cost-volume build) + 12 B (aggregation) per pixel*label of SURVEY 8d"
                                               if FUSE else "12 B per pixel*label (aggregation, SURVEY 8d); the cost volume is built by k_cost_adgrad"),
                         "frac_12B_aggregation_only": ALG_BYTES_PER_PXLABEL * pxl / (per_launch_ms * 1e-3) / 1e9 / peak,
                         "fp64_traffic_model_gbs": (16.0 if FUSE else 20.0) * pxl / (per_launch_ms * 1e-3) / 1e9},
            "e2e": {"value": e2e_value, "unit": METRIC, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 2 * W * H * 3 * B, "d2h_bytes_per_step": 2 * W * H * 4 * B},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(extra)
        if label_sharded is not None:
            line["label_sharded"] = label_sharded
            if not (label_sharded.get("check") and label_sharded.get("c5_merge_check")):
                rc = 3
        if cpu:
            line["cpu_baseline"] = cpu
        emit(json.dumps(line))
    if world > 1:
        t = torch.tensor([rc], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rc = int(t.item())
        dist.destroy_process_group()
    if rc:
        sys.stderr.write("bench.py: label-sharded result differs from the oracle / the merged partial results\n")
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="stereo pairs per step per GPU")
    ap.add_argument("--fh-ctas", type=int, default=36, help="CTAs of the forest kernel per frame when batching")
    ap.add_argument("--fh-threads", type=int, default=1024, help="threads per CTA of the forest kernel when batching")
    ap.add_argument("--fh-cluster", type=int, default=0, help="CTAs of the thread-block cluster a view's forest kernel runs in (0: cooperative grid)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c1_flir / c2 objects")
    ap.add_argument("--no-fuse", action="store_true", help="build the cost volume with k_cost_adgrad first (params.fuse_cost = -1)")
    ap.add_argument("--no-label-sharded", action="store_true", help="skip the C5 leg at N >= 2")
    args = ap.parse_args()
    global FUSE
    FUSE = not args.no_fuse
    guard_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
