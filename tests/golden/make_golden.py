"""tests/golden/make_golden.py — regenerates tests/golden/ref_small.npz.

Runs the REFERENCE's own code (oracle/_ref/libref3dmst.so = /root/reference/src/Stereo3DMST.cpp
compiled unmodified, see oracle/ref_driver.cpp) on a small seeded input and stores inputs and
outputs, so that the oracle can be checked against reference output on machines where
/root/reference is not mounted.  Run from the repo root in the build container:
    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Ref  # noqa: E402
from stereomatch_b200 import synth  # noqa: E402


def main():
    O, R = Oracle(), Ref()
    W, H, D = 72, 48, 12
    L, Rt, gt = synth.make_pair(W, H, D, seed=11)
    N = W * H
    # cost source for the fixture: the a2' volume scaled into the reference's [0, 0.5] range
    lv_raw, rv_raw = O.cost_adgrad(L, Rt, D)
    lv_raw = (lv_raw * np.float32(1 / 6.0)).astype(np.float32)
    rv_raw = (rv_raw * np.float32(1 / 6.0)).astype(np.float32)
    lv_raw[2, 17] = np.nan  # exercises the NaN scrub at Stereo3DMST.cpp:788
    out = dict(W=W, H=H, D=D, left=L, right=Rt, lv_raw=lv_raw, rv_raw=rv_raw)
    for tag, (c, ms) in dict(a=(5000.0, 200), b=(300.0, 20)).items():
        R.srand(1)
        V = R.view(L, D, c=c, min_size=ms)
        out[f"{tag}_params"] = np.float64([c, ms])
        for k in ("tree_start", "node_pixel", "parent", "child_count", "weight", "weight2", "adj_ptr", "adj", "abc"):
            out[f"{tag}_{k}"] = getattr(V, k).copy()
        lv = O.ingest(lv_raw)  # ingest itself is pinned by the full-pipeline golden below
        min_r = np.full(N, np.finfo(np.float64).max)
        agg_r = np.zeros(N)
        rng = np.random.default_rng(3)
        labs, trees, aggs = [], [], []
        for k in range(12):
            t = int(rng.integers(0, V.T))
            lab = np.float32([rng.uniform(-.05, .05), rng.uniform(-.05, .05), rng.uniform(-2, D + 2)])
            if k == 3:
                lab = np.float32([0, 0, 5.0])
            V.eval_proposal(lv, t, lab, min_r, agg_r)
            labs.append(lab); trees.append(t); aggs.append(agg_r.copy())
        out[f"{tag}_prop_labels"] = np.stack(labs)
        out[f"{tag}_prop_trees"] = np.int32(trees)
        out[f"{tag}_prop_agg"] = np.stack(aggs)
        out[f"{tag}_prop_min"] = min_r.copy()
        out[f"{tag}_prop_abc"] = V.get_abc()
        R.srand(1)
        for it in range(2):
            V.mst_pms(lv, min_r, agg_r)
        out[f"{tag}_pms_min"] = min_r.copy()
        out[f"{tag}_pms_abc"] = V.get_abc()
        out[f"{tag}_pms_disp"] = V.label_to_disp()
    rng = np.random.default_rng(5)
    left = rng.uniform(-2, D + 2, N).astype(np.float32)
    right = (left + rng.normal(0, 1.0, N)).astype(np.float32)
    out["lr_left"], out["lr_right"] = left, right
    out["lr_nofill"] = R.lr_check(left, right, W, H, D, 0)
    out["lr_fill"] = R.lr_check(left, right, W, H, D, 1)
    with tempfile.TemporaryDirectory() as td:
        R.srand(1)
        dl, dr = R.stereo3dmst(td, L, Rt, lv_raw, rv_raw, D)
    out["full_left_disp"], out["full_right_disp"] = dl, dr
    path = os.path.join(ROOT, "tests", "golden", "ref_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    make_medium(O, R)


def make_medium(O, R):
    """ref_medium.npz: the reference's full pipeline on an input that segments into SEVERAL trees under its own
    literals (c = 5000, min size 200), so the 100 x 2 rounds exercise neighbour-tree propagation (ref_small is one tree).
    Only what cannot be regenerated is stored: the images and the reference's two disparity maps; the cost volume is
    the oracle's a2' volume x 1/6 (an input, rebuilt by the tests from the stored images)."""
    W, H, D = 200, 150, 16
    L, Rt, _ = synth.make_natural_pair(W, H, D, seed=23)
    lv_raw, rv_raw = O.cost_adgrad(L, Rt, D)
    lv_raw = (lv_raw * np.float32(1 / 6.0)).astype(np.float32)
    rv_raw = (rv_raw * np.float32(1 / 6.0)).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        R.srand(1)
        dl, dr = R.stereo3dmst(td, L, Rt, lv_raw, rv_raw, D)
    T = (O.forest(L).T, O.forest(Rt).T)
    assert min(T) >= 4, T
    path = os.path.join(ROOT, "tests", "golden", "ref_medium.npz")
    np.savez_compressed(path, W=W, H=H, D=D, left=L, right=Rt, trees=np.int32(T), full_left_disp=dl, full_right_disp=dr)
    print("wrote", path, os.path.getsize(path), "bytes, trees", T)


if __name__ == "__main__":
    main()
