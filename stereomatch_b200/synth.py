"""Synthetic stereo inputs for the configs in BASELINE.json.

Random-dot slanted-plane pairs with known left-view ground truth: the right image is a field of
2x2-pixel uniform-random BGR dots (so the 3x3 median of the forest stage keeps texture); the left
image is the right one sampled bilinearly at x - d(x, y), where d is one slanted plane per cell of
a 3x2 grid, d in [4, D-5]; left pixels that look outside the right image get fresh noise.
Pure numpy; no dependency on the CUDA library or the oracle.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20261018


def make_pair(W: int, H: int, D: int, seed: int = BASE_SEED):
    """Returns (left_bgr u8 [H,W,3], right_bgr u8 [H,W,3], gt_left f32 [H,W])."""
    rng = np.random.default_rng(seed)
    dots = rng.integers(0, 256, size=((H + 1) // 2, (W + 1) // 2, 3), dtype=np.uint8)
    right = np.repeat(np.repeat(dots, 2, axis=0), 2, axis=1)[:H, :W].copy()

    gt = np.zeros((H, W), np.float32)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    lo, hi = 4.0, float(max(5, D - 5))
    for gy in range(2):
        for gx in range(3):
            y0, y1 = gy * H // 2, (gy + 1) * H // 2
            x0, x1 = gx * W // 3, (gx + 1) * W // 3
            a, b = rng.uniform(-0.02, 0.02, size=2)
            if y1 <= y0 or x1 <= x0:
                continue
            xx, yy = xs[y0:y1, x0:x1], ys[y0:y1, x0:x1]
            slope = a * (xx - x0) + b * (yy - y0)
            smin, smax = float(slope.min()), float(slope.max())
            span = smax - smin
            if span > hi - lo:  # squeeze very steep planes into range
                slope = slope * ((hi - lo) / span)
                smin, smax = float(slope.min()), float(slope.max())
            c = rng.uniform(lo - smin, hi - smax)
            gt[y0:y1, x0:x1] = slope + c

    src = xs - gt
    x0i = np.floor(src).astype(np.int64)
    fx = (src - x0i).astype(np.float32)[..., None]
    valid = (x0i >= 0) & (x0i + 1 < W)
    x0c = np.clip(x0i, 0, W - 1)
    x1c = np.clip(x0i + 1, 0, W - 1)
    rows = np.arange(H)[:, None]
    r0 = right[rows, x0c].astype(np.float32)
    r1 = right[rows, x1c].astype(np.float32)
    left = np.rint((1.0 - fx) * r0 + fx * r1).astype(np.uint8)
    noise = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    left[~valid] = noise[~valid]
    return left, right, gt


def make_natural_pair(W: int, H: int, D: int, seed: int = BASE_SEED):
    """Smoother, image-like texture (sum of blurred noise octaves): gives FLIR-like forests
    (few large, deep trees) instead of the shallow forests of pure random dots."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((H, W, 3), np.float32)
    amp = 1.0
    k = 32
    while k >= 1:
        g = rng.normal(size=((H + k - 1) // k + 1, (W + k - 1) // k + 1, 3)).astype(np.float32)
        up = np.repeat(np.repeat(g, k, axis=0), k, axis=1)[:H, :W]
        acc += amp * up
        amp *= 0.6
        k //= 2
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    right = np.rint(acc * 255).astype(np.uint8)
    gt = np.zeros((H, W), np.float32)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    a, b = rng.uniform(-0.01, 0.01, size=2)
    slope = a * xs + b * ys
    gt[:] = slope - slope.min() + 4.0
    gt = np.minimum(gt, D - 5).astype(np.float32)
    src = xs - gt
    x0i = np.floor(src).astype(np.int64)
    fx = (src - x0i).astype(np.float32)[..., None]
    valid = (x0i >= 0) & (x0i + 1 < W)
    x0c = np.clip(x0i, 0, W - 1)
    x1c = np.clip(x0i + 1, 0, W - 1)
    rows = np.arange(H)[:, None]
    left = np.rint((1.0 - fx) * right[rows, x0c].astype(np.float32) + fx * right[rows, x1c].astype(np.float32)).astype(np.uint8)
    noise = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    left[~valid] = noise[~valid]
    return left, right, gt


def make_proposals(tree_sizes, abc_gt_of_tree, D: int, n_iter: int = 8, per_tree: int = 12, seed: int = BASE_SEED + 2):
    """Fixed injected proposal sequence (config 3): for each iteration k and tree, `per_tree`
    labels = a reference plane of the tree perturbed by N(0, 2^-k * D/4) in c and small tilts."""
    rng = np.random.default_rng(seed)
    T = len(tree_sizes)
    trees, labels = [], []
    for k in range(n_iter):
        sd = (2.0 ** -k) * D / 4.0
        for t in range(T):
            a0, b0, c0 = abc_gt_of_tree[t]
            for _ in range(per_tree):
                trees.append(t)
                labels.append((a0 + rng.normal(0, 0.01 * 2.0 ** -k), b0 + rng.normal(0, 0.01 * 2.0 ** -k), c0 + rng.normal(0, sd)))
    return np.asarray(trees, np.int32), np.asarray(labels, np.float32)
