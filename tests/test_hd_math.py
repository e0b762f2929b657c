"""The per-element arithmetic shared by the CUDA kernels (stereomatch_b200/csrc/hd_math.h), compiled for the
host and compared bit-for-bit with the oracle.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from stereomatch_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hd(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hd") / "libhd.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "models", "hd_math_host.cpp")])
    L = C.CDLL(out)
    L.hd_label_cost.restype = C.c_float
    L.hd_label_cost.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float]
    L.hd_label_disp.restype = C.c_float
    L.hd_label_disp.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
    L.hd_ingest.restype = C.c_float
    L.hd_ingest.argtypes = [C.c_float] * 4
    return L


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_median_network(hd, oracle):
    rng = np.random.default_rng(0)
    for shape in ((1, 1), (3, 9), (40, 57)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        out = np.empty_like(img)
        hd.hd_median3(p(img), shape[1], shape[0], p(out))
        assert np.array_equal(out, oracle.median3(img))
    img = rng.integers(0, 4, (64, 64), dtype=np.uint8)  # many ties
    out = np.empty_like(img)
    hd.hd_median3(p(img), 64, 64, p(out))
    assert np.array_equal(out, oracle.median3(img))


def test_adgrad_cost(hd, oracle):
    for (W, H, D, seed) in ((40, 9, 12, 1), (17, 5, 20, 2), (64, 8, 64, 3)):
        L, R, _ = synth.make_pair(W, H, max(D, 12), seed=seed)
        lv, rv = oracle.cost_adgrad(L, R, D)
        lv2 = np.empty_like(lv)
        rv2 = np.empty_like(rv)
        hd.hd_cost_adgrad(p(L), p(R), W, H, D, p(lv2), p(rv2))
        assert np.array_equal(lv.view(np.uint32), lv2.view(np.uint32))
        assert np.array_equal(rv.view(np.uint32), rv2.view(np.uint32))


def test_label_cost_against_oracle_single_node_trees(hd, oracle):
    """compute3DLabelCost through the oracle: with min_size=2 and c tiny most trees are small; the aggregated
    cost of a proposal on the ROOT of a 2-node tree with weight w is cost(root) + w*cost(leaf) — instead of
    unpicking that, compare on a forest of one-pixel-wide images where every tree is a path and use the
    brute-force identity.  Simpler and exact: a 1x1 image is one single-node tree, agg == cost."""
    rng = np.random.default_rng(5)
    D = 10
    labels = [np.float32(l) for l in ([0, 0, 3.0], [0.01, -0.02, 4.5], [np.nan, 0, 1], [0, 0, -1.0], [0, 0, 9.0], [0, 0, 9.5],
                                      [0, 0, 3e9], [0, 0, -3e9], [0.3, 0.1, 7.25], [0, 0, np.inf], [0, 0, -0.5], [0, 0, 8.999])]
    img = np.zeros((1, 1, 3), np.uint8)
    F = oracle.forest(img, c=1.0, min_size=1)
    assert F.T == 1 and F.N == 1
    for lab in labels:
        vol = rng.uniform(0, 0.5, (D, 1)).astype(np.float32)
        mn = np.full(1, np.finfo(np.float64).max)
        abc = np.zeros((1, 3), np.float32)
        _, agg = oracle.eval_proposal(F, vol, D, 0, lab, mn, abc)
        row = np.ascontiguousarray(vol[:, 0])
        got = hd.hd_label_cost(p(row), lab[0], lab[1], lab[2], 0, 0, D, 0.5)
        assert np.float64(got).view(np.uint64) == agg[0:1].view(np.uint64)[0], lab


def test_label_disp_and_ingest(hd, oracle):
    rng = np.random.default_rng(6)
    W, H, D = 31, 7, 10
    N = W * H
    abc = rng.uniform(-0.1, 0.1, (N, 3)).astype(np.float32)
    abc[:, 2] = rng.uniform(-3, D + 3, N)
    abc[3] = np.nan
    want = oracle.label_to_disp(abc, W, H, D) * np.float32(D - 1.0)
    got = np.float32([hd.hd_label_disp(abc[i, 0], abc[i, 1], abc[i, 2], i % W, i // W, D) for i in range(N)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    v = np.float32([np.nan, 0.7, 0.2, -1.0, 0.5])
    for (cap, off, sc) in ((0.5, 0.0, 1.0), (0.5, 1.0, 0.5)):
        want = oracle.ingest(v, cap, off, sc)
        got = np.float32([hd.hd_ingest(x, cap, off, sc) for x in v])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_plane_cost_and_gradients(hd, oracle):
    """pms_cost_mode 1 (north-star item 1, slanted variant): the shared header's gradients equal cv2's (cvtColor + Sobel/8,
    pm.cpp:70-88) and the oracle's; its per-pixel plane cost equals the oracle's restatement of pm.cpp:97-154 bit for bit,
    for both views, incl. planes leaving [0, Dmax], matches beyond both image borders and integer disparities."""
    import cv2
    hd.hd_pm_gradients.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    hd.hd_plane_cost_map.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int] + [C.c_float] * 5 + [C.c_void_p]
    rng = np.random.default_rng(3)
    for (W, H, D, seed, nat) in ((97, 61, 24, 3, 1), (64, 7, 16, 4, 0), (2, 5, 4, 5, 0)):
        L, R, _ = (synth.make_natural_pair if nat else synth.make_pair)(W, H, max(D, 12), seed=seed)
        grads = []
        for img in (L, R):
            g = np.empty((H, W, 2), np.float32)
            hd.hd_pm_gradients(p(img), W, H, p(g))
            gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
            assert np.array_equal(g[..., 0], cv2.Sobel(gray, cv2.CV_32F, 1, 0, ksize=3) / np.float32(8))
            assert np.array_equal(g[..., 1], cv2.Sobel(gray, cv2.CV_32F, 0, 1, ksize=3) / np.float32(8))
            assert np.array_equal(g, oracle.pm_gradients(img))
            grads.append(g)
        labels = [(0.0, 0.0, 3.0), (0.0, 0.0, 3.5), (0.01, -0.02, 5.25), (-0.3, 0.1, 9.0), (0.0, 0.0, -1.0), (0.0, 0.0, D + 0.5), (0.0, 0.0, float(D))]
        labels += [tuple(x) for x in np.stack([rng.uniform(-.05, .05, 6), rng.uniform(-.05, .05, 6), rng.uniform(-2, D + 2, 6)], 1)]
        for view in (0, 1):
            for lab in labels:
                lab = np.float32(lab)
                out = np.empty(W * H, np.float32)
                hd.hd_plane_cost_map(view, p(L), p(R), p(grads[0]), p(grads[1]), W, H, lab[0], lab[1], lab[2], D, 0.9, 10.0, 2.0, 0.25, 0.5, p(out))
                want = oracle.plane_cost_map(view, L, R, lab, D, scale=0.25)
                assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (W, view, lab)
