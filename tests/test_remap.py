"""Rectification front-end: the oracle's restatement of OpenCV's fixed-point bilinear remap against
OpenCV itself and against the committed golden vectors, and the library's weight table against the oracle's.
The GPU kernel is checked in tests/test_gpu_parity.py::test_remap_matches_opencv."""
import os

import numpy as np
import pytest

from oracle import remap_oracle as ro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "remap_small.npz")


def test_table_sums_and_saturation():
    tab = ro.bilinear_tab().astype(np.int64)
    assert tab.shape == (1024, 4) and (tab.sum(1) == 32768).all()
    assert tab[0].tolist() == [32767, 0, 0, 1]          # the saturated unit weight and OpenCV's fix-up of it
    assert tab[33].tolist() == [30752, 992, 992, 32]    # fy = fx = 1/32


def test_oracle_matches_golden_vectors():
    g = np.load(GOLD)
    for key in ("00", "01", "10", "11"):
        got = ro.remap_fixed(g["src_" + key], g["xy_" + key], g["fxy_" + key])
        assert np.array_equal(got, g["exp_" + key]), key


def test_oracle_matches_opencv_on_random_maps():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(12):
        Hs, Ws = int(rng.integers(2, 60)), int(rng.integers(2, 60))
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        src = rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8)
        mxy = np.stack([rng.integers(-4, Ws + 4, (H, W)), rng.integers(-4, Hs + 4, (H, W))], -1).astype(np.int16)
        mf = rng.integers(0, 1024, (H, W)).astype(np.uint16)
        assert np.array_equal(ro.remap_fixed(src, mxy, mf), cv2.remap(src, mxy, mf, cv2.INTER_LINEAR))


def test_library_weight_table_is_the_oracles():
    from stereomatch_b200 import api, build
    build.build()
    assert np.array_equal(api.remap_table(), ro.bilinear_tab())
