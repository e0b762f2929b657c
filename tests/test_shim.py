"""The header-compatible C++ entry point (stereomatch_b200/csrc/stereo3dmst_shim.cpp): same symbol, argument
list and error behaviour as include/Stereo3DMST.h:7 / src/Stereo3DMST.cpp:714-759, compiled here against the
cv::Mat stand-in of oracle/ref_shims (no OpenCV C++ headers in this image) and called like stereo_Yin.cpp:205-210."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_build")
SO = os.path.join(OUT, "libshimtest.so")


@pytest.fixture(scope="module")
def shim():
    from stereomatch_b200 import build
    build.build()
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(ROOT, "stereomatch_b200", "csrc", "stereo3dmst_shim.cpp"), os.path.join(ROOT, "tests", "models", "shim_driver.cpp")]
    libdir = os.path.join(ROOT, "stereomatch_b200")
    if True:   # always rebuilt (2 s): the parameter struct lives in the header, and a copied tree's mtimes prove nothing
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w", "-I", os.path.join(ROOT, "oracle", "ref_shims"),
                               "-I", os.path.join(ROOT, "include"), *srcs, "-L", libdir, "-ls3dmst", f"-Wl,-rpath,{libdir}", "-o", SO])
    return C.CDLL(SO)


def _call(lib, L, R, D, cost):
    H, W = L.shape[:2]
    ol = np.full((H, W), -7.0, np.float32); orr = np.full((H, W), -7.0, np.float32)
    ms = C.c_double(); rows = C.c_int(); cols = C.c_int()
    lib.shim_call.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = lib.shim_call(L.ctypes.data, R.ctypes.data, W, H, D, cost.encode(), ol.ctypes.data, orr.ctypes.data, C.byref(ms),
                       C.byref(rows), C.byref(cols))
    return rc, ol, orr, ms.value, rows.value, cols.value


def test_shim_exports_the_reference_symbols(shim):
    out = subprocess.check_output(["nm", "-D", "--defined-only", SO], text=True)
    assert " T stereo3dmst" in out            # unmangled, as `nm build/StereoYin` shows for the reference
    assert "_Z10startTimerv" in out and "_Z8getTimerv" in out


def test_shim_unknown_data_cost_behaves_like_the_reference(shim, capfd):
    """Stereo3DMST.cpp:756-759: prints "wrong data cost" and returns; outputs are created (:722-723) but unfilled."""
    from stereomatch_b200 import synth
    L, R, _ = synth.make_pair(64, 48, 16, seed=1)
    rc, ol, orr, ms, rows, cols = _call(shim, L, R, 16, "SGM")
    assert rc == 0 and (rows, cols) == (48, 64) and ms >= 0.0
    assert "wrong data cost" in capfd.readouterr().out


def test_shim_missing_mccnn_volume_returns_like_a_failed_step(shim, capfd, tmp_path, monkeypatch):
    from stereomatch_b200 import synth
    monkeypatch.chdir(tmp_path)
    L, R, _ = synth.make_pair(64, 48, 16, seed=1)
    rc, ol, orr, ms, rows, cols = _call(shim, L, R, 16, "MCCNN_acrt")
    assert rc == 0 and (rows, cols) == (48, 64)
    assert "left.bin" in capfd.readouterr().out


def test_shim_rejects_bad_inputs_before_touching_them(shim, capfd):
    """Dmax <= 0 (and, in C++ callers, empty / non-CV_8UC3 / differently sized Mats) return with a message before any
    buffer is sized from them; the reference would compute Dmax*rows*cols from them unchecked."""
    from stereomatch_b200 import synth
    L, R, _ = synth.make_pair(64, 48, 16, seed=1)
    rc, ol, orr, ms, rows, cols = _call(shim, L, R, 0, "ADGRAD")
    assert rc == -1 and (rows, cols) == (0, 0)          # outputs not even created
    assert "Dmax > 0" in capfd.readouterr().out
    shim.shim_call_mismatched.argtypes = [C.c_int, C.c_int]
    assert shim.shim_call_mismatched(64, 48) == 0       # right image smaller / other type: returns, nothing read
    assert "CV_8UC3" in capfd.readouterr().out


@pytest.mark.gpu
def test_shim_default_is_the_reference_pipeline(shim, monkeypatch):
    """With unchanged arguments the drop-in runs what the reference runs (plane init, 100 rounds of MST_PMS per view,
    LabelToDisp, LR check without fill): sub-pixel disparities, identical to s3dmst_run on the same inputs and seed."""
    from stereomatch_b200 import api, synth
    monkeypatch.delenv("S3DMST_MODE", raising=False)
    W, H, D = 160, 96, 24
    L, R, gt = synth.make_pair(W, H, D, seed=9)
    rc, dl, dr, ms, _, _ = _call(shim, L, R, D, "ADGRAD")
    assert rc == 0
    eng = api.Stereo3DMST(cost_scale=1 / 6.0)
    eng.set_images(L, R)
    wl, wr = eng.run(D, seed=1, fill=False)
    eng.close()
    assert np.array_equal(dl.view(np.uint32).ravel(), wl.view(np.uint32)) and np.array_equal(dr.view(np.uint32).ravel(), wr.view(np.uint32))
    assert np.any(dl != np.round(dl)) and dl.min() >= 0.0 and dl.max() <= D - 1.0
    valid = dl > 0
    assert valid.mean() > 0.3 and (np.abs(dl - gt)[valid] <= 1.0).mean() > 0.5


@pytest.mark.gpu
def test_shim_dense_matches_oracle(shim, tmp_path, monkeypatch):
    from oracle.pyoracle import Oracle
    from stereomatch_b200 import synth
    O = Oracle()
    monkeypatch.setenv("S3DMST_MODE", "dense")
    W, H, D = 160, 96, 24
    L, R, _ = synth.make_pair(W, H, D, seed=9)
    rc, dl, dr, ms, _, _ = _call(shim, L, R, D, "ADGRAD")
    assert rc == 0
    lv, rv = O.cost_adgrad(L, R, D)
    dlo = O.aggregate_dense(O.forest(L), lv)[0].astype(np.float32)
    dro = O.aggregate_dense(O.forest(R), rv)[0].astype(np.float32)
    want, _ = O.lr_check(dlo, dro, W, H, D, False)
    assert np.array_equal(dl.view(np.uint32).ravel(), want.view(np.uint32).ravel())
    assert np.array_equal(dr.view(np.uint32).ravel(), dro.view(np.uint32).ravel())
    # the reference's on-disk input: mc-cnn-master/left.bin, right.bin = float32 [1][Dmax][rows][cols] (:764-775)
    monkeypatch.chdir(tmp_path)
    os.makedirs("mc-cnn-master")
    rng = np.random.default_rng(5)
    vl = rng.random((D, H, W), dtype=np.float32); vr = rng.random((D, H, W), dtype=np.float32)
    vl[3, 5, 7] = np.nan
    vl.tofile("mc-cnn-master/left.bin"); vr.tofile("mc-cnn-master/right.bin")
    rc, dl, dr, ms, _, _ = _call(shim, L, R, D, "MCCNN_acrt")
    assert rc == 0
    il, ir = O.ingest(vl.reshape(D, -1)), O.ingest(vr.reshape(D, -1))
    dlo = O.aggregate_dense(O.forest(L), il)[0].astype(np.float32)
    dro = O.aggregate_dense(O.forest(R), ir)[0].astype(np.float32)
    want, _ = O.lr_check(dlo, dro, W, H, D, False)
    assert np.array_equal(dl.view(np.uint32).ravel(), want.view(np.uint32).ravel())
    assert np.array_equal(dr.view(np.uint32).ravel(), dro.view(np.uint32).ravel())


@pytest.mark.gpu
def test_shim_mccnn_fst_rescale(shim, tmp_path, monkeypatch):
    """"MCCNN_fst": the fast nets score in [-1, 1]; the ingest maps (c + 1) / 2 before the 0.5 cap
    (Stereo3DMST.cpp:792, PatchMatchStereoGPU.cu:4713-4745)."""
    from oracle.pyoracle import Oracle
    from stereomatch_b200 import synth
    O = Oracle()
    monkeypatch.setenv("S3DMST_MODE", "dense")
    monkeypatch.chdir(tmp_path)
    W, H, D = 96, 64, 12
    L, R, _ = synth.make_pair(W, H, D, seed=4)
    os.makedirs("mc-cnn-master")
    rng = np.random.default_rng(6)
    vl = rng.uniform(-1, 1, (D, H, W)).astype(np.float32); vr = rng.uniform(-1, 1, (D, H, W)).astype(np.float32)
    vl[1, 2, 3] = np.nan
    vl.tofile("mc-cnn-master/left.bin"); vr.tofile("mc-cnn-master/right.bin")
    rc, dl, dr, ms, _, _ = _call(shim, L, R, D, "MCCNN_fst")
    assert rc == 0
    il, ir = O.ingest(vl.reshape(D, -1), 0.5, 1.0, 0.5), O.ingest(vr.reshape(D, -1), 0.5, 1.0, 0.5)
    assert il.min() >= 0.0 and il.max() <= 0.5
    dlo = O.aggregate_dense(O.forest(L), il)[0].astype(np.float32)
    dro = O.aggregate_dense(O.forest(R), ir)[0].astype(np.float32)
    want, _ = O.lr_check(dlo, dro, W, H, D, False)
    assert np.array_equal(dl.view(np.uint32).ravel(), want.view(np.uint32).ravel())
    assert np.array_equal(dr.view(np.uint32).ravel(), dro.view(np.uint32).ravel())
