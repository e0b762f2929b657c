"""The label-sharded path called from plain C through include/s3dmst.h: tests/models/comm_driver.c
is compiled with gcc against libs3dmst.so and run with a one-rank communicator on the test box's single GPU (NCCL still
executes both all-reduces of the MIN-LOC); bench.py --gpus N covers N >= 2 on the driver's multi-GPU runs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_build")
SO = os.path.join(OUT, "libcommtest.so")


@pytest.fixture(scope="module")
def drv():
    import torch  # noqa: F401  (its bundled NCCL must be the copy this process holds: see stereomatch_b200.api._prefer_torch_nccl)
    from stereomatch_b200 import build
    build.build()
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "tests", "models", "comm_driver.c")
    libdir = os.path.join(ROOT, "stereomatch_b200")
    if True:   # always rebuilt (a copied tree's mtimes prove nothing)
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                               "-L", libdir, "-ls3dmst", f"-Wl,-rpath,{libdir}", "-o", SO])
    return C.CDLL(SO)


def test_header_compiles_as_c_and_unique_id_needs_no_gpu(drv):
    """include/s3dmst.h is a C header (the driver is compiled as C99 with -Wall -Werror); the ncclUniqueId comes from NCCL
    bound at run time."""
    buf = C.create_string_buffer(128)
    assert drv.shard_make_id(buf) == 0
    assert any(b != 0 for b in buf.raw)


@pytest.mark.gpu
def test_sharded_path_from_c_single_rank(drv, oracle):
    from stereomatch_b200 import synth
    W, H, D = 200, 120, 44
    L, R, _ = synth.make_pair(W, H, D, seed=17)
    idb = C.create_string_buffer(128)
    assert drv.shard_make_id(idb) == 0
    dl = np.empty(W * H, np.float32); dr = np.empty(W * H, np.float32)
    d0, d1, ms = C.c_int(), C.c_int(), C.c_double()
    err = C.create_string_buffer(512)
    rc = drv.shard_run(idb, 0, 1, 0, L.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p), W, H, D, dl.ctypes.data_as(C.c_void_p),
                       dr.ctypes.data_as(C.c_void_p), C.byref(d0), C.byref(d1), C.byref(ms), err, 512)
    assert rc == 0, err.value.decode()
    assert (d0.value, d1.value) == (0, D) and ms.value > 0.0
    lv, rv = oracle.cost_adgrad(L, R, D)
    dlo = oracle.aggregate_dense(oracle.forest(L), lv)[0].astype(np.float32)
    dro = oracle.aggregate_dense(oracle.forest(R), rv)[0].astype(np.float32)
    want, _ = oracle.lr_check(dlo, dro, W, H, D, True)
    assert np.array_equal(dl.view(np.uint32), want.view(np.uint32)) and np.array_equal(dr.view(np.uint32), dro.view(np.uint32))
