"""The C-ABI shared library loads and exports every symbol include/s3dmst.h declares; without a GPU it must
refuse to create a context (no CPU fallback).  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from stereomatch_b200 import build
    build.build()
    from stereomatch_b200 import api
    return api.load_library()


def test_exports_match_header(lib):
    from stereomatch_b200 import api
    hdr = open(os.path.join(ROOT, "include", "s3dmst.h")).read()
    declared = sorted(set(re.findall(r"\b(s3dmst_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(api.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_params_defaults_are_the_reference_literals(lib):
    from stereomatch_b200 import api
    p = api.default_params()
    assert (p.fh_c, p.min_cc_size, p.median, p.cost_cap, p.oob_cost, p.num_iter, p.exact) == (5000.0, 200, 3, 0.5, 0.5, 100, 1)
    assert abs(p.gamma - 1.0 / 12.0) < 1e-7 and abs(p.refine_floor - 0.1) < 1e-7


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stereomatch_b200 import api
    with pytest.raises(api.S3Error):
        api.Stereo3DMST()


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "stereomatch_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "import oracle" not in src, f


def test_params_struct_layout_matches_the_header(tmp_path):
    """Field order, types, offsets and size of s3dmst_params: the header compiled by gcc against the ctypes mirror."""
    import ctypes as C
    import subprocess
    from stereomatch_b200 import api
    hdr = open(os.path.join(ROOT, "include", "s3dmst.h")).read()
    body = re.search(r"typedef struct s3dmst_params \{(.*?)\} s3dmst_params;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(float|int)\s+([a-z_0-9]+)\s*;", body)
    ctype = {"float": C.c_float, "int": C.c_int}
    assert [(n, ctype[t]) for t, n in fields] == list(api.S3Params._fields_)
    src = tmp_path / "layout.c"
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "s3dmst.h"', "int main(void) {",
             '    printf("%zu\\n", sizeof(s3dmst_params));']
    lines += [f'    printf("%zu\\n", offsetof(s3dmst_params, {n}));' for _, n in fields]
    lines += ["    return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == C.sizeof(api.S3Params)
    assert out[1:] == [getattr(api.S3Params, n).offset for _, n in fields]
