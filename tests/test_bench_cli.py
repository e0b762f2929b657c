"""bench.py host-side contract (no GPU): the one-JSON-line guard, the workload description shared by both arms, the
algorithmic-byte model and the provenance of the traffic figure."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_only_the_json_line_reaches_stdout():
    code = (
        "import os, sys, json; sys.path.insert(0, %r); import bench\n"
        "bench.guard_stdout()\n"
        "os.write(1, b'NCCL version 2.x (a library printing to fd 1)\\n')\n"
        "print('a python print after the guard')\n"
        "bench.emit(json.dumps({'metric': 'x', 'value': 1}))\n" % ROOT
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout) == {"metric": "x", "value": 1}          # stdout is exactly one JSON line
    assert "NCCL version" in r.stderr and "a python print" in r.stderr   # everything else went to stderr


def test_config_and_byte_model():
    import bench
    bench.FUSE = True
    cfg = bench.gpu_config(1, 8)
    assert cfg == bench.gpu_config(1, 8) and "workload" in cfg and "model" not in cfg
    assert "no cost volume" in cfg["mode"]
    assert bench.alg_bytes_per_pxlabel() == 16.0       # 4 B cost build + 12 B aggregation: the fused kernel does both
    tpp, src = bench.read_traffic()
    assert tpp is not None and 15.0 < tpp < 22.0 and "profiles/" in src and os.path.exists(os.path.join(ROOT, src.split(":")[0]))
    bench.FUSE = False
    assert bench.alg_bytes_per_pxlabel() == 12.0
    assert "cost + running sums" in bench.gpu_config(1, 8)["l2"]
    tpp2, src2 = bench.read_traffic()
    assert tpp2 is not None and 19.0 < tpp2 < 22.0 and os.path.exists(os.path.join(ROOT, src2.split(":")[0]))
    bench.FUSE = True


def test_cli_flags():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--no-fuse", "--no-cpu", "--no-extras"):
        assert flag in r.stdout
