"""Development probe: the reference's own pipeline (s3dmst_run: plane init, num_iter x MST_PMS per view, LabelToDisp, LR check) on
one pair.  argv: case (flir|c2) [num_iter] [pms_cost_mode]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereomatch_b200 import api, synth
case = sys.argv[1] if len(sys.argv) > 1 else "flir"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if case == "flir":
    import cv2
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    L = cv2.imread(os.path.join(g, "flir_000020_left.jpg")); R = cv2.imread(os.path.join(g, "flir_000020_right.jpg")); D = 100
else:
    L, R, _ = synth.make_pair(1280, 720, 128); D = 128
kw = dict(num_iter=iters, cost_scale=1 / 6.0)
if mode:
    kw.update(pms_cost_mode=1, cost_scale=0.25, oob_cost=30.0)
eng = api.Stereo3DMST(**kw)
eng.set_images(L, R)
eng.run(D, seed=1, fetch=False); eng.sync()
t0 = time.perf_counter()
dl, dr = eng.run(D, seed=2)
dt = time.perf_counter() - t0
T = [eng.forest_info(v)[0] for v in (0, 1)]
print(json.dumps({"case": case, "num_iter": iters, "pms_cost_mode": mode, "ms_per_pair": round(dt * 1e3, 1), "ms_per_round_per_view": round(dt * 1e3 / (2 * iters), 3),
                  "trees": T, "valid_fraction": float((dl > 0).mean()), "launches": eng.launch_count()}), flush=True)
eng.close()
