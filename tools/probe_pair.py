"""Development probe: stage times of ONE pair alone on the GPU (argv: case [cluster_nodes]); case in flir, c2, c4."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereomatch_b200 import api, synth
case = sys.argv[1] if len(sys.argv) > 1 else "flir"
cl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fhc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if case == "flir":
    import cv2
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    L = cv2.imread(os.path.join(g, "flir_000020_left.jpg")); R = cv2.imread(os.path.join(g, "flir_000020_right.jpg")); D = 100
elif case == "c2":
    L, R, _ = synth.make_pair(1280, 720, 128); D = 128
else:
    L, R, _ = synth.make_pair(1920, 1080, 256, seed=synth.BASE_SEED + 10); D = 256
eng = api.Stereo3DMST(agg_cluster_nodes=cl, fh_cluster=fhc)
eng.set_images(L, R)
for _ in range(2):
    eng.run_dense(D, fill=True, fetch=False)
eng.sync()
st = np.zeros(4); reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    eng.run_dense(D, fill=True, fetch=False); eng.sync()
    st += [eng.stage_ms(i) for i in range(4)]
    if os.environ.get("PROBE_VERBOSE"): print("rep", [round(eng.stage_ms(i), 3) for i in range(4)], flush=True)
dt = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps({"case": case, "cl": cl, "fh_cluster": fhc, "env": {k: v for k, v in os.environ.items() if k.startswith("S3_")}, "ms_per_pair": round(dt, 3),
                  "forest": round(st[0] / reps, 3), "cost": round(st[1] / reps, 3), "aggregate": round(st[2] / reps, 3), "post": round(st[3] / reps, 3)}), flush=True)
eng.close()
