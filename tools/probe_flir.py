"""Development probe: aggregation time of the bundled FLIR pair for one giant-tree threshold (argv[1])."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from stereomatch_b200 import api
if os.environ.get('S3_LIB'):
    api.LIB_PATH = os.environ['S3_LIB']
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
L = cv2.imread(os.path.join(g, "flir_000020_left.jpg")); R = cv2.imread(os.path.join(g, "flir_000020_right.jpg"))
cl = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = api.Stereo3DMST(agg_cluster_nodes=cl)
eng.set_images(L, R)
eng.run_dense(100, fill=True, fetch=False); eng.sync()
agg = []
for _ in range(reps):
    eng.aggregate_dense(0, 0, 100, fetch=False); eng.sync()
    agg.append(eng.stage_ms(api.T_AGG))
print(json.dumps({"cl": cl, "env": {k: v for k, v in os.environ.items() if k.startswith("S3_")}, "left_view_agg_ms": [round(a, 3) for a in agg]}), flush=True)
eng.close()
