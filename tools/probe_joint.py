"""Development probe: front of a C4 batch of 8 (forests + costs); with S3_FH_JOINT=1 the 16 views share ONE forest-kernel
launch (the regime a profiler can capture).  argv: [reps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereomatch_b200 import api, synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
W, H, D, B = 1920, 1080, 256, 8
engs = []
for i in range(B):
    L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + 10 + i)
    e = api.Stereo3DMST(fh_ctas=36)
    e.set_images(L, R)
    engs.append(e)
for _ in range(2): api.batch_front(engs, D)
for e in engs: e.sync()
t0 = time.perf_counter()
for _ in range(reps):
    api.batch_front(engs, D)
    for e in engs: e.sync()
ms = (time.perf_counter() - t0) / reps * 1e3
T = [e.forest_info(0)[0] for e in engs]
print(json.dumps({"front_ms": round(ms, 2), "trees": T, "env": {k: v for k, v in os.environ.items() if k.startswith("S3_")}}), flush=True)
