"""Development probe: how the front of a batch (forests + cost volumes) scales with the number of frames in flight.
argv: [fh_ctas] [case c4|c2]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereomatch_b200 import api, synth
fhc = int(sys.argv[1]) if len(sys.argv) > 1 else 36
case = sys.argv[2] if len(sys.argv) > 2 else "c4"
W, H, D = (1920, 1080, 256) if case == "c4" else (1280, 720, 128)
B = 8
engs = []
for i in range(B):
    L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + 10 + i)
    e = api.Stereo3DMST(fh_ctas=fhc)
    e.set_images(L, R)
    engs.append(e)
out = {"fh_ctas": fhc, "case": case, "env": {k: v for k, v in os.environ.items() if k.startswith("S3_")}}
for n in (1, 2, 4, 8):
    sub = engs[:n]
    for _ in range(2): api.batch_front(sub, D)
    for e in sub: e.sync()
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        api.batch_front(sub, D)
        for e in sub: e.sync()
    ms = (time.perf_counter() - t0) / reps * 1e3
    out["front_%d" % n] = round(ms, 2)
    out["forest_%d" % n] = round(float(np.mean([e.stage_ms(0) for e in sub])), 2)
    out["cost_%d" % n] = round(float(np.mean([e.stage_ms(1) for e in sub])), 2)
print(json.dumps(out), flush=True)
