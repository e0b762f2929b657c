"""Development probe: a C4 batch of 8 frames as ONE batch call vs software-pipelined groups (front(A) back(A) front(B) back(B)):
a group's forests (latency / random-access bound) run beside the other group's aggregation (bandwidth bound).
argv: [groups] [reps] [fh_ctas]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stereomatch_b200 import api, synth
groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
fhc = int(sys.argv[3]) if len(sys.argv) > 3 else 36
W, H, D, B = 1920, 1080, 256, 8
engs = []
for i in range(B):
    L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + 10 + i)
    e = api.Stereo3DMST(fh_ctas=fhc)
    e.set_images(L, R)
    engs.append(e)
def sync():
    for e in engs: e.sync()
def step_joint():
    api.run_dense_batch(engs, D, fill=True, fetch=False)
def step_piped():
    per = B // groups
    for g in range(groups):
        api.batch_front(engs[g * per:(g + 1) * per], D)
        api.batch_back(engs[g * per:(g + 1) * per], D, fill=True)
res = {}
ref = None
for name, fn in (("joint", step_joint), ("piped", step_piped)):
    for _ in range(3): fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    sync()
    res[name] = round((time.perf_counter() - t0) / reps * 1e3, 2)
    maps = [np.concatenate([e.get_disparity(0), e.get_disparity(1)]) for e in (engs[0], engs[-1])]
    if ref is None: ref = maps
    else: res["same"] = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(ref, maps))
print(json.dumps({"groups": groups, "fh_ctas": fhc, **res}), flush=True)
