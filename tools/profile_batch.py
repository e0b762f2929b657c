"""A few batched C2 steps for ncu: python tools/profile_batch.py [batch] [n_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereomatch_b200 import api, synth
W, H, D = 1280, 720, 128
if len(sys.argv) > 4:
    W, H = int(sys.argv[3]), int(sys.argv[4])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
engs = []
for i in range(B):
    L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + i)
    e = api.Stereo3DMST(fh_ctas=36)
    e.set_images(L, R)
    engs.append(e)
for _ in range(steps):
    api.run_dense_batch(engs, D, fill=True, fetch=False)
for e in engs:
    e.sync()
print("agg ms:", round(engs[0].stage_ms(api.T_AGG), 3), "launches", sum(e.launch_count() for e in engs))
for e in engs:
    e.close()
