"""One C2 step for ncu: python tools/profile_step.py [n_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereomatch_b200 import api, synth
W, H, D = 1280, 720, 128
L, R, _ = synth.make_pair(W, H, D)
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("="); kw[k] = int(v)
eng = api.Stereo3DMST(**kw)
eng.set_images(L, R)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    eng.run_dense(D, fill=True, fetch=False)
eng.sync()
print("stage ms:", [round(eng.stage_ms(s), 3) for s in range(4)], "launches", eng.launch_count())
eng.close()
