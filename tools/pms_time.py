"""Timing of the proposal-search mode (the reference's own mode: MST_PMS rounds) at C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereomatch_b200 import api, synth
W, H, D = 1280, 720, 128
L, R, _ = synth.make_pair(W, H, D)
eng = api.Stereo3DMST(cost_scale=1 / 6.0)
eng.set_images(L, R)
eng.build_forest(0); eng.build_forest(1)
eng.build_cost_volume(D, ingest=True)
eng.init_labels(0, D)
eng.pms_iterate(0, 2, seed=1)
t0 = time.perf_counter()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
eng.pms_iterate(0, n, seed=2)
eng.sync()
dt = time.perf_counter() - t0
info = eng.forest_info(0, want_adj=True)
print(f"{n} MST_PMS rounds, one view: {dt * 1e3:.1f} ms = {dt / n * 1e3:.2f} ms per round (device k_pms {eng.stage_ms(api.T_PMS):.2f} ms last round); forest info {info}")
eng.close()
