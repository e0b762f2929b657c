"""Probe: software pipelining of consecutive batches (front of batch k+1 overlaps back of batch k)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereomatch_b200 import api, synth
W, H, D = 1280, 720, 128
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sets = []
for s in range(2):
    engs = []
    for i in range(B):
        L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + s * B + i)
        e = api.Stereo3DMST(fh_ctas=36)
        e.set_images(L, R)
        engs.append(e)
    sets.append(engs)
# plain: run_dense_batch one after the other
for _ in range(2):
    api.run_dense_batch(sets[0], D, fill=True, fetch=False)
t0 = time.perf_counter()
for k in range(steps):
    api.run_dense_batch(sets[k & 1], D, fill=True, fetch=False)
for e in sets[0] + sets[1]: e.sync()
plain = (time.perf_counter() - t0) * 1e3 / steps
# pipelined
api.batch_front(sets[0], D)
t0 = time.perf_counter()
for k in range(steps):
    a, b = sets[k & 1], sets[(k + 1) & 1]
    t1 = threading.Thread(target=api.batch_back, args=(a, D, True))
    t2 = threading.Thread(target=api.batch_front, args=(b, D))
    t1.start(); t2.start(); t1.join(); t2.join()
for e in sets[0] + sets[1]: e.sync()
pipe = (time.perf_counter() - t0) * 1e3 / steps
print(f"B={B}: plain {plain:.2f} ms per batch ({plain / B:.2f} per pair), pipelined {pipe:.2f} ms per batch ({pipe / B:.2f} per pair)")
for e in sets[0] + sets[1]: e.close()
