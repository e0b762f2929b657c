"""Concurrency probe: B contexts build their forests at the same time (one host thread each)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereomatch_b200 import api, synth
W, H, D = 1280, 720, 128
ctas = int(sys.argv[1]); threads = int(sys.argv[2])
for B in (1, 2, 4, 8):
    engs = []
    for i in range(B):
        L, R, _ = synth.make_pair(W, H, D, seed=synth.BASE_SEED + i)
        e = api.Stereo3DMST(fh_ctas=ctas, fh_threads=threads)
        e.set_images(L, R)
        engs.append(e)
    def work(e, n):
        for _ in range(n):
            e.build_forest(0)
        e.sync()
    for n in (2, 6):
        ts = [threading.Thread(target=work, args=(e, n)) for e in engs]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        dt = (time.perf_counter() - t0) * 1e3
    print(f"ctas {ctas} threads {threads} B={B}: {dt / 6:.2f} ms per round of {B} single-view forests -> {dt / 6 / B:.2f} ms each; stage(dev) {engs[0].stage_ms(api.T_FOREST):.2f}")
    for e in engs: e.close()
