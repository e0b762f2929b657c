"""Development probe: stage times of one pair alone on the GPU for several giant-tree thresholds
(params.agg_cluster_nodes): the bundled FLIR pair (C1) and the C2 synthetic pair.  Prints one JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereomatch_b200 import api, synth


def run(tag, L, R, D, cl, reps=5):
    eng = api.Stereo3DMST(agg_cluster_nodes=cl)
    eng.set_images(L, R)
    for _ in range(2):
        eng.run_dense(D, fill=True, fetch=False)
    eng.sync()
    st = np.zeros(4)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.run_dense(D, fill=True, fetch=False)
        eng.sync()
        st += [eng.stage_ms(i) for i in range(4)]
    dt = (time.perf_counter() - t0) / reps * 1e3
    dl, dr = eng.run_dense(D, fill=True)
    eng.close()
    print(json.dumps({"case": tag, "agg_cluster_nodes": cl, "ms_per_pair": round(dt, 3), "forest": round(st[0] / reps, 3), "cost": round(st[1] / reps, 3),
                      "aggregate": round(st[2] / reps, 3), "post": round(st[3] / reps, 3)}), flush=True)
    return dl, dr


def main():
    import cv2
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    FL = cv2.imread(os.path.join(g, "flir_000020_left.jpg")); FR = cv2.imread(os.path.join(g, "flir_000020_right.jpg"))
    C2 = synth.make_pair(1280, 720, 128)
    C4 = synth.make_pair(1920, 1080, 256, seed=synth.BASE_SEED + 10)
    for tag, (L, R), D in (("C1 FLIR 2048x1536 D=100", (FL, FR), 100), ("C2 1280x720 D=128", C2[:2], 128), ("C4 frame 1920x1080 D=256", C4[:2], 256)):
        ref = None
        for cl in (-1, 32768, 8192, 2048):
            out = run(tag, L, R, D, cl)
            if ref is None:
                ref = out
            else:
                assert np.array_equal(ref[0].view(np.uint32), out[0].view(np.uint32)) and np.array_equal(ref[1].view(np.uint32), out[1].view(np.uint32)), "cluster walk changed the result"


if __name__ == "__main__":
    main()
