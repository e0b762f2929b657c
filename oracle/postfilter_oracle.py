"""oracle/postfilter_oracle.py — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's weightedMedianFilter (src/PatchMatchStereoGPU.cu:2436-2599), numpy + plain loops
(small cases only).  Per invalid pixel (mask == 1), :2469-2528:
  * window (2r+1)^2, row-major (i = dy outer, j = dx inner); an entry inside the image carries
    weight = exp(-sqrt(|dR|+|dG|+|dB|) * gamma) against the centre pixel and the neighbour's disparity, an entry outside
    carries weight 0 and disparity 0; weight_sum accumulates the in-image weights in window order (fp32);
  * insertion sort by disparity with '<' (:2496-2511) = stable sort;
  * running sum of weight[i] / weight_sum in sorted order (fp32); the first entry at which it reaches 0.5 gives the
    output disparity (:2514-2525); if it never does the pixel keeps its value.
Defined where the reference races (it filters in place): every pixel reads the input map.  The 766-entry weight table
is an input (the library's s3dmst_wmf_table: glibc expf(-sqrtf(i) * gamma)), so both sides use the same weights.
"""
import numpy as np


def weighted_median(disp, mask, bgr, tab, radius=10):
    H, W = disp.shape
    out = disp.copy()
    img = bgr.astype(np.int32)
    ws = 2 * radius + 1
    f32 = np.float32
    for y, x in zip(*np.nonzero(mask)):
        wts = np.zeros(ws * ws, f32)
        dsp = np.zeros(ws * ws, f32)
        wsum = f32(0.0)
        k = 0
        for dy in range(-radius, radius + 1):
            for dx in range(-radius, radius + 1):
                yy, xx = y + dy, x + dx
                if 0 <= xx < W and 0 <= yy < H:
                    w = tab[int(np.abs(img[yy, xx] - img[y, x]).sum())]
                    wts[k] = w
                    dsp[k] = disp[yy, xx]
                    wsum = f32(wsum + w)
                k += 1
        order = np.argsort(dsp, kind="stable")
        acc = f32(0.0)
        for i in order:
            acc = f32(acc + f32(wts[i] / wsum))
            if acc >= f32(0.5):
                out[y, x] = dsp[i]
                break
    return out
