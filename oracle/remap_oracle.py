"""oracle/remap_oracle.py — CPU restatement of cv::remap(src, dst, map1 CV_16SC2, map2 CV_16UC1, INTER_LINEAR,
BORDER_CONSTANT 0) for 8-bit 3-channel images: the rectification step of the reference's only caller
(src/stereo_Yin.cpp:139-144: initUndistortRectifyMap(..., CV_16SC2, ...) then remap(..., INTER_LINEAR)).

TEST INFRASTRUCTURE ONLY (tests/ and bench legs may import it; the product never does).

The arithmetic lives in a third-party dependency that is not under /root/reference: OpenCV imgproc (the reference
links 3.4.3; cv2 4.13 in this image implements the same fixed-point scheme).  Published algorithm restated here:
  * map1[y][x] = (sx, sy) integer source corner, map2[y][x] = fy * 32 + fx, fractions in 1/32 (INTER_BITS = 5);
  * weights: a 1024-entry table of four int16, round((1-fy)(1-fx) * 32768) ... with saturation to 32767 and the
    sum forced back to 32768 (imgwarp.cpp initInterTab2D);
  * value = (sum_i w_i * tap_i + 16384) >> 15 (FixedPtCast<int, uchar, 15>), taps outside the source = border (0);
  * a destination pixel whose 2x2 footprint lies completely outside the source gets the border value.
Pinned by tests/test_remap.py against cv2.remap itself (random maps incl. borders) and tests/golden/remap_small.npz.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS


def bilinear_tab() -> np.ndarray:
    """int16 [1024][4]: weights of (S[y][x], S[y][x+1], S[y+1][x], S[y+1][x+1]) for table index fy*32+fx."""
    scale = np.float32(1.0) / np.float32(INTER_TAB_SIZE)
    t1 = np.zeros((INTER_TAB_SIZE, 2), np.float32)
    for i in range(INTER_TAB_SIZE):
        x = np.float32(i) * scale
        t1[i, 0] = np.float32(1.0) - x
        t1[i, 1] = x
    # the table is one flat array that the fix-up loop below indexes past the current entry (entries not yet
    # written are still zero), exactly like the static array in imgwarp.cpp
    flat = np.zeros(INTER_TAB_SIZE * INTER_TAB_SIZE * 4 + 8, np.int32)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            base = (i * INTER_TAB_SIZE + j) * 4
            isum = 0
            for k1 in range(2):
                vy = t1[i, k1]
                for k2 in range(2):
                    v = np.float32(vy * t1[j, k2])
                    iv = int(np.rint(np.float64(v) * COEF_SCALE))
                    iv = max(-32768, min(32767, iv))  # saturate_cast<short>
                    flat[base + k1 * 2 + k2] = iv
                    isum += iv
            if isum != COEF_SCALE:
                diff = isum - COEF_SCALE
                mk = Mk = (1, 1)
                for k1 in range(1, 3):
                    for k2 in range(1, 3):
                        val = flat[base + k1 * 2 + k2]
                        if val < flat[base + mk[0] * 2 + mk[1]]:
                            mk = (k1, k2)
                        elif val > flat[base + Mk[0] * 2 + Mk[1]]:
                            Mk = (k1, k2)
                if diff < 0:
                    idx = base + Mk[0] * 2 + Mk[1]
                else:
                    idx = base + mk[0] * 2 + mk[1]
                flat[idx] = int(np.int16(flat[idx] - diff))
    return flat[: INTER_TAB_SIZE * INTER_TAB_SIZE * 4].reshape(-1, 4).astype(np.int16)


def remap_fixed(src: np.ndarray, map_xy: np.ndarray, map_fxy: np.ndarray, border: int = 0) -> np.ndarray:
    """src uint8 [Hs][Ws][C]; map_xy int16 [H][W][2] (x, y); map_fxy uint16 [H][W] -> uint8 [H][W][C]."""
    tab = bilinear_tab().astype(np.int64)
    Hs, Ws, C = src.shape
    sx = map_xy[..., 0].astype(np.int64)
    sy = map_xy[..., 1].astype(np.int64)
    w = tab[map_fxy.astype(np.int64) & (INTER_TAB_SIZE * INTER_TAB_SIZE - 1)]  # [H][W][4]
    acc = np.zeros(map_fxy.shape + (C,), np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < Hs) & (xx >= 0) & (xx < Ws)
        tap = np.where(ok[..., None], src[np.clip(yy, 0, Hs - 1), np.clip(xx, 0, Ws - 1)].astype(np.int64), border)
        acc += tap * w[..., k][..., None]
    out = (acc + (1 << (COEF_BITS - 1))) >> COEF_BITS
    out = np.clip(out, 0, 255)
    outside = (sx >= Ws) | (sx + 1 < 0) | (sy >= Hs) | (sy + 1 < 0)
    out[outside] = border
    return out.astype(np.uint8)
