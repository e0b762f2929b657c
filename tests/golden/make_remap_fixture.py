"""Builds tests/golden/remap_small.npz: golden vectors for the rectification remap, produced by OpenCV itself.

The maps come from the reference's own calibration (cam_stereo_pheno.yml, rectified exactly as its caller does,
src/stereo_Yin.cpp:135-140: stereoRectify(CALIB_ZERO_DISPARITY, alpha 0) + initUndistortRectifyMap(CV_16SC2)); to keep
the fixture small only two 96x64 windows of the left/right maps are kept (one at the image corner, so that the
footprints leave the source), shifted so that they index a 160x112 random source image.  Expected outputs =
cv2.remap(src, map1, map2, INTER_LINEAR) (BORDER_CONSTANT 0, OpenCV's default, what the reference calls).

Run here (needs /root/reference and cv2):  python tests/golden/make_remap_fixture.py"""
import os
import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

fs = cv2.FileStorage(os.path.join(REF, "cam_stereo_pheno.yml"), cv2.FILE_STORAGE_READ)
M1, D1, M2, D2, R, T = (fs.getNode(k).mat() for k in ("M1", "D1", "M2", "D2", "R", "T"))
size = (2048, 1536)
R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(M1, D1, M2, D2, size, R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0, newImageSize=size)
maps = [cv2.initUndistortRectifyMap(M1, D1, R1, P1, size, cv2.CV_16SC2), cv2.initUndistortRectifyMap(M2, D2, R2, P2, size, cv2.CV_16SC2)]
rng = np.random.default_rng(20261018)
out = {}
for i, (mxy, mf) in enumerate(maps):
    for j, (x0, y0) in enumerate(((0, 0), (900, 700))):
        wxy = mxy[y0:y0 + 64, x0:x0 + 96].astype(np.int32)
        wf = mf[y0:y0 + 64, x0:x0 + 96].copy()
        # shift the window's source coordinates into a small image; the corner window keeps some footprints outside
        off = wxy.reshape(-1, 2).min(0) - (np.array([0, 0]) if j else np.array([-3, -3]))
        wxy = (wxy - off + (0 if j else -6)).astype(np.int16)
        src = rng.integers(0, 256, (112, 160, 3), dtype=np.uint8)
        exp = cv2.remap(src, wxy, wf, cv2.INTER_LINEAR)
        out[f"src_{i}{j}"] = src
        out[f"xy_{i}{j}"] = wxy
        out[f"fxy_{i}{j}"] = wf
        out[f"exp_{i}{j}"] = exp
        print(i, j, "outside-footprint pixels:", int(((wxy[..., 0] < -1) | (wxy[..., 1] < -1) | (wxy[..., 0] >= 160) | (wxy[..., 1] >= 112)).sum()),
              "distinct fractions:", len(np.unique(wf)))
np.savez_compressed(os.path.join(HERE, "remap_small.npz"), **out)
