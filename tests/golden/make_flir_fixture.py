"""Builds tests/golden/flir_0000{20,40,60,61,80}_{left,right}.jpg: the reference's five bundled FLIR pairs (BASELINE config
C1 = pair 000020; config C4 adds the other four),
rectified exactly as its only caller does (src/stereo_Yin.cpp:122-147: left = ...42.jpg, right = ...39.jpg,
stereoRectify(CALIB_ZERO_DISPARITY, alpha 0) + initUndistortRectifyMap(CV_16SC2) + remap(INTER_LINEAR) with the
calibration of cam_stereo_pheno.yml), stored as JPEG (quality 90) to keep the fixture small.  The GPU box has no
/root/reference, so the parity test reads these two files; it compares GPU and oracle on the SAME decoded pixels, so
the recompression does not matter.  Also stores Q (for the disparity -> 3-D step) in flir_000020_Q.npy.

Run here (needs /root/reference and cv2):  python tests/golden/make_flir_fixture.py"""
import os
import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

fs = cv2.FileStorage(os.path.join(REF, "cam_stereo_pheno.yml"), cv2.FILE_STORAGE_READ)
M1, D1, M2, D2, R, T = (fs.getNode(k).mat() for k in ("M1", "D1", "M2", "D2", "R", "T"))
PAIRS = ("000020", "000040", "000060", "000061", "000080")
size = (2048, 1536)
R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(M1, D1, M2, D2, size, R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0, newImageSize=size)
m11, m12 = cv2.initUndistortRectifyMap(M1, D1, R1, P1, size, cv2.CV_16SC2)
m21, m22 = cv2.initUndistortRectifyMap(M2, D2, R2, P2, size, cv2.CV_16SC2)
for tag in PAIRS:
    left = cv2.imread(os.path.join(REF, "build", tag + "_191400042.jpg"))
    right = cv2.imread(os.path.join(REF, "build", tag + "_191400039.jpg"))
    assert (left.shape[1], left.shape[0]) == size
    l = cv2.remap(left, m11, m12, cv2.INTER_LINEAR)
    r = cv2.remap(right, m21, m22, cv2.INTER_LINEAR)
    for side, img in (("left", l), ("right", r)):
        path = os.path.join(HERE, "flir_%s_%s.jpg" % (tag, side))
        if tag == "000020" and os.path.exists(path):
            continue  # pair 000020 is pinned by round-1 tests: keep the committed bytes
        cv2.imwrite(path, img, [cv2.IMWRITE_JPEG_QUALITY, 90])
np.save(os.path.join(HERE, "flir_000020_Q.npy"), Q)
print(l.shape, Q)
