"""tests/golden/make_orb.py -- golden vectors for SURVEY.md section 8 row f-3 (ORB extraction), generated IN THIS CONTAINER with the cv2
4.13 binding, i.e. with the very OpenCV call the reference makes (`ORB::create(5000)->detectAndCompute`, SfM2DFeatureUtilities.cpp:39, 48).

    python tests/golden/make_orb.py        (needs /root/reference/dataset/crazyhorse)

Writes
  orb_crazyhorse_gray0.png   the first crazyhorse image (sorted order) after cvtColor(BGR2GRAY), lossless -- the only reference image
                             that travels with the repo (414 KB); the BGR original stays in /root/reference
  orb_golden.npz             cv2's key points [n, 6] (x, y, size, angle, response, octave) + descriptors for that image (nfeatures 5000
                             and 1000) and for a 701 x 511 crop, a 96 x 96 BGR patch with its cv2 grey conversion, cv2's INTER_LINEAR_EXACT
                             pyramid level 1..7 checksums, the FAST detections of level 0 and cv2.sepFilter2D of level 0 with the float
                             Gaussian (what ORB's blur computes), the float kernel itself, and fastAtan2 samples.
"""
import glob
import os
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
files = sorted(glob.glob("/root/reference/dataset/crazyhorse/*.JPG"))
bgr = cv2.imread(files[0])
gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
cv2.imwrite(os.path.join(HERE, "orb_crazyhorse_gray0.png"), gray, [cv2.IMWRITE_PNG_COMPRESSION, 9])


def orb(img, nf):
    k, d = cv2.ORB_create(nf).detectAndCompute(img, None)
    return np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in k], np.float32).reshape(-1, 6), d


out = {"cv2_version": cv2.__version__, "source": os.path.basename(files[0])}
out["kp_5000"], out["desc_5000"] = orb(bgr, 5000)          # from the BGR image, as the reference calls it
kg, dg = orb(gray, 5000)
assert np.array_equal(kg, out["kp_5000"]) and np.array_equal(dg, out["desc_5000"])      # grey input gives the same answer
out["kp_1000"], out["desc_1000"] = orb(gray, 1000)
crop = gray[100:611, 200:901]
out["crop_rect"] = np.array([100, 611, 200, 901], np.int32)
out["kp_crop"], out["desc_crop"] = orb(np.ascontiguousarray(crop), 5000)
out["bgr_patch"] = bgr[300:396, 400:496].copy(); out["gray_patch"] = gray[300:396, 400:496].copy()
prev = gray; crcs = []
for l in range(1, 8):
    s = np.float32(np.power(np.float64(np.float32(1.2)), l))
    w = int(np.rint(np.float32(gray.shape[1]) / s)); h = int(np.rint(np.float32(gray.shape[0]) / s))
    prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR_EXACT)
    crcs.append((w, h, zlib.crc32(prev.tobytes())))
out["pyramid_crc"] = np.array(crcs, np.int64)
fk = cv2.FastFeatureDetector_create(20, True).detect(gray)
out["fast0"] = np.array([(p.pt[0], p.pt[1], p.response) for p in fk], np.int32)
k32 = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
out["gauss_kernel"] = k32
out["blur0_crc"] = np.int64(zlib.crc32(cv2.sepFilter2D(gray, cv2.CV_8U, k32, k32, borderType=cv2.BORDER_REFLECT_101).tobytes()))
rng = np.random.RandomState(0)
yx = rng.randint(-300000, 300000, (4000, 2)).astype(np.float32); yx[:50] = rng.randint(-3, 4, (50, 2))
out["atan_yx"] = yx; out["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
np.savez_compressed(os.path.join(HERE, "orb_golden.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
