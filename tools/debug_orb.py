"""tools/debug_orb.py -- stage-by-stage comparison of the GPU ORB extraction with the oracle / cv2 (mismatch counts instead of asserts)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from orb_util import CASES, real_gray  # noqa: E402
from oracle import orb_oracle as O  # noqa: E402
from sfm_toy_library_b200 import capi  # noqa: E402

ctx = capi.Context(0)
cases = [("real", real_gray, 5000)] + CASES
for name, make, nf in cases:
    img = make()
    t = time.time(); k, d = ctx.orb_detect_and_compute(img, nf); dt = time.time() - t
    gray = O.to_gray(img) if img.ndim == 3 else img
    imgs = O.pyramid(gray)
    st = []
    for l in range(8):
        h, w = imgs[l].shape
        if h == 0:
            continue
        p = int((ctx.orb_download_level(0, 0, l, w, h) != imgs[l]).sum())
        f = int((ctx.orb_download_level(2, 0, l, w, h).astype(np.int32) != O.fast_score_map(imgs[l])).sum())
        b = int((ctx.orb_download_level(1, 0, l, w, h) != O.gaussian_blur_orb(imgs[l])).sum())
        st.append((p, f, b))
    rk, rd = O.cv2_detect_and_compute(img, nf)
    msg = "n %d vs %d" % (len(k), len(rk))
    if len(k) == len(rk):
        msg += " kp field mismatches %s desc rows %d" % ([int((k[:, c] != rk[:, c]).sum()) for c in range(6)], int((d != rd).any(1).sum()))
    print("%-24s %.1f ms  (pyr, fast, blur) mismatches per level %s  %s" % (name, dt * 1e3, st, msg), flush=True)
ctx.close()
