"""The C++ drop-in shim (sfm-toy-library_b200/host/shim.cpp: matchFeatures / triangulateViews / adjustBundle with the
reference's signatures over the C ABI) exercised by its C++ test program, which replays the reference's
triangulate_from_2_views scene (SfMUnitTests.cpp:221-251)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_shim_program():
    exe = os.path.join(ROOT, "sfm-toy-library_b200", "host", "build", "test_shim")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "SHIM_TEST PASS" in r.stdout


def test_cpp_shim_extract_features_equals_cv2(tmp_path):
    """SfM2DFeatureUtilities::extractFeatures of the C++ shim (reference signature, cv::Mat in, Features out) against cv2's
    ORB_create(5000).detectAndCompute -- the call it replaces (SfM2DFeatureUtilities.cpp:39, 46-51): key points and descriptors bit for bit."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from orb_util import real_gray, textured
    from oracle import orb_oracle as O
    exe = os.path.join(ROOT, "sfm-toy-library_b200", "host", "build", "test_shim")
    for name, img in (("real", real_gray()), ("bgr", np.stack([textured(300, 400, 90 + c) for c in range(3)], 2))):
        raw = tmp_path / (name + ".raw"); out = tmp_path / (name + ".bin")
        np.ascontiguousarray(img).tofile(raw)
        h, w = img.shape[:2]; ch = 1 if img.ndim == 2 else 3
        r = subprocess.run([exe, "--orb", str(raw), str(w), str(h), str(ch), str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        blob = open(out, "rb").read()
        n = int(np.frombuffer(blob[:4], np.int32)[0])
        rec = np.frombuffer(blob[4:4 + 28 * n], np.dtype([("f", np.float32, 5), ("i", np.int32, 2)]))
        desc = np.frombuffer(blob[4 + 28 * n:], np.uint8).reshape(n, 32)
        rk, rd = O.cv2_detect_and_compute(img, 5000)
        assert n == len(rk) and np.array_equal(rec["f"], rk[:, :5]) and np.array_equal(rec["i"][:, 0], rk[:, 5].astype(np.int32))
        assert np.all(rec["i"][:, 1] == -1) and np.array_equal(desc, rd)
