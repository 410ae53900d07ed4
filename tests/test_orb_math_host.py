"""The arithmetic of the ORB kernels (csrc/orb_math.cuh, __host__ __device__) compiled for the CPU (tests/host_orb_math.cpp walks
images the way the kernels of csrc/orb.cu do) and compared with the oracle (oracle/orb_oracle.py, itself pinned to cv2): every stage
bit for bit.  Also the host-side pieces the C ABI exports (pyramid layout, resize taps, retainBest) and the generated pattern table."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from orb_util import blobs, noise, real_gray, textured

from oracle import orb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(ROOT, "tests", "_build", "libhost_orb_math.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "host_orb_math.cpp"), "-o", so],
                   check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


IMAGES = [("real", real_gray), ("noise", lambda: noise(300, 400, 21)), ("blobs", lambda: blobs(333, 517, 22)), ("textured", lambda: textured(240, 320, 23))]


@pytest.mark.parametrize("name,make", IMAGES, ids=[c[0] for c in IMAGES])
def test_pixel_stages(hm, name, make):
    g = np.ascontiguousarray(make()); h, w = g.shape
    # FAST score map
    s = np.zeros((h, w), np.uint8); hm.orbh_fast_score_map(_p(g), w, h, 20, _p(s))
    assert np.array_equal(s.astype(np.int32), O.fast_score_map(g))
    # float Gaussian
    b = np.zeros((h, w), np.uint8); hm.orbh_blur(_p(g), w, h, _p(b))
    assert np.array_equal(b, O.gaussian_blur_orb(g))
    # one pyramid step through the exported tap tables
    from sfm_toy_library_b200 import capi
    dw, dh = int(np.rint(np.float32(w) / np.float32(1.2))), int(np.rint(np.float32(h) / np.float32(1.2)))
    x0, x1, ax = capi.orb_linear_exact_taps(w, dw); y0, y1, ay = capi.orb_linear_exact_taps(h, dh)
    r = np.zeros((dh, dw), np.uint8)
    hm.orbh_resize(_p(g), w, _p(x0), _p(x1), _p(ax), _p(y0), _p(y1), _p(ay), dw, dh, _p(r))
    assert np.array_equal(r, O.resize_linear_exact(g, dw, dh))
    # Harris responses, angles, descriptors at the FAST corners away from the border
    xs, ys, _ = O.fast_detect(g)
    m = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
    xs, ys = np.ascontiguousarray(xs[m]), np.ascontiguousarray(ys[m]); n = len(xs)
    assert n > 50
    hr = np.zeros(n, np.float32); hm.orbh_harris(_p(g), w, _p(xs), _p(ys), n, _p(hr))
    assert np.array_equal(hr, O.harris_responses(g, xs, ys))
    an = np.zeros(n, np.float32); hm.orbh_angles(_p(g), w, _p(xs), _p(ys), n, _p(an))
    assert np.array_equal(an, O.ic_angles(g, xs, ys))
    for scale in (np.float32(1.0), O.level_scales()[3]):
        inv = np.float32(1.0) / scale
        pt = np.zeros((n, 2), np.float32); d = np.zeros((n, 32), np.uint8)
        hm.orbh_describe(_p(b), w, _p(xs), _p(ys), _p(an), n, C.c_float(scale), C.c_float(inv), _p(pt), _p(d))
        assert np.array_equal(pt, np.stack([xs.astype(np.float32) * scale, ys.astype(np.float32) * scale], 1))
        cx = np.rint(pt[:, 0] * inv).astype(np.int64); cy = np.rint(pt[:, 1] * inv).astype(np.int64)
        assert np.array_equal(d, O.brief_descriptors(b, cx, cy, an))


def test_gray_and_atan2(hm, golden):
    gold = golden("orb_golden.npz")
    bgr = np.ascontiguousarray(gold["bgr_patch"]); out = np.zeros(bgr.shape[:2], np.uint8)
    hm.orbh_gray(_p(bgr), out.size, _p(out))
    assert np.array_equal(out, gold["gray_patch"])
    rnd = noise(64, 64, 5, 3); out = np.zeros((64, 64), np.uint8); hm.orbh_gray(_p(rnd), out.size, _p(out))
    assert np.array_equal(out, O.to_gray(rnd))
    yx = gold["atan_yx"]; y = np.ascontiguousarray(yx[:, 0]); x = np.ascontiguousarray(yx[:, 1]); a = np.zeros(len(y), np.float32)
    hm.orbh_atan2(_p(y), _p(x), len(y), _p(a))
    assert np.array_equal(a, gold["atan_deg"])


def test_pattern_header_is_the_table_of_the_cv2_binary(hm):
    hm.orbh_pattern.restype = C.POINTER(C.c_byte * 1024)
    tab = np.frombuffer(hm.orbh_pattern().contents, np.int8).reshape(256, 4)
    assert np.array_equal(tab.astype(np.int32), O.bit_pattern())


def test_exported_host_pieces_match_the_oracle():
    from sfm_toy_library_b200 import capi
    for (w, h, nf) in [(1024, 768, 5000), (517, 333, 500), (100, 90, 100), (64, 64, 100), (4000, 3000, 20000), (9, 8, 5)]:
        lw, lh, s, q = capi.orb_layout(w, h, nf)
        assert [(a, b) for a, b in zip(lw, lh)] == [x if x[0] > 0 and x[1] > 0 else (0, 0) for x in O.level_sizes(w, h)]
        assert np.array_equal(s, np.array(O.level_scales(), np.float32)) and list(q) == O.features_per_level(nf)
    for src, dst in [(1024, 853), (768, 640), (853, 711), (100, 83), (7, 6), (2, 1), (5, 5)]:
        for a, b in zip(capi.orb_linear_exact_taps(src, dst), O.linear_exact_coefficients(src, dst)):
            assert np.array_equal(a, b)
    rng = np.random.RandomState(0)
    for n, k in [(1000, 100), (5000, 4999), (10, 20), (300, 0), (20000, 2172), (0, 5)]:
        r = rng.randint(20, 60, n).astype(np.float32)          # integer-valued like FAST scores: many ties
        assert np.array_equal(capi.orb_retain_best(r, k), O.retain_best(r, k))
        r = rng.rand(n).astype(np.float32)
        assert np.array_equal(capi.orb_retain_best(r, k), O.retain_best(r, k))


def test_host_thread_pool():
    """csrc/host_pool.h: the pool the ORB stage stages images and runs its retainBest tasks on (many short loops, late-waking workers)."""
    import ctypes as C
    from sfm_toy_library_b200 import capi
    f = capi.lib().sfmb200_host_pool_selftest; f.restype = C.c_int64
    want = None
    for threads in (1, 2, 8, 16):
        got = f(threads, 3000, 64)
        assert got >= 0
        want = got if want is None else want
        assert got == want
    assert f(4, 200, 5000) > 0 and f(4, 10, 0) == 0 and f(0, 1, 1) == -1


def test_retain_best_properties():
    """KeyPointsFilter::retainBest through the C ABI: survivors are exactly the responses >= the n-th largest (ties kept), every index
    appears once, and the order is the oracle's (libstdc++ nth_element + partition) for tie-heavy and tie-free inputs alike."""
    from sfm_toy_library_b200 import capi
    rng = np.random.RandomState(11)
    for trial in range(60):
        n = int(rng.randint(1, 4000)); k = int(rng.randint(0, n + 50))
        r = (rng.randint(0, 1 + rng.randint(1, 50), n) if trial % 2 else rng.rand(n)).astype(np.float32)
        o = capi.orb_retain_best(r, k)
        assert np.array_equal(o, O.retain_best(r, k))
        assert len(set(o.tolist())) == len(o)
        if k >= n:
            assert np.array_equal(o, np.arange(n))
        elif k == 0:
            assert len(o) == 0
        else:
            thr = np.sort(r)[::-1][k - 1]
            assert set(o.tolist()) == set(np.nonzero(r >= thr)[0].tolist())
