"""oracle/orb_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy) of what `mDetector->detectAndCompute(image, noArray(), keyPoints, descriptors)` computes for
`mDetector = ORB::create(5000)` (reference: SfM2DFeatureUtilities.cpp:39, 46-51).

The arithmetic lives in an un-vendored dependency, OpenCV (>= 3.1 per CMakeLists.txt:28; this image has cv2 4.13.0): the
restatement follows OpenCV's published ORB (Rublee et al. 2011; features2d/orb.cpp, fast.cpp, fast_score.cpp, keypoint.cpp;
imgproc resize / filter / color / mathfuncs) and is PINNED against the cv2 binary of this image stage by stage and end to end
(tests/test_oracle_orb.py): grey conversion, INTER_LINEAR_EXACT pyramid, FAST-9/16 + non-maximum suppression in raster order,
both retainBest selections (order included), Harris responses, intensity-centroid angles, the float separable Gaussian that
cv::GaussianBlur runs on a sub-matrix, and the 256-bit steered BRIEF descriptors -- bit for bit on real and synthetic images.
Parameters are ORB::create's defaults: scaleFactor 1.2f, nlevels 8, edgeThreshold 31, firstLevel 0, WTA_K 2, HARRIS_SCORE,
patchSize 31, fastThreshold 20.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_select.so")
f32 = np.float32
f64 = np.float64

NLEVELS, EDGE, PATCH, HALF_PATCH, FAST_T, HARRIS_BLOCK = 8, 31, 31, 15, 20, 7
SCALE_FACTOR = float(f32(1.2))          # ORB::create(int, float scaleFactor = 1.2f, ...) stored in a double
# FAST circle of radius 3, OpenCV's order (fast_score.cpp makeOffsets)
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0),
          (-3, 1), (-2, 2), (-1, 3)]


def build(force=False):
    src = os.path.join(_HERE, "orb_select.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", _SO, src], check=True, capture_output=True)
    return _SO


_sel = None


def retain_best(response, n_points):
    """KeyPointsFilter::retainBest: indices of the survivors in the order std::nth_element + std::partition leave them."""
    global _sel
    if _sel is None:
        _sel = ctypes.CDLL(build())
    r = np.ascontiguousarray(response, np.float32)
    order = np.zeros(max(len(r), 1), np.int32)
    n = _sel.orb_oracle_retain_best(r.ctypes.data_as(ctypes.c_void_p), len(r), int(n_points), order.ctypes.data_as(ctypes.c_void_p))
    return order[:n].copy()


_pattern = None


def bit_pattern():
    """The learned 256 x (x0, y0, x1, y1) test pattern of rBRIEF (orb.cpp `bit_pattern_31_`), read from the cv2 binary of this
    image: it is data of the pinned dependency, not something to restate."""
    global _pattern
    if _pattern is None:
        import cv2
        so = [f for f in os.listdir(os.path.dirname(cv2.__file__)) if f.startswith("cv2") and f.endswith(".so")][0]
        blob = open(os.path.join(os.path.dirname(cv2.__file__), so), "rb").read()
        key = np.array([8, -3, 9, 5, 4, 2, 7, -12, -11, 9, -8, 2], "<i4").tobytes()
        at = blob.find(key)
        assert at >= 0 and blob.find(key, at + 1) < 0
        _pattern = np.frombuffer(blob[at:at + 4096], "<i4").reshape(256, 4).copy()
        assert np.abs(_pattern).max() <= 13 and tuple(_pattern[-1]) == (-1, -6, 0, -11)
    return _pattern


# ------------------------------------------------------------------------------------------------ pyramid
def to_gray(bgr):
    """cvtColor(COLOR_BGR2GRAY) for 8-bit input: fixed point, 15 fractional bits (imgproc color_rgb: R2Y 9798, G2Y 19235, B2Y 3735)."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def level_scales(nlevels=NLEVELS):
    return [f32(np.power(f64(SCALE_FACTOR), f64(l))) for l in range(nlevels)]          # (float)std::pow(scaleFactor, level)


def level_sizes(w, h, nlevels=NLEVELS):
    return [(int(np.rint(f32(w) / s)), int(np.rint(f32(h) / s))) for s in level_scales(nlevels)]   # cvRound(cols / scale)


def features_per_level(nfeatures, nlevels=NLEVELS):
    factor = f32(1.0 / SCALE_FACTOR)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(np.power(f64(factor), f64(nlevels))))
    out, total = [], 0
    for _ in range(nlevels - 1):
        n = int(np.rint(nd)); out.append(n); total += n; nd = nd * factor
    out.append(max(nfeatures - total, 0))
    return out


def linear_exact_coefficients(src, dst):
    """resize(..., INTER_LINEAR_EXACT) for 8-bit data: source index pair and the 8-bit weight of the right/lower tap."""
    f = (f64(src) / f64(dst)) * (np.arange(dst, dtype=f64) + 0.5) - 0.5
    i = np.floor(f).astype(np.int64)
    a = np.rint((f - i) * 256).astype(np.int64)            # ufixedpoint16(softdouble): cvRound, ties to even
    lo, hi = i < 0, i >= src - 1
    i0 = np.where(hi, src - 1, np.clip(i, 0, src - 1)); i1 = np.where(hi, src - 1, np.clip(i + 1, 0, src - 1))
    a = np.where(lo | hi, 0, a)
    return i0, i1, a


def resize_linear_exact(S, w, h):
    x0, x1, ax = linear_exact_coefficients(S.shape[1], w); y0, y1, ay = linear_exact_coefficients(S.shape[0], h)
    S = S.astype(np.int64)
    H = (256 - ax)[None, :] * S[:, x0] + ax[None, :] * S[:, x1]                          # horizontal pass, 8.8 fixed point
    return (((256 - ay)[:, None] * H[y0, :] + ay[:, None] * H[y1, :] + 32768) >> 16).astype(np.uint8)


def pyramid(gray, nlevels=NLEVELS):
    """Level l is resized from level l-1 (orb.cpp: `resize(prevImg, currImg, sz, 0, 0, INTER_LINEAR_EXACT)`)."""
    sizes = level_sizes(gray.shape[1], gray.shape[0], nlevels)
    imgs = [gray]
    for l in range(1, nlevels):
        w, h = sizes[l]
        imgs.append(resize_linear_exact(imgs[-1], w, h) if w > 0 and h > 0 else np.zeros((max(h, 0), max(w, 0)), np.uint8))
    return imgs


# ------------------------------------------------------------------------------------------------ FAST
def fast_score_map(I, threshold=FAST_T):
    """FAST-9/16: cornerScore<16> (fast_score.cpp) = max over the 16 arcs of 9 of min(centre - p) (dark arc) or min(p - centre)
    (bright arc), minus 1; a pixel is a corner iff that is >= threshold (fast.cpp: 9 contiguous pixels beyond +-threshold,
    strictly).  Rows/columns within 3 pixels of the border are never corners.  0 where there is no corner."""
    h, w = I.shape
    S = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return S
    J = I.astype(np.int32)
    c = J[3:h - 3, 3:w - 3]
    d = np.stack([c - J[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in CIRCLE])
    d2 = np.concatenate([d, d[:9]])
    A = np.full(c.shape, -1000); B = np.full(c.shape, -1000)
    for k in range(16):
        A = np.maximum(A, d2[k:k + 9].min(0)); B = np.maximum(B, (-d2[k:k + 9]).min(0))
    s = np.maximum(A, B) - 1
    S[3:h - 3, 3:w - 3] = np.where(s >= threshold, s, 0)
    return S


def fast_nms(S):
    """fast.cpp non-maximum suppression: strictly greater than all 8 neighbours' scores (non-corners count as 0)."""
    h, w = S.shape
    P = np.pad(S, 1)
    keep = S > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= S > P[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    return keep


def fast_detect(I, threshold=FAST_T):
    """FastFeatureDetector(threshold, nonmaxSuppression=true, TYPE_9_16)::detect: x, y, score in raster order."""
    S = fast_score_map(I, threshold)
    ys, xs = np.nonzero(fast_nms(S))
    return xs.astype(np.int32), ys.astype(np.int32), S[ys, xs].astype(np.int32)


# ------------------------------------------------------------------------------------------------ responses, angles
def harris_responses(I, xs, ys, block=HARRIS_BLOCK, k=0.04):
    """orb.cpp HarrisResponses: integer Sobel-like sums over a block x block window, float expression evaluated left to right."""
    if len(xs) == 0:
        return np.zeros(0, np.float32)
    J = I.astype(np.int64); r = block // 2
    oy, ox = np.mgrid[-r:r + 1, -r:r + 1]
    Y = ys[:, None, None] + oy[None]; X = xs[:, None, None] + ox[None]
    p = lambda dy, dx: J[Y + dy, X + dx]
    Ix = (p(0, 1) - p(0, -1)) * 2 + (p(-1, 1) - p(-1, -1)) + (p(1, 1) - p(1, -1))
    Iy = (p(1, 0) - p(-1, 0)) * 2 + (p(1, -1) - p(-1, -1)) + (p(1, 1) - p(-1, 1))
    a = (Ix * Ix).sum((1, 2)).astype(f32); b = (Iy * Iy).sum((1, 2)).astype(f32); c = (Ix * Iy).sum((1, 2)).astype(f32)
    scale = f32(1.0) / (f32(4 * block) * f32(255.0)); ssq = scale * scale * scale * scale
    return (((a * b - c * c) - (f32(k) * (a + b)) * (a + b)) * ssq).astype(f32)


def umax_table():
    vmax = int(np.floor(f32(HALF_PATCH) * np.sqrt(f32(2.0)) / 2 + 1)); vmin = int(np.ceil(f32(HALF_PATCH) * np.sqrt(f32(2.0)) / 2))
    u = [0] * (HALF_PATCH + 2)
    for v in range(vmax + 1):
        u[v] = int(np.rint(np.sqrt(f64(HALF_PATCH * HALF_PATCH - v * v))))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while u[v0] == u[v0 + 1]:
            v0 += 1
        u[v] = v0; v0 += 1
    return u[:HALF_PATCH + 1]


def fast_atan2(y, x):
    """cv::fastAtan2 (mathfuncs_core: 7th-order odd polynomial in float, no fused multiply-add; the constants are float PRODUCTS
    `0.99978784f * (float)(180 / CV_PI)` ...)."""
    y = np.asarray(y, f32); x = np.asarray(x, f32)
    s = f32(180.0 / np.pi)
    p1, p3, p5, p7 = f32(0.9997878412794807) * s, f32(-0.3258083974640975) * s, f32(0.1555786518463281) * s, f32(-0.04432655554792128) * s
    ax, ay = np.abs(x), np.abs(y); eps = f32(2.220446049250313e-16)
    swap = ~(ax >= ay)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (np.where(swap, ax, ay) / (np.where(swap, ay, ax) + eps)).astype(f32)
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(swap, f32(90.0) - a, a)
    a = np.where(x < 0, f32(180.0) - a, a)
    a = np.where(y < 0, f32(360.0) - a, a)
    return a.astype(f32)


def ic_angles(I, xs, ys):
    """orb.cpp ICAngles: first-order moments over the circular patch of radius 15, angle = fastAtan2(m01, m10) in degrees."""
    if len(xs) == 0:
        return np.zeros(0, np.float32)
    J = I.astype(np.int64); um = umax_table()
    oy, ox = np.mgrid[-HALF_PATCH:HALF_PATCH + 1, -HALF_PATCH:HALF_PATCH + 1]
    mask = np.abs(ox) <= np.array(um)[np.abs(oy)]
    P = J[ys[:, None, None] + oy[None], xs[:, None, None] + ox[None]] * mask[None]
    m10 = (P * ox[None]).sum((1, 2)); m01 = (P * oy[None]).sum((1, 2))
    return fast_atan2(m01.astype(f32), m10.astype(f32))


# ------------------------------------------------------------------------------------------------ blur + descriptors
def gaussian_kernel_7_2():
    """getGaussianKernel(7, 2, CV_32F) (what createGaussianKernels hands to sepFilter2D for 8-bit images)."""
    x = np.arange(-3, 4, dtype=f64)
    k = np.exp(-(x * x) / 8.0)
    return (k / k.sum()).astype(f32)


def _fma(f, s, acc):                     # float fused multiply-add through double (product exact, one extra rounding < 2^-29)
    return (f64(f) * s.astype(f64) + acc.astype(f64)).astype(f32)


def gaussian_blur_orb(I):
    """What `GaussianBlur(workingMat, workingMat, Size(7, 7), 2, 2, BORDER_REFLECT_101)` does inside ORB: the level is a sub-matrix
    of the pyramid buffer, so cv::GaussianBlur skips its fixed-point 8-bit path and runs sepFilter2D with the FLOAT kernel:
    row filter u8 -> f32 as a left-to-right chain of fused multiply-adds (the AVX2 build of filter.simd.hpp), column filter as
    centre tap then symmetric pairs `s = fma(k[j], up + down, s)`, result cvRound-ed and saturated.  Pinned on whole images against
    cv2.sepFilter2D with that kernel."""
    k = gaussian_kernel_7_2()
    h, w = I.shape
    if h < 4 or w < 4:
        return I.copy()
    P = np.pad(I, 3, mode="reflect").astype(f32)
    R = (k[0] * P[:, 0:w]).astype(f32)
    for i in range(1, 7):
        R = _fma(k[i], P[:, i:i + w], R)
    C = (k[3] * R[3:3 + h]).astype(f32)
    for j in (1, 2, 3):
        C = _fma(k[3 + j], (R[3 + j:3 + j + h] + R[3 - j:3 - j + h]).astype(f32), C)
    return np.clip(np.rint(C), 0, 255).astype(np.uint8)


def brief_descriptors(B, cx, cy, angle_deg):
    """orb.cpp computeOrbDescriptors, WTA_K = 2: pattern rotated by the key point angle in float (no fused multiply-add),
    cvRound, bit = (I[p0] < I[p1]), least significant bit first."""
    n = len(cx)
    if n == 0:
        return np.zeros((0, 32), np.uint8)
    pat = bit_pattern().reshape(512, 2).astype(f32)
    ang = (np.asarray(angle_deg, f32) * f32(np.pi / 180.0)).astype(f32)
    a = np.cos(ang.astype(f64)).astype(f32)[:, None]; b = np.sin(ang.astype(f64)).astype(f32)[:, None]
    x = (pat[None, :, 0] * a).astype(f32) - (pat[None, :, 1] * b).astype(f32)
    y = (pat[None, :, 0] * b).astype(f32) + (pat[None, :, 1] * a).astype(f32)
    ix = np.rint(x).astype(np.int64); iy = np.rint(y).astype(np.int64)
    v = B[cy[:, None] + iy, cx[:, None] + ix].astype(np.int32)
    bits = (v[:, 0::2] < v[:, 1::2]).astype(np.uint8)
    return np.packbits(bits, axis=1, bitorder="little")


# ------------------------------------------------------------------------------------------------ whole call
def detect_and_compute(image, nfeatures=5000, return_stages=False):
    """ORB::detectAndCompute(image, noArray(), ...) with ORB::create(nfeatures) defaults.
    Returns key points [n, 6] float32 (x, y, size, angle, response, octave) in OpenCV's output order and descriptors [n, 32]."""
    gray = to_gray(image) if image.ndim == 3 else image
    h, w = gray.shape
    scales = level_scales(); imgs = pyramid(gray); per_level = features_per_level(nfeatures)
    kp = []; centres = []; stages = []
    for l in range(NLEVELS):
        I = imgs[l]; lh, lw = I.shape
        if lh <= 0 or lw <= 0:
            continue
        xs, ys, sc = fast_detect(I)
        m = (xs >= EDGE) & (xs < lw - EDGE) & (ys >= EDGE) & (ys < lh - EDGE)        # KeyPointsFilter::runByImageBorder
        xs, ys, sc = xs[m], ys[m], sc[m]
        o = retain_best(sc.astype(f32), 2 * per_level[l]); xs, ys = xs[o], ys[o]      # by FAST score, twice the quota
        hr = harris_responses(I, xs, ys)
        o = retain_best(hr, per_level[l]); xs, ys, hr = xs[o], ys[o], hr[o]            # by Harris response
        ang = ic_angles(I, xs, ys)
        s = scales[l]
        kp.append(np.stack([xs.astype(f32) * s, ys.astype(f32) * s, np.full(len(xs), f32(PATCH) * s, f32), ang, hr,
                            np.full(len(xs), l, f32)], 1).astype(f32))
        centres.append((l, xs, ys))
        stages.append(dict(level=l, n=len(xs)))
    kp = np.concatenate(kp) if kp else np.zeros((0, 6), f32)
    desc = []
    off = 0
    for l, xs, ys in centres:
        B = gaussian_blur_orb(imgs[l])
        k = kp[off:off + len(xs)]; off += len(xs)
        inv = f32(1.0) / scales[l]
        cx = np.rint(k[:, 0] * inv).astype(np.int64); cy = np.rint(k[:, 1] * inv).astype(np.int64)   # cvRound(pt * (1 / scale))
        desc.append(brief_descriptors(B, cx, cy, k[:, 3]))
    desc = np.concatenate(desc) if desc else np.zeros((0, 32), np.uint8)
    if return_stages:
        return kp, desc, dict(gray=gray, pyramid=imgs)
    return kp, desc


def cv2_detect_and_compute(image, nfeatures=5000):
    """The reference's own call through the cv2 binding (SfM2DFeatureUtilities.cpp:39, 48)."""
    import cv2
    k, d = cv2.ORB_create(nfeatures).detectAndCompute(image, None)
    kp = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in k], np.float32).reshape(-1, 6)
    return kp, (d if d is not None else np.zeros((0, 32), np.uint8))
