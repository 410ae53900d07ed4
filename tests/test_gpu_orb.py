"""SURVEY.md section 8 row f-3 on the GPU: sfmb200_orb_detect_and_compute[_batch] (csrc/orb.cu) against OpenCV -- the reference's own call
`ORB::create(nfeatures)->detectAndCompute` (SfM2DFeatureUtilities.cpp:39, 48) through cv2 on the same box, the committed cv2 goldens of a
real crazyhorse image, and the oracle stage by stage.  Everything is compared BIT FOR BIT: key point order, pt, size, angle, response,
octave, descriptors; pyramid, FAST score map and blurred pyramid of every level."""
import threading

import numpy as np
import pytest

from orb_util import CASES, blobs, real_gray, textured

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sfm_toy_library_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gold(golden):
    return golden("orb_golden.npz")


def _same(got, kp, desc):
    k, d = got
    return k.shape[0] == kp.shape[0] and np.array_equal(k[:, :6], kp) and np.all(k[:, 6] == -1) and np.array_equal(d, desc)


def test_real_image_matches_cv2_golden(ctx, gold):
    g = real_gray()
    assert _same(ctx.orb_detect_and_compute(g, 5000), gold["kp_5000"], gold["desc_5000"])
    assert _same(ctx.orb_detect_and_compute(g, 1000), gold["kp_1000"], gold["desc_1000"])
    y0, y1, x0, x1 = gold["crop_rect"]
    assert _same(ctx.orb_detect_and_compute(np.ascontiguousarray(g[y0:y1, x0:x1]), 5000), gold["kp_crop"], gold["desc_crop"])
    bgr = np.stack([g, g, g], 2)                     # B = G = R converts to the same grey level
    assert _same(ctx.orb_detect_and_compute(bgr, 5000), gold["kp_5000"], gold["desc_5000"])


@pytest.mark.parametrize("name,make,nf", CASES, ids=[c[0] for c in CASES])
def test_equals_cv2(ctx, name, make, nf):
    from oracle import orb_oracle as O
    img = make()
    rk, rd = O.cv2_detect_and_compute(img, nf)
    assert _same(ctx.orb_detect_and_compute(img, nf), rk, rd), name


def test_stages_equal_oracle(ctx):
    from oracle import orb_oracle as O
    for img in (real_gray(), textured(333, 517, 31), blobs(200, 150, 6)):
        ctx.orb_detect_and_compute(img, 2000)
        imgs = O.pyramid(img)
        for l in range(8):
            h, w = imgs[l].shape
            assert np.array_equal(ctx.orb_download_level(0, 0, l, w, h), imgs[l]), ("pyramid", l)
            assert np.array_equal(ctx.orb_download_level(2, 0, l, w, h).astype(np.int32), O.fast_score_map(imgs[l])), ("fast", l)
            assert np.array_equal(ctx.orb_download_level(1, 0, l, w, h), O.gaussian_blur_orb(imgs[l])), ("blur", l)


def test_batch_equals_single_calls(ctx):
    imgs = [textured(480, 640, 40 + i) for i in range(5)] + [blobs(480, 640, 50), np.full((480, 640), 9, np.uint8)]
    single = [ctx.orb_detect_and_compute(im, 3000) for im in imgs]
    batch = ctx.orb_detect_and_compute(imgs, 3000)
    assert len(batch) == len(imgs)
    for (k0, d0), (k1, d1) in zip(single, batch):
        assert np.array_equal(k0, k1) and np.array_equal(d0, d1)
    assert len(batch[-1][0]) == 0 and len(batch[0][0]) > 1000


def test_small_capacity_reports_the_full_count(ctx):
    img = textured(480, 640, 60)
    full_k, full_d = ctx.orb_detect_and_compute(img, 2000)
    k, d = ctx.orb_detect_and_compute(img, 2000, capacity=100)          # the wrapper retries with the reported count
    assert np.array_equal(k, full_k) and np.array_equal(d, full_d)


def test_calls_from_several_threads(ctx):
    imgs = [textured(300, 400, 70 + i) for i in range(6)]
    want = [ctx.orb_detect_and_compute(im, 1000) for im in imgs]
    got = [None] * len(imgs)

    def work(i):
        got[i] = ctx.orb_detect_and_compute(imgs[i], 1000)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    [t.start() for t in ts]; [t.join() for t in ts]
    for (k0, d0), (k1, d1) in zip(want, got):
        assert np.array_equal(k0, k1) and np.array_equal(d0, d1)


def test_extract_then_match_equals_cv2_chain(ctx):
    """extractFeatures -> matchFeatures on the GPU = cv2 ORB -> cv2 knnMatch + ratio test (SfM.cpp:141-154, 157-212)."""
    import cv2
    from oracle import cv2_reference
    a = textured(480, 640, 80)
    M = np.float32([[0.98, 0.05, 6.0], [-0.04, 0.99, -3.0]])
    b = cv2.warpAffine(a, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    (ka, da), (kb, db) = ctx.orb_detect_and_compute([a, b], 5000)
    q, t, d = ctx.match_knn2_ratio(da, db)
    _, ra = cv2.ORB_create(5000).detectAndCompute(a, None); _, rb = cv2.ORB_create(5000).detectAndCompute(b, None)
    rq, rt, rd = cv2_reference.match_features(ra, rb)
    assert len(q) > 200 and np.array_equal(q, rq) and np.array_equal(t, rt) and np.array_equal(d, rd)


def test_bad_arguments(ctx):
    from sfm_toy_library_b200 import capi
    with pytest.raises(capi.SfmB200Error):
        ctx.orb_detect_and_compute(np.zeros((4, 4), np.uint8), 100)
    with pytest.raises(capi.SfmB200Error):
        ctx.orb_detect_and_compute(np.zeros((64, 64, 2), np.uint8), 100)
    with pytest.raises(capi.SfmB200Error):
        ctx.orb_detect_and_compute([np.zeros((64, 64), np.uint8), np.zeros((64, 65), np.uint8)], 100)
    assert len(ctx.orb_detect_and_compute(np.zeros((64, 64), np.uint8), 0)[0]) == 0
