"""Pins oracle/ransac_oracle.py (how OpenCV's RANSAC scores one hypothesis; SfMStereoUtilities.cpp:51-118, 208-243) to cv2:
the inlier mask cv2.findEssentialMat returns for the E it returns, cv2.projectPoints for the pose error, cv2.perspectiveTransform
for the homography transfer error -- on the real crazyhorse matches of tests/golden/cfg1_crazyhorse.npz."""
import numpy as np
import pytest

from cfg1_util import Cfg1
from oracle import ransac_oracle as ro

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def cfg1():
    return Cfg1()


def _pair_points(cfg1, p):
    i, j = cfg1.pairs[p]
    q, t, _ = cfg1.matches[p]
    return cfg1.features[i].points[q], cfg1.features[j].points[t]


@pytest.mark.parametrize("p", [0, 6, 11, 20])
def test_essential_mask_equals_cv2(cfg1, p):
    """cv::findEssentialMat(RANSAC) returns the best hypothesis un-refined together with its inlier mask: the restated Sampson
    error + threshold rule reproduces that mask bit for bit."""
    a, b = _pair_points(cfg1, p)
    focal, pp = 2500.0, (512.0, 384.0)
    E, mask = cv2.findEssentialMat(a, b, focal, pp, cv2.RANSAC, 0.999, 1.0)
    err = ro.err_essential(E, a, b, focal, pp[0], pp[1])
    got = err <= ro.thresh2(1.0 / focal)
    np.testing.assert_array_equal(got.astype(np.uint8), mask.reshape(-1))
    assert 20 < got.sum() < len(a)


def test_pose_error_equals_cv2_projectpoints(cfg1):
    g = cfg1.g
    for k in range(cfg1.n_pnp):
        X = g[f"pnp_{k}_p3"]; uv = g[f"pnp_{k}_p2"]; pose = g[f"pnp_{k}_pose"].astype(np.float64)
        K = np.array([[2500, 0, 512], [0, 2500, 384], [0, 0, 1]], np.float64)
        rvec, _ = cv2.Rodrigues(pose[:, :3])
        proj, _ = cv2.projectPoints(X.reshape(-1, 1, 3).astype(np.float64), rvec, pose[:, 3], K, None)
        d = uv - proj.reshape(-1, 2).astype(np.float32)
        ref = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
        err = ro.err_pose(pose, K, X, uv)
        np.testing.assert_allclose(err, ref, rtol=2e-3, atol=1e-3)            # Rodrigues round trip of R costs ~1e-7 relative
        t2 = ro.thresh2(10.0)
        away = np.abs(ref - t2) > 0.05
        np.testing.assert_array_equal((err <= t2)[away], (ref <= t2)[away])


@pytest.mark.parametrize("p", [0, 11])
def test_homography_error_equals_cv2_transfer(cfg1, p):
    a, b = _pair_points(cfg1, p)
    H, mask = cv2.findHomography(a, b, cv2.RANSAC, 10.0)
    proj = cv2.perspectiveTransform(a.reshape(-1, 1, 2), H).reshape(-1, 2)
    d = proj - b
    ref = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
    err = ro.err_homography(H, a, b)
    np.testing.assert_allclose(err, ref, rtol=1e-3, atol=1e-2)
    t2 = ro.thresh2(10.0)
    away = np.abs(ref - t2) > 1.0
    np.testing.assert_array_equal((err <= t2)[away], (ref <= t2)[away])
