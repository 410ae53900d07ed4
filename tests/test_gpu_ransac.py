"""batched RANSAC hypothesis scoring (sfmb200_ransac_score) for the reference's three RANSAC stages
(SfMStereoUtilities.cpp:51-72 findHomographyInliers, :74-118 findCameraMatricesFromMatch, :208-243 findCameraPoseFrom2D3DMatch).
Per hypothesis the inlier counts / masks are EXACTLY those of the restatement of OpenCV's computeError + findInliers
(oracle/ransac_oracle.py, pinned to cv2 in tests/test_oracle_ransac.py); the RANSAC outcome is compared with cv2 statistically."""
import numpy as np
import pytest

from cfg1_util import Cfg1
from oracle import ransac_oracle as ro
from sfm_toy_library_b200 import capi, ransac, runsfm, stages

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
K = np.array([[2500, 0, 512], [0, 2500, 384], [0, 0, 1]], np.float32)


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def cfg1():
    return Cfg1()


def _pair(cfg1, p):
    i, j = cfg1.pairs[p]; q, t, _ = cfg1.matches[p]
    return cfg1.features[i].points[q], cfg1.features[j].points[t]


@pytest.mark.parametrize("p", [0, 7, 20])
def test_homography_scores_exact(ctx, cfg1, p):
    a, b = _pair(cfg1, p)
    hyps = ransac.homography_hypotheses(a, b, 300, np.random.RandomState(p))
    counts, best, mask = ctx.ransac_score(capi.MODEL_HOMOGRAPHY, a, b, hyps, None, 10.0)
    oc, ob, om = ro.score(0, a, b, hyps, None, 10.0)
    np.testing.assert_array_equal(counts, oc); assert best == ob; np.testing.assert_array_equal(mask, om)
    # statistical: the best of 300 seeded hypotheses reaches cv::findHomography's inlier count to within 15 %
    _, cvmask = cv2.findHomography(a, b, cv2.RANSAC, 10.0)
    assert counts[best] >= 0.85 * int(cvmask.sum())


@pytest.mark.parametrize("p", [0, 7, 20])
def test_essential_scores_exact(ctx, cfg1, p):
    a, b = _pair(cfg1, p)
    hyps = ransac.essential_hypotheses(a, b, 2500.0, (512.0, 384.0), 150, np.random.RandomState(p))
    assert len(hyps) > 150                                   # several solutions per sample
    aux = (2500.0, 512.0, 384.0)
    counts, best, mask = ctx.ransac_score(capi.MODEL_ESSENTIAL, a, b, hyps, aux, 1.0 / 2500.0)
    oc, ob, om = ro.score(1, a, b, hyps, aux, 1.0 / 2500.0)
    np.testing.assert_array_equal(counts, oc); assert best == ob; np.testing.assert_array_equal(mask, om)
    # cv2's own E scored here gives cv2's own mask (the pin of the restatement, now through the GPU)
    E, cvmask = cv2.findEssentialMat(a, b, 2500.0, (512.0, 384.0), cv2.RANSAC, 0.999, 1.0)
    c2, b2, m2 = ctx.ransac_score(capi.MODEL_ESSENTIAL, a, b, E.reshape(1, 9), aux, 1.0 / 2500.0)
    np.testing.assert_array_equal(m2, cvmask.reshape(-1))
    assert counts[best] >= 0.85 * int(cvmask.sum())


def test_pose_scores_exact_and_reference_unit_test_scene(ctx, cfg1):
    g = cfg1.g
    for k in range(cfg1.n_pnp):
        X = g[f"pnp_{k}_p3"]; uv = g[f"pnp_{k}_p2"]
        hyps = ransac.pose_hypotheses(X, uv, K, 100, np.random.RandomState(k))
        counts, best, mask = ctx.ransac_score(capi.MODEL_POSE, X, uv, hyps, K.reshape(-1), 10.0)
        oc, ob, om = ro.score(2, X, uv, hyps, K.astype(np.float64), 10.0)
        # the pose error goes through a division and a float cast: allow a correspondence sitting on the threshold
        assert np.abs(counts - oc).max() <= 1 and abs(int(counts[best]) - int(oc[ob])) <= 1
    # the reference's find_camera_pose_from_2d3d_match (SfMUnitTests.cpp:194-216): 12 exact correspondences, R to 0.01, t to 0.1
    from sfm_toy_library_b200 import synth
    Pl, _ = synth.fixture_poses()
    rvec, _ = cv2.Rodrigues(Pl[:, :3].copy())
    proj, _ = cv2.projectPoints(synth.CANNED_POINTS, rvec, Pl[:, 3].copy(), synth.TEST_K, None)
    ok, pose = ransac.findCameraPoseFrom2D3DMatch(stages.Intrinsics(synth.TEST_K), proj.reshape(-1, 2).astype(np.float32), synth.CANNED_POINTS, ctx=ctx)
    assert ok
    assert np.abs(pose[:, :3] - Pl[:, :3]).max() < 0.01 and np.abs(pose[:, 3] - Pl[:, 3]).max() < 0.1


def test_stage_functions_and_whole_replay(ctx, cfg1):
    """The three stage functions against cv2's (statistical), then the whole runSfM replay with GPU-scored RANSAC."""
    intr = stages.Intrinsics(K.copy())
    m = np.zeros(len(cfg1.matches[0][0]), stages.DMATCH)
    m["queryIdx"], m["trainIdx"], m["distance"] = cfg1.matches[0]
    i, j = cfg1.pairs[0]
    n_gpu = ransac.findHomographyInliers(cfg1.features[i], cfg1.features[j], m, ctx=ctx)
    n_cv = runsfm.findHomographyInliers_cv2(cfg1.features[i], cfg1.features[j], m)
    assert 0.85 * n_cv <= n_gpu <= len(m)
    ok, pruned, Pl, Pr = ransac.findCameraMatricesFromMatch(intr, m, cfg1.features[i], cfg1.features[j], ctx=ctx)
    ok2, pruned2, Pl2, Pr2 = runsfm.findCameraMatricesFromMatch_cv2(intr, m, cfg1.features[i], cfg1.features[j])
    assert ok and len(pruned) >= 0.8 * len(pruned2)
    R1, R2 = Pr[:, :3].astype(np.float64), Pr2[:, :3].astype(np.float64)
    ang = np.degrees(np.arccos(np.clip((np.trace(R1 @ R2.T) - 1) / 2, -1, 1)))
    assert ang < 3.0, ang                                    # same relative rotation to a few degrees (two-view geometry is only statistical)
    sfm = runsfm.SfM(cfg1.features, cfg1.size,
                     matchAllPairs=lambda f, pr: stages.matchAllPairs(f, pr, ctx=ctx), matchFeatures=lambda a, b: stages.matchFeatures(a, b, ctx=ctx),
                     triangulateViews=lambda *a: stages.triangulateViews(*a, ctx=ctx), adjustBundle=lambda *a: stages.adjustBundle(*a, ctx=ctx),
                     findHomographyInliers=lambda *a: ransac.findHomographyInliers(*a, ctx=ctx),
                     findCameraMatricesFromMatch=lambda *a: ransac.findCameraMatricesFromMatch(*a, ctx=ctx),
                     findCameraPoseFrom2D3DMatch=lambda *a: ransac.findCameraPoseFrom2D3DMatch(*a, ctx=ctx))
    sfm.runSfM()
    # statistical agreement with the cv2-RANSAC replay of the fixture (1430 points, focal refined from 2500 to 965.6 by the six
    # bundle adjustments): all 7 views registered, a cloud of the same order (the essential-matrix inlier sets differ by sample,
    # which changes how many matches survive into triangulation), the same refined focal length to 15 %
    assert len(sfm.mDoneViews) == 7
    assert 0.5 * len(cfg1.g["final_cloud"]) < len(sfm.mReconstructionCloud) < 2.0 * len(cfg1.g["final_cloud"])
    assert abs(float(sfm.mIntrinsics.K[0, 0]) - float(cfg1.g["final_K"][0, 0])) < 0.15 * float(cfg1.g["final_K"][0, 0])
