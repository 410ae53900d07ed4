"""GPU parity: sfmb200_triangulate (CUDA, through the C ABI) vs cv2 golden vectors, the oracle restatement of
triangulateViews (SfMStereoUtilities.cpp:120-206) and the reference's own unit-test fixture (SfMUnitTests.cpp:221-251)."""
import numpy as np
import pytest

from sfm_toy_library_b200 import capi, stages, synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5        # 3D points, relative (float32 outputs)
ERR_TOL = 1e-3        # px: keep mask may differ only this close to the 10 px threshold


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _check(X, keep, Xr, kr, er):
    good = kr.astype(bool)
    rel = np.linalg.norm(X[good] - Xr[good], axis=1) / np.linalg.norm(Xr[good], axis=1)
    assert rel.max() < REL_TOL, rel.max()
    diff = keep != kr
    if diff.any():
        assert (np.abs(np.max(er[diff], axis=1) - 10.0) < ERR_TOL).all()


def test_reference_unit_test_fixture(ctx, golden):
    """triangulate_from_2_views, through the mirror of the reference's interface."""
    g = golden("triangulate_cv2.npz")
    left = stages.Features(points=g["fx_ptsL"]); right = stages.Features(points=g["fx_ptsR"])
    cloud = []
    ok = stages.triangulateViews(stages.Intrinsics(g["fx_K"]), stages.ImagePair(0, 1), stages.GetAlignedMatching(len(g["fx_ptsL"])),
                                 left, right, g["fx_Pl"], g["fx_Pr"], cloud, ctx=ctx)
    assert ok and len(cloud) == 12
    for i, p in enumerate(cloud):
        assert np.linalg.norm(p.p - g["fx_points3d"][i]) < 0.01             # the reference's tolerance
        assert p.originatingViews == {0: i, 1: i}
    X = np.array([p.p for p in cloud])
    assert np.abs(X - g["fx_X"]).max() < 1e-4                                # and cv2's own output


@pytest.mark.parametrize("pre,with_matches", [("sy", True), ("ch", False)])
def test_golden_cv2(ctx, golden, pre, with_matches):
    g = golden("triangulate_cv2.npz")
    mq = g[f"{pre}_mq"] if with_matches else None; mt = g[f"{pre}_mt"] if with_matches else None
    X, keep, nk = ctx.triangulate(g[f"{pre}_K"], g[f"{pre}_Pl"], g[f"{pre}_Pr"], g[f"{pre}_ptsL"], g[f"{pre}_ptsR"], mq, mt)
    _check(X, keep, g[f"{pre}_X"], g[f"{pre}_keep"], g[f"{pre}_err"])
    assert nk == keep.sum()


@pytest.mark.parametrize("m", [1, 31, 129, 100_000])
def test_vs_oracle(ctx, oracle, m):
    p = synth.make_triangulation_problem(m, seed=m)
    X, keep, nk = ctx.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"])
    Xo, ko, eo = oracle.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"])
    _check(X, keep, Xo, ko, eo)
    assert nk == int(keep.sum())


def test_empty_and_invalid(ctx, golden):
    g = golden("triangulate_cv2.npz")
    X, keep, nk = ctx.triangulate(g["fx_K"], g["fx_Pl"], g["fx_Pr"], np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert X.shape == (0, 3) and nk == 0
    with pytest.raises(capi.SfmB200Error):
        ctx.triangulate(g["fx_K"], g["fx_Pl"], g["fx_Pr"], g["fx_ptsL"], g["fx_ptsR"], np.array([0, 99], np.int32), np.array([0, 1], np.int32))


def test_config5_one_million_points(ctx):
    """BASELINE.json config 5 (M = 1e6): properties that need no O(M) CPU SVD: noise-free points are recovered, the
    gross outliers are dropped, every kept point reprojects within 10 px (float64 check), result independent of batch split."""
    p = synth.make_triangulation_problem(1_000_000, seed=0)
    X, keep, nk = ctx.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"])
    assert nk == int(keep.sum()) and 0.95e6 < nk < 1e6
    k = keep.astype(bool)
    assert np.median(np.linalg.norm(X[k] - p["X_true"][k], axis=1)) < 0.5
    for P, pts in ((p["Pl"], p["ptsL"]), (p["Pr"], p["ptsR"])):
        Xc = X[k].astype(np.float64) @ P[:, :3].astype(np.float64).T + P[:, 3]
        uv = Xc[:, :2] / Xc[:, 2:3] * 700.0 + [320.0, 240.0]
        assert np.linalg.norm(uv - pts[k], axis=1).max() < 10.0 + 1e-2
    X2, keep2, _ = ctx.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"][:1000], p["ptsR"][:1000])
    np.testing.assert_array_equal(X2, X[:1000]); np.testing.assert_array_equal(keep2, keep[:1000])
