"""Pins oracle/triangulate_oracle.c against cv2 golden vectors (tests/golden/triangulate_cv2.npz) incl. the reference's
own triangulate_from_2_views fixture (SfMUnitTests.cpp:221-251, tolerance 0.01)."""
import numpy as np
import pytest

REL_TOL = 1e-5        # 3D points, relative (float32 outputs)
ERR_TOL = 1e-3        # reprojection error, px; keep mask may differ only this close to the 10 px threshold


def _check(oracle, g, pre, with_matches=False):
    mq = g[f"{pre}_mq"] if with_matches else None
    mt = g[f"{pre}_mt"] if with_matches else None
    X, keep, err = oracle.triangulate(g[f"{pre}_K"], g[f"{pre}_Pl"], g[f"{pre}_Pr"], g[f"{pre}_ptsL"], g[f"{pre}_ptsR"], mq, mt)
    Xr, kr, er = g[f"{pre}_X"], g[f"{pre}_keep"], g[f"{pre}_err"]
    good = kr.astype(bool)
    rel = np.linalg.norm(X[good] - Xr[good], axis=1) / np.linalg.norm(Xr[good], axis=1)
    assert rel.max() < REL_TOL, rel.max()
    assert np.abs(err[good] - er[good]).max() < ERR_TOL
    diff = keep != kr
    if diff.any():
        assert (np.abs(np.max(er[diff], axis=1) - 10.0) < ERR_TOL).all()
    return X, keep


def test_reference_fixture(oracle, golden):
    g = golden("triangulate_cv2.npz")
    X, keep = _check(oracle, g, "fx")
    assert keep.all()
    assert np.linalg.norm(X - g["fx_points3d"], axis=1).max() < 0.01      # the reference's own tolerance


def test_synthetic_with_matches_and_outliers(oracle, golden):
    g = golden("triangulate_cv2.npz")
    X, keep = _check(oracle, g, "sy", with_matches=True)
    assert 0 < keep.sum() < len(keep)


def test_crazyhorse_like_geometry(oracle, golden):
    _check(oracle, golden("triangulate_cv2.npz"), "ch")


def test_empty(oracle, golden):
    g = golden("triangulate_cv2.npz")
    X, keep, err = oracle.triangulate(g["fx_K"], g["fx_Pl"], g["fx_Pr"], np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert X.shape == (0, 3) and keep.shape == (0,)


def test_vs_cv2_live(oracle):
    pytest.importorskip("cv2")
    from oracle import cv2_reference as ref
    from sfm_toy_library_b200 import synth
    p = synth.make_triangulation_problem(20000, seed=9)
    X, keep, err = oracle.triangulate(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"])
    Xr, kr, er = ref.triangulate_views(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"])
    good = kr.astype(bool)
    rel = np.linalg.norm(X[good] - Xr[good], axis=1) / np.linalg.norm(Xr[good], axis=1)
    assert rel.max() < REL_TOL
    diff = keep != kr
    assert diff.sum() == 0 or (np.abs(np.max(er[diff], axis=1) - 10.0) < ERR_TOL).all()
