"""oracle/ba_oracle.c: the projection model is pinned by the reference's ceres_reprojection_test fixture
(SfMUnitTests.cpp:153-189, golden made with cv2.projectPoints); the solver (un-vendored Ceres, PARITY UNPINNED by the
reference's own tests) is cross-checked against finite differences and scipy's independent optimum."""
import numpy as np
import pytest

from sfm_toy_library_b200 import synth


def test_projection_model_matches_cv_projectpoints(oracle, golden):
    g = golden("reproj_fixture.npz")
    aa = oracle.rotmat_to_angle_axis_f32(g["R"])            # RotationMatrixToAngleAxis<float>(R.t().val), :126
    np.testing.assert_allclose(aa, g["rvec"], atol=2e-6)    # == cv::Rodrigues(R)
    np.testing.assert_allclose(aa, [0.08334807, 0.09095846, 0.08334807], atol=2e-6)
    cam = np.concatenate([aa.astype(np.float64), g["t"].astype(np.float64)])
    f, cx, cy = float(g["K"][0, 0]), float(g["K"][0, 2]), float(g["K"][1, 2])
    for X, uv in zip(g["points3d"], g["points2d"]):
        r, _, _, _ = oracle.ba_residual_jacobian(cam, X.astype(np.float64), f, 0.0, 0.0)
        assert abs(r[0] + cx - uv[0]) < 1e-3 and abs(r[1] + cy - uv[1]) < 1e-3     # reference tolerance is 0.1 px


def test_rotation_round_trip(oracle):
    rs = np.random.RandomState(0)
    for _ in range(50):
        w = rs.normal(0, 1.2, 3)
        R = oracle.angle_axis_to_rotmat(w)
        np.testing.assert_allclose(R, synth.angle_axis_to_rotmat(w), atol=1e-14)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-14)
        w2 = oracle.rotmat_to_angle_axis_f32(R.astype(np.float32))
        np.testing.assert_allclose(synth.angle_axis_to_rotmat(w2), R, atol=5e-6)
    # trace < 0 branch of the quaternion conversion and the zero rotation
    R = synth.angle_axis_to_rotmat([0.0, 3.0, 0.3]).astype(np.float32)
    np.testing.assert_allclose(synth.angle_axis_to_rotmat(oracle.rotmat_to_angle_axis_f32(R)), R, atol=5e-6)
    assert np.all(oracle.rotmat_to_angle_axis_f32(np.eye(3, dtype=np.float32)) == 0)


def test_jets_equal_closed_form_and_finite_differences(oracle):
    rs = np.random.RandomState(1)
    for trial in range(40):
        cam = np.concatenate([rs.normal(0, 0.7, 3), rs.normal(0, 1, 3) + [0, 0, 8]])
        if trial % 8 == 0:
            cam[:3] = rs.normal(0, 1e-9, 3)                 # Taylor branch of AngleAxisRotatePoint
        pt = rs.uniform(-2, 2, 3); f = 2500.0 * rs.uniform(0.9, 1.1); ox, oy = rs.normal(0, 100, 2)
        r0, Jc0, Jp0, Jf0 = oracle.ba_residual_jacobian(cam, pt, f, ox, oy, mode=0)
        r1, Jc1, Jp1, Jf1 = oracle.ba_residual_jacobian(cam, pt, f, ox, oy, mode=1)
        np.testing.assert_allclose(r0, r1, rtol=0, atol=1e-9)
        for a, b in ((Jc0, Jc1), (Jp0, Jp1), (Jf0, Jf1)):
            np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-8)
        if trial % 8 == 0:
            continue                                        # finite differences straddle the branch
        x = np.concatenate([cam, pt, [f]]); J = np.concatenate([Jc0, Jp0, Jf0[:, None]], 1)
        for k in range(10):
            h = 1e-6 * max(1.0, abs(x[k])); xp = x.copy(); xm = x.copy(); xp[k] += h; xm[k] -= h
            rp = oracle.ba_residual_jacobian(xp[:6], xp[6:9], xp[9], ox, oy)[0]
            rm = oracle.ba_residual_jacobian(xm[:6], xm[6:9], xm[9], ox, oy)[0]
            np.testing.assert_allclose((rp - rm) / (2 * h), J[:, k], rtol=2e-5, atol=1e-4)


def test_cost_matches_numpy_model(oracle, golden):
    g = golden("ba_scipy.npz")
    c = oracle.ba_cost(g["cams"], g["pts"], float(g["focal"]), g["obs_xy"], g["obs_cam"], g["pt_off"])
    assert abs(c - float(g["cost0"])) < 1e-8 * float(g["cost0"])


@pytest.mark.parametrize("mode", [0, 1])
def test_solve_reaches_scipy_optimum(oracle, golden, mode):
    g = golden("ba_scipy.npz")
    o = oracle.ba_default_options(jacobian_mode=mode, max_solver_time_in_seconds=0.0, function_tolerance=1e-12,
                                  parameter_tolerance=1e-12, max_num_iterations=200)
    cams, pts, f, s = oracle.ba_solve(g["cams"], g["pts"], float(g["focal"]), g["obs_xy"], g["obs_cam"], g["pt_off"], o)
    assert s["termination_type"] == 0, s
    assert abs(s["initial_cost"] - float(g["cost0"])) < 1e-8 * float(g["cost0"])
    assert s["final_cost"] <= float(g["cost_opt"]) * (1 + 1e-6), (s["final_cost"], float(g["cost_opt"]))
    assert abs(s["final_cost"] - float(g["cost_opt"])) < 2e-3 * float(g["cost_opt"])


def test_default_options_converge_on_cfg2_like(oracle):
    p = synth.make_ba_problem(n_cams=20, n_pts=2000, obs_per_pt=8, seed=0)
    o = oracle.ba_default_options(jacobian_mode=1)
    cams, pts, f, s, trace = oracle.ba_solve(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], o, want_trace=True)
    assert s["termination_type"] == 0 and s["num_iterations"] < 60, s
    rms0 = np.sqrt(2 * s["initial_cost"] / p["nobs"]); rms1 = np.sqrt(2 * s["final_cost"] / p["nobs"])
    assert rms0 > 10 and rms1 < 0.75, (rms0, rms1)       # noise is 0.5 px/axis -> ~0.7 px RMS at the optimum
    assert np.all(np.diff(trace[:, 0]) <= 1e-12)          # monotone (use_nonmonotonic_steps = false)
    assert abs(f - p["focal_true"]) < 5.0


def test_reduced_system_is_the_schur_complement(oracle):
    """S, rhs from the oracle equal the dense Schur complement of (J^T J + D^2) computed with numpy."""
    p = synth.make_ba_problem(n_cams=4, n_pts=30, obs_per_pt=3, seed=5)
    out = oracle.ba_reduced_system(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], radius=1e4)
    nc, npt, nobs = p["nc"], p["np"], p["nobs"]
    n = 6 * nc + 3 * npt + 1
    J = np.zeros((2 * nobs, n)); r = np.zeros(2 * nobs)
    for o in range(nobs):
        c, q = p["obs_cam"][o], p["obs_pt"][o]
        ro, Jc, Jp, Jf = oracle.ba_residual_jacobian(p["cams"][c], p["pts"][q], p["focal"], *p["obs_xy"][o].astype(float), mode=1)
        J[2 * o:2 * o + 2, 6 * c:6 * c + 6] = Jc; J[2 * o:2 * o + 2, 6 * nc + 3 * q:6 * nc + 3 * q + 3] = Jp
        J[2 * o:2 * o + 2, -1] = Jf; r[2 * o:2 * o + 2] = ro
    np.testing.assert_allclose(out["grad"], J.T @ r, rtol=1e-10)
    scale = 1.0 / (1.0 + np.linalg.norm(J, axis=0)); np.testing.assert_allclose(out["scale"], scale, rtol=1e-12)
    Js = J * scale
    H = Js.T @ Js; d = np.clip(np.diag(H), 1e-6, 1e32) / 1e4; H = H + np.diag(d); b = Js.T @ r
    ic = np.r_[0:6 * nc, n - 1]; ip = np.r_[6 * nc:6 * nc + 3 * npt]
    Hpp_inv = np.linalg.inv(H[np.ix_(ip, ip)])
    S = H[np.ix_(ic, ic)] - H[np.ix_(ic, ip)] @ Hpp_inv @ H[np.ix_(ip, ic)]
    rhs = b[ic] - H[np.ix_(ic, ip)] @ Hpp_inv @ b[ip]
    np.testing.assert_allclose(out["S"], S, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["rhs"], rhs, rtol=1e-9, atol=1e-12)


def test_cfg2_reaches_the_independent_scipy_optimum(oracle, golden):
    """BASELINE configs[1] at full size: the oracle's LM + DENSE_SCHUR lands on the optimum that scipy's TRF (sparse analytic
    Jacobian by complex-step differentiation of an independent numpy model, tests/golden/make_golden.py:golden_ba_cfg2)
    finds -- cost to 1e-9 relative, well inside the 1e-6 the round-1 review asked for."""
    g = golden("ba_scipy_cfg2.npz")
    p = synth.make_ba_problem(seed=0, **synth.BA_CONFIGS["cfg2"])
    o = oracle.ba_default_options(jacobian_mode=1, num_threads=4, max_solver_time_in_seconds=0.0, function_tolerance=1e-14,
                                  parameter_tolerance=1e-14, max_num_iterations=60)
    cams, pts, f, s = oracle.ba_solve(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], o)
    assert abs(s["initial_cost"] - float(g["cost0"])) < 1e-10 * float(g["cost0"])
    assert abs(s["final_cost"] - float(g["cost_opt"])) < 1e-9 * float(g["cost_opt"]), (s["final_cost"], float(g["cost_opt"]))
    assert abs(f - float(g["focal"])) < 1e-5 * f
    # (cameras / points are only defined up to the 7-dof similarity gauge: the cost and the focal length are the invariants)
    # Ceres' default tolerances stop a little earlier: still within 1e-6 of the optimum
    o = oracle.ba_default_options(jacobian_mode=1, num_threads=4, max_solver_time_in_seconds=0.0)
    s = oracle.ba_solve(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], o)[3]
    assert s["termination_type"] == 0 and 0 <= s["final_cost"] - float(g["cost_opt"]) < 1e-6 * float(g["cost_opt"]), s
