"""Pins the LM trajectory of oracle/ba_oracle.c (Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_SCHUR as
used by adjustBundle, SfMBundleAdjustmentUtils.cpp:171-179) against oracle/dense_lm.py: an independently written solver
with an EXPLICIT dense Jacobian (complex-step differentiation of a numpy model) and the FULL damped normal equations --
no dual numbers, no closed-form derivative, no Schur complement.  Same accept/reject sequence, same per-iteration costs
and radii, same termination, same final state."""
import numpy as np
import pytest

from oracle import dense_lm
from sfm_toy_library_b200 import synth

CASES = [dict(n_cams=4, n_pts=60, obs_per_pt=3, seed=5), dict(n_cams=6, n_pts=150, obs_per_pt=4, seed=11),
         dict(n_cams=10, n_pts=300, obs_per_pt=5, seed=2), dict(n_cams=3, n_pts=80, obs_per_pt=2, seed=9)]


def _compare(o, trace, s, d):
    assert ["CONVERGENCE", "NO_CONVERGENCE", "FAILURE"][s["termination_type"]] == d["termination"], (s, d["message"])
    assert s["num_iterations"] == d["iterations"], (s["num_iterations"], d["iterations"])
    assert (s["num_successful_steps"], s["num_unsuccessful_steps"]) == (d["successful"], d["unsuccessful"])
    n = min(len(trace), len(d["costs"]))
    np.testing.assert_allclose(trace[:n, 0], d["costs"][:n], rtol=1e-9)
    np.testing.assert_allclose(trace[:n, 1], d["radii"][:n], rtol=1e-6)


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("mode", [0, 1])
def test_trajectory_equals_dense_lm(oracle, case, mode):
    p = synth.make_ba_problem(**CASES[case])
    o = oracle.ba_default_options(jacobian_mode=mode, max_solver_time_in_seconds=0.0)
    cams, pts, f, s, trace = oracle.ba_solve(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"], o, want_trace=True)
    d = dense_lm.solve(p["cams"], p["pts"], p["focal"], p["obs_xy"], p["obs_cam"], p["pt_off"])
    _compare(o, trace, s, d)
    np.testing.assert_allclose(cams, d["cams"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(pts, d["pts"], rtol=1e-7, atol=1e-9)
    assert abs(f - d["focal"]) < 1e-7 * f


@pytest.mark.parametrize("k", [0, 1, 2])
def test_crazyhorse_calls(oracle, k):
    """The first three adjustBundle calls of the crazyhorse replay (2, 3, 4 cameras of real data): call 0 runs 108 LM
    iterations along the gauge valley of a two-view problem, call 1 contains REJECTED steps (22 accepted, 2 rejected)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_crazyhorse.npz"))
    pre = f"ba_{k}_"
    args = (g[pre + "cams"], g[pre + "pts"], float(g[pre + "focal"]), g[pre + "obs_xy"], g[pre + "obs_cam"], g[pre + "pt_off"])
    o = oracle.ba_default_options(jacobian_mode=0, max_solver_time_in_seconds=0.0)
    c, q, f, s, trace = oracle.ba_solve(*args, o, want_trace=True)
    d = dense_lm.solve(*args)
    _compare(o, trace, s, d)
    if k == 1:
        assert d["unsuccessful"] == 2
    np.testing.assert_allclose(q, d["pts"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose([s["initial_cost"], s["final_cost"]], g[pre + "cost"], rtol=1e-12)
