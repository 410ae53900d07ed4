"""oracle/dense_lm.py -- TEST INFRASTRUCTURE ONLY.

A third, independently written statement of what `ceres::Solve` does for adjustBundle (SfMBundleAdjustmentUtils.cpp:171-179),
used to pin the LM trajectory of oracle/ba_oracle.c (and through it the GPU solver) to something that shares no code and
no algebra with either:

  * the Jacobian is formed EXPLICITLY as a dense (2*nobs) x (6*nc + 3*np + 1) matrix by complex-step differentiation of a
    numpy statement of SimpleReprojectionError (SfMBundleAdjustmentUtils.cpp:62-88) -- no dual numbers (ba_oracle.c) and
    no closed-form derivative (csrc/ba_math.cuh);
  * the LM step solves the FULL damped normal equations (J^T J + D^2) with numpy's Cholesky -- no Schur complement, no
    per-point elimination;
  * the control flow follows Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy defaults:
    Jacobi scaling 1/(1+||col||) fixed at x0, diagonal clamp [1e-6, 1e32] / radius, rho = cost_change / model_cost_change,
    radius /= max(1/3, 1-(2rho-1)^3) on success, radius /= decrease_factor (2, 4, 8, ...) on failure, the
    gradient / parameter / function tolerance tests in Ceres' order.

Only practical for small problems (a dense J): <= ~10 cameras, a few hundred points.  Parameter order: cameras (6 each),
points (3 each), focal.
"""
import numpy as np


def _rotate(w, X):
    """ceres::AngleAxisRotatePoint (rotation.h), complex-safe: theta^2 = w.w without conjugation."""
    th2 = (w * w).sum(1, keepdims=True)
    small = np.abs(th2) <= np.finfo(np.float64).eps
    th2s = np.where(small, 1.0, th2)
    th = np.sqrt(th2s)
    k = w / th
    c, s = np.cos(th), np.sin(th)
    big = X * c + np.cross(k, X) * s + k * (k * X).sum(1, keepdims=True) * (1 - c)
    lin = X + np.cross(w, X)
    return np.where(small, lin, big)


def residuals(cams, pts, focal, obs_xy, obs_cam, obs_pt):
    """SimpleReprojectionError::operator() for every observation: [nobs, 2] (x then y)."""
    P = _rotate(cams[obs_cam, :3], pts[obs_pt]) + cams[obs_cam, 3:]
    return focal * P[:, :2] / P[:, 2:3] - obs_xy


def dense_jacobian(cams, pts, focal, obs_xy, obs_cam, obs_pt, h=1e-40):
    nc, npt, nobs = cams.shape[0], pts.shape[0], obs_cam.shape[0]
    n = 6 * nc + 3 * npt + 1
    J = np.zeros((2 * nobs, n))
    rows = 2 * np.arange(nobs)
    cc = cams.astype(np.complex128); pc = pts.astype(np.complex128)
    for k in range(6):                                                  # d r / d camera parameter k (of the observing camera)
        c2 = cc.copy(); c2[:, k] += 1j * h
        d = residuals(c2, pc, focal, obs_xy, obs_cam, obs_pt).imag / h
        J[rows, 6 * obs_cam + k] = d[:, 0]; J[rows + 1, 6 * obs_cam + k] = d[:, 1]
    for k in range(3):
        p2 = pc.copy(); p2[:, k] += 1j * h
        d = residuals(cc, p2, focal, obs_xy, obs_cam, obs_pt).imag / h
        J[rows, 6 * nc + 3 * obs_pt + k] = d[:, 0]; J[rows + 1, 6 * nc + 3 * obs_pt + k] = d[:, 1]
    d = residuals(cc, pc, focal + 1j * h, obs_xy, obs_cam, obs_pt).imag / h
    J[rows, n - 1] = d[:, 0]; J[rows + 1, n - 1] = d[:, 1]
    return J


def solve(cams, pts, focal, obs_xy, obs_cam, pt_off, max_num_iterations=500, function_tolerance=1e-6, gradient_tolerance=1e-10,
          parameter_tolerance=1e-8, initial_radius=1e4, max_radius=1e16, min_radius=1e-32, min_relative_decrease=1e-3,
          min_lm_diagonal=1e-6, max_lm_diagonal=1e32, max_consecutive_invalid=5):
    """Returns dict(cams, pts, focal, termination ('CONVERGENCE' | 'NO_CONVERGENCE' | 'FAILURE'), iterations,
    successful, unsuccessful, costs [per iteration: cost of x after the iteration], radii, message)."""
    cams = np.array(cams, np.float64).reshape(-1, 6); pts = np.array(pts, np.float64).reshape(-1, 3); focal = float(focal)
    obs_xy = np.asarray(obs_xy, np.float32).astype(np.float64).reshape(-1, 2)
    obs_cam = np.asarray(obs_cam, np.int64); pt_off = np.asarray(pt_off, np.int64)
    obs_pt = np.repeat(np.arange(len(pts)), np.diff(pt_off))
    nc, npt = len(cams), len(pts)

    def unpack(x):
        return x[:6 * nc].reshape(nc, 6), x[6 * nc:6 * nc + 3 * npt].reshape(npt, 3), x[-1]

    def evaluate(x, want_jac):
        c, p, f = unpack(x)
        r = residuals(c, p, f, obs_xy, obs_cam, obs_pt).reshape(-1)
        if not want_jac:
            return 0.5 * r @ r, r, None
        return 0.5 * r @ r, r, dense_jacobian(c, p, f, obs_xy, obs_cam, obs_pt)

    x = np.concatenate([cams.ravel(), pts.ravel(), [focal]])
    cost, r, J = evaluate(x, True)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))                       # jacobi_scaling, once, at x0
    g = J.T @ r                                                         # gradient of the UNSCALED problem
    Js = J * scale
    x_norm = np.linalg.norm(x)
    radius, decrease_factor = initial_radius, 2.0
    reuse_diagonal = False; diag = None
    costs = [cost]; radii = [radius]
    out = dict(initial_cost=cost, iterations=0, successful=0, unsuccessful=0)
    invalid = 0; it = 0
    term, msg = None, ""
    if np.abs(g).max() <= gradient_tolerance:
        term, msg = "CONVERGENCE", "gradient tolerance at x0"
    while term is None:
        if it >= max_num_iterations:
            term, msg = "NO_CONVERGENCE", "max iterations"; break
        if np.abs(g).max() <= gradient_tolerance:
            term, msg = "CONVERGENCE", "gradient tolerance"; break
        if radius <= min_radius:
            term, msg = "CONVERGENCE", "min trust region radius"; break
        it += 1
        if not reuse_diagonal:
            diag = np.clip((Js * Js).sum(0), min_lm_diagonal, max_lm_diagonal)
        D2 = diag / radius
        H = Js.T @ Js + np.diag(D2)
        try:
            L = np.linalg.cholesky(H)
            step = -np.linalg.solve(L.T, np.linalg.solve(L, Js.T @ r))
            ok = np.all(np.isfinite(step))
        except np.linalg.LinAlgError:
            ok = False
        if ok:
            m = Js @ step
            model_cost_change = -m @ (r + 0.5 * m)
            ok = model_cost_change > 0
        if not ok:
            invalid += 1
            if invalid >= max_consecutive_invalid:
                term, msg = "FAILURE", "too many invalid steps"; break
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = True   # StepIsInvalid() == StepRejected(0)
            out["unsuccessful"] += 1
            costs.append(cost); radii.append(radius)
            continue
        invalid = 0
        delta = step * scale
        x_new = x + delta
        cost_new, _, _ = evaluate(x_new, False)
        step_norm = np.linalg.norm(delta)
        if step_norm <= parameter_tolerance * (x_norm + parameter_tolerance):
            term, msg = "CONVERGENCE", "parameter tolerance"; break
        cost_change = cost - cost_new
        if abs(cost_change) <= function_tolerance * cost:
            # Ceres reports convergence BEFORE accepting the step: the iterate stays x (FunctionToleranceReached)
            term, msg = "CONVERGENCE", "function tolerance"; break
        rho = cost_change / model_cost_change
        if rho > min_relative_decrease:
            x = x_new; x_norm = np.linalg.norm(x)
            cost, r, J = evaluate(x, True)
            g = J.T @ r; Js = J * scale
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease_factor = 2.0; reuse_diagonal = False
            out["successful"] += 1
        else:
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = True
            out["unsuccessful"] += 1
        costs.append(cost); radii.append(radius)
    c, p, f = unpack(x)
    out.update(cams=c.copy(), pts=p.copy(), focal=float(f), termination=term, message=msg, iterations=it, final_cost=cost,
               costs=np.array(costs), radii=np.array(radii))
    return out
