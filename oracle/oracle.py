"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings to the CPU oracle (oracle/*.c), the plain-C restatement of the reference hot path
(SfM2DFeatureUtilities.cpp:53-71, SfMStereoUtilities.cpp:120-206, SfMBundleAdjustmentUtils.cpp:58-222).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsfm_oracle.so")
_lib = None

RATIO_REFERENCE = float(np.float64(np.float32(0.8)))  # (double)0.8f, SfM2DFeatureUtilities.cpp:35


def build(force=False):
    """Compile oracle/*.c -> oracle/_build/libsfm_oracle.so (gcc, seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("match_oracle.c", "triangulate_oracle.c", "ba_oracle.c", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


class BAOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("max_solver_time_in_seconds", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("jacobi_scaling", C.c_int),
                ("max_num_consecutive_invalid_steps", C.c_int), ("jacobian_mode", C.c_int),
                ("num_threads", C.c_int), ("verbose", C.c_int)]


class BASummary(C.Structure):
    _fields_ = [("termination_type", C.c_int), ("num_iterations", C.c_int),
                ("num_successful_steps", C.c_int), ("num_unsuccessful_steps", C.c_int),
                ("num_jacobian_evals", C.c_int), ("num_residual_evals", C.c_int), ("num_linear_solves", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("total_time_s", C.c_double), ("jacobian_time_s", C.c_double), ("linear_solve_time_s", C.c_double),
                ("message", C.c_char * 160)]


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        _lib.sfm_oracle_ba_cost.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


# ---------------------------------------------------------------- matching (a-1)
def knn2_hamming(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nb = q.shape; nt = t.shape[0]
    idx = np.empty((nq, 2), np.int32); dist = np.empty((nq, 2), np.int32)
    lib().sfm_oracle_knn2_hamming(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb, _p(idx, C.c_int32), _p(dist, C.c_int32))
    return idx, dist


def match_hamming(q, t, ratio=RATIO_REFERENCE):
    """matchFeatures restated: returns (queryIdx, trainIdx, distance) of the ratio-test survivors."""
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq = q.shape[0]; nb = q.shape[1] if q.ndim == 2 else 0; nt = t.shape[0]
    oq = np.empty(max(nq, 1), np.int32); ot = np.empty(max(nq, 1), np.int32); od = np.empty(max(nq, 1), np.float32)
    si = np.empty((max(nq, 1), 2), np.int32); sd = np.empty((max(nq, 1), 2), np.int32)
    n = lib().sfm_oracle_match_hamming(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb, C.c_double(ratio),
                                       _p(oq, C.c_int32), _p(ot, C.c_int32), _p(od, C.c_float),
                                       _p(si, C.c_int32), _p(sd, C.c_int32))
    return oq[:n].copy(), ot[:n].copy(), od[:n].copy()


def match_l2(q, t, ratio=RATIO_REFERENCE):
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    nq, dim = q.shape; nt = t.shape[0]
    oq = np.empty(max(nq, 1), np.int32); ot = np.empty(max(nq, 1), np.int32); od = np.empty(max(nq, 1), np.float32)
    si = np.empty((max(nq, 1), 2), np.int32); sd = np.empty((max(nq, 1), 2), np.float32)
    n = lib().sfm_oracle_match_l2(_p(q, C.c_float), nq, _p(t, C.c_float), nt, dim, C.c_double(ratio),
                                  _p(oq, C.c_int32), _p(ot, C.c_int32), _p(od, C.c_float),
                                  _p(si, C.c_int32), _p(sd, C.c_float))
    return oq[:n].copy(), ot[:n].copy(), od[:n].copy()


# ---------------------------------------------------------------- triangulation (a-2, a-6)
def triangulate(K, Pl, Pr, ptsL, ptsR, mq=None, mt=None, max_reproj=10.0):
    """triangulateViews restated.  Returns X [m,3] float32 (all), keep [m] uint8, err [m,2] float64."""
    K = np.ascontiguousarray(K, np.float32).reshape(9); Pl = np.ascontiguousarray(Pl, np.float32).reshape(12)
    Pr = np.ascontiguousarray(Pr, np.float32).reshape(12)
    ptsL = np.ascontiguousarray(ptsL, np.float32).reshape(-1, 2); ptsR = np.ascontiguousarray(ptsR, np.float32).reshape(-1, 2)
    if mq is not None:
        mq = np.ascontiguousarray(mq, np.int32); mt = np.ascontiguousarray(mt, np.int32); m = mq.shape[0]
    else:
        m = ptsL.shape[0]
    X = np.empty((max(m, 1), 3), np.float32); keep = np.empty(max(m, 1), np.uint8); err = np.empty((max(m, 1), 2), np.float64)
    lib().sfm_oracle_triangulate(_p(K, C.c_float), _p(Pl, C.c_float), _p(Pr, C.c_float), _p(ptsL, C.c_float),
                                 _p(ptsR, C.c_float), _p(mq, C.c_int32), _p(mt, C.c_int32), m, C.c_float(max_reproj),
                                 _p(X, C.c_float), _p(keep, C.c_uint8), _p(err, C.c_double))
    return X[:m], keep[:m], err[:m]


# ---------------------------------------------------------------- bundle adjustment (a-3, a-4)
def ba_default_options(**kw):
    o = BAOptions()
    lib().sfm_oracle_ba_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def rotmat_to_angle_axis_f32(R_rowmajor):
    """ceres::RotationMatrixToAngleAxis<float>(R.t().val) as at SfMBundleAdjustmentUtils.cpp:126:
    R.t() stored row-major == R stored column-major."""
    R = np.ascontiguousarray(R_rowmajor, np.float32).reshape(3, 3)
    # column-major storage of R is the row-major storage of R^T
    Rc = np.ascontiguousarray(R.T.reshape(-1), np.float32)
    aa = np.empty(3, np.float32)
    lib().sfm_oracle_rotmat_colmajor_to_angle_axis_f32(_p(Rc, C.c_float), _p(aa, C.c_float))
    return aa


def angle_axis_to_rotmat(aa):
    """ceres::AngleAxisToRotationMatrix + the reference's transpose-on-write-back (:203-209): row-major R."""
    aa = np.ascontiguousarray(aa, np.float64); Rc = np.empty(9, np.float64)
    lib().sfm_oracle_angle_axis_to_rotmat_colmajor(_p(aa, C.c_double), _p(Rc, C.c_double))
    return Rc.reshape(3, 3).T.copy()  # column-major buffer -> row-major matrix


def ba_residual_jacobian(cam, pt, focal, ox, oy, mode=0):
    cam = np.ascontiguousarray(cam, np.float64); pt = np.ascontiguousarray(pt, np.float64)
    r = np.empty(2); Jc = np.empty((2, 6)); Jp = np.empty((2, 3)); Jf = np.empty(2)
    lib().sfm_oracle_ba_residual_jacobian(_p(cam, C.c_double), _p(pt, C.c_double), C.c_double(focal), C.c_double(ox),
                                          C.c_double(oy), mode, _p(r, C.c_double), _p(Jc, C.c_double), _p(Jp, C.c_double),
                                          _p(Jf, C.c_double))
    return r, Jc, Jp, Jf


def _flat(cams, pts, obs_xy, obs_cam, pt_off):
    cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 6); pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    obs_xy = np.ascontiguousarray(obs_xy, np.float32).reshape(-1, 2); obs_cam = np.ascontiguousarray(obs_cam, np.int32)
    pt_off = np.ascontiguousarray(pt_off, np.int32)
    return cams, pts, obs_xy, obs_cam, pt_off


def ba_cost(cams, pts, focal, obs_xy, obs_cam, pt_off, nthreads=1):
    cams, pts, obs_xy, obs_cam, pt_off = _flat(cams, pts, obs_xy, obs_cam, pt_off)
    return lib().sfm_oracle_ba_cost(cams.shape[0], pts.shape[0], obs_cam.shape[0], _p(cams, C.c_double), _p(pts, C.c_double),
                                    C.c_double(focal), _p(obs_xy, C.c_float), _p(obs_cam, C.c_int32), _p(pt_off, C.c_int32),
                                    nthreads)


def ba_reduced_system(cams, pts, focal, obs_xy, obs_cam, pt_off, radius=1e4, jacobi_scaling=1, scale=None,
                      min_diag=1e-6, max_diag=1e32, mode=1):
    """Reduced camera(+focal) system S, rhs of DENSE_SCHUR at x; also Jacobi scale, gradient (unscaled), cost."""
    cams, pts, obs_xy, obs_cam, pt_off = _flat(cams, pts, obs_xy, obs_cam, pt_off)
    nc, np_, nobs = cams.shape[0], pts.shape[0], obs_cam.shape[0]
    nr = 6 * nc + 1; n = 6 * nc + 3 * np_ + 1
    S = np.empty((nr, nr)); rhs = np.empty(nr); sc = np.empty(n); g = np.empty(n); cost = C.c_double()
    if scale is not None:
        scale = np.ascontiguousarray(scale, np.float64)
    ok = lib().sfm_oracle_ba_reduced_system(nc, np_, nobs, _p(cams, C.c_double), _p(pts, C.c_double), C.c_double(focal),
                                            _p(obs_xy, C.c_float), _p(obs_cam, C.c_int32), _p(pt_off, C.c_int32),
                                            jacobi_scaling, _p(scale, C.c_double), C.c_double(radius), C.c_double(min_diag),
                                            C.c_double(max_diag), mode, _p(S, C.c_double), _p(rhs, C.c_double),
                                            _p(sc, C.c_double), _p(g, C.c_double), C.byref(cost))
    return dict(ok=bool(ok), S=S, rhs=rhs, scale=sc, grad=g, cost=cost.value)


def ba_solve(cams, pts, focal, obs_xy, obs_cam, pt_off, options=None, want_trace=False):
    """Ceres-equivalent LM/DENSE_SCHUR.  Returns (cams, pts, focal, summary dict[, trace])."""
    cams, pts, obs_xy, obs_cam, pt_off = _flat(cams, pts, obs_xy, obs_cam, pt_off)
    cams = cams.copy(); pts = pts.copy(); f = C.c_double(float(focal))
    o = options or ba_default_options()
    s = BASummary()
    trace = np.zeros((o.max_num_iterations + 1, 4)) if want_trace else None
    lib().sfm_oracle_ba_solve(C.byref(o), cams.shape[0], pts.shape[0], obs_cam.shape[0], _p(cams, C.c_double),
                              _p(pts, C.c_double), C.byref(f), _p(obs_xy, C.c_float), _p(obs_cam, C.c_int32),
                              _p(pt_off, C.c_int32), C.byref(s), _p(trace, C.c_double))
    summ = {k: getattr(s, k) for k, _ in BASummary._fields_}
    summ["message"] = s.message.decode()
    if want_trace:
        return cams, pts, f.value, summ, trace[: s.num_iterations + 1]
    return cams, pts, f.value, summ
