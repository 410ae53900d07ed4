"""Generates tests/golden/*.npz with the reference's third-party math run HERE:
  * cv2 4.13 (the OpenCV the reference calls, Python binding) for matching / triangulation / projection;
  * scipy.optimize.least_squares (independent solver) for the BA optimum (Ceres is not installable here).
Run from the repo root:  python tests/golden/make_golden.py        (needs cv2 + scipy; the build container has both)
The .npz files are committed; tests never need /root/reference or this script at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from scipy.optimize import least_squares  # noqa: E402
from scipy.sparse import lil_matrix  # noqa: E402

from oracle import cv2_reference as ref  # noqa: E402
from sfm_toy_library_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def golden_match():
    cases = {}
    # (a) ORB-like sets with copies (ratio-test survivors) and duplicate rows (tie-break), ragged sizes
    a = synth.make_descriptors(0, 300); b = synth.make_descriptors(1, 257, prev=a)
    cases["a"] = (b, a)
    # (b) identical rows / d0 = d1 = 0 (strict '<' rejects) and the 4-vs-5 tie the (double)0.8f constant flips
    q = np.zeros((4, 32), np.uint8); t = np.zeros((6, 32), np.uint8)
    t[1] = 0; t[2, 0] = 0x0F; t[3, 0] = 0x1F          # train rows at distance 0,0,4,5 from q[0]
    q[1, 0] = 0x0F; q[1, 1] = 0xFF                    # some other geometry
    q[2] = 0xFF; t[4] = 0xFF; t[4, 0] = 0xF0          # d0 = 4
    t[5] = 0xFF; t[5, 0] = 0xE0; t[5, 1] = 0xFE       # d = 6
    q[3, 5] = 0x3C
    cases["b"] = (q, t)
    # (c) d0=4, d1=5 exactly: 4 < 0.8f*5 = 4.00000006 keeps; exact 0.8 would drop
    q = np.zeros((1, 32), np.uint8); t = np.full((3, 32), 0xFF, np.uint8)
    t[0] = 0; t[0, 0] = 0x0F; t[1] = 0; t[1, 0] = 0x1F
    cases["c"] = (q, t)
    # (d) nt == 2 (minimum defined size), nq == 1
    cases["d"] = (synth.make_descriptors(7, 1), synth.make_descriptors(8, 2))
    # (e) 64-byte descriptors (desc_bytes is a parameter of the ABI)
    cases["e"] = (synth.make_descriptors(9, 120, nbytes=64), synth.make_descriptors(10, 90, nbytes=64))
    out = {}
    for k, (q, t) in cases.items():
        mq, mt, md = ref.match_features(q, t)
        knn = cv2.DescriptorMatcher_create("BruteForce-Hamming").knnMatch(q, t, 2)
        out[f"{k}_q"] = q; out[f"{k}_t"] = t; out[f"{k}_mq"] = mq; out[f"{k}_mt"] = mt; out[f"{k}_md"] = md
        out[f"{k}_knn_idx"] = np.array([[m.trainIdx for m in row] for row in knn], np.int32)
        out[f"{k}_knn_dist"] = np.array([[m.distance for m in row] for row in knn], np.float32)
    assert len(out["c_mq"]) == 1, "the (double)0.8f constant must keep d0=4,d1=5"
    # L2 / SIFT-like (BASELINE.json config 4 wording)
    a = synth.make_sift_like(0, 200); b = synth.make_sift_like(1, 180, prev=a)
    mq, mt, md = ref.match_features(b, a, norm="l2")
    out.update(l2_q=b, l2_t=a, l2_mq=mq, l2_mt=mt, l2_md=md)
    np.savez_compressed(os.path.join(OUT, "match_cv2.npz"), **out)
    print("match:", {k: len(out[f"{k}_mq"]) for k in cases}, "l2:", len(mq))


def golden_triangulate():
    out = {}
    Pl, Pr = synth.fixture_poses()
    K = synth.TEST_K
    # reference unit-test fixture: generateStereoViews (SfMUnitTests.cpp:105-146)
    imgs = []
    for P in (Pl, Pr):
        rvec, _ = cv2.Rodrigues(P[:, :3].copy())
        proj, _ = cv2.projectPoints(synth.CANNED_POINTS, rvec, P[:, 3].copy(), K, None)
        imgs.append(proj.reshape(-1, 2).astype(np.float32))
    X, keep, err = ref.triangulate_views(K, Pl, Pr, imgs[0], imgs[1])
    out.update(fx_K=K, fx_Pl=Pl, fx_Pr=Pr, fx_ptsL=imgs[0], fx_ptsR=imgs[1], fx_X=X, fx_keep=keep, fx_err=err,
               fx_points3d=synth.CANNED_POINTS)
    assert np.abs(X - synth.CANNED_POINTS).max() < 0.01 and keep.all()      # triangulate_from_2_views tolerance
    # synthetic with outliers + a match list (gather path, SfMCommon.cpp:63-87)
    p = synth.make_triangulation_problem(3000, seed=3)
    rs = np.random.RandomState(5)
    mq = rs.permutation(3000)[:2500].astype(np.int32); mt = mq.copy()
    X, keep, err = ref.triangulate_views(p["K"], p["Pl"], p["Pr"], p["ptsL"], p["ptsR"], mq, mt)
    out.update(sy_K=p["K"], sy_Pl=p["Pl"], sy_Pr=p["Pr"], sy_ptsL=p["ptsL"], sy_ptsR=p["ptsR"], sy_mq=mq, sy_mt=mt,
               sy_X=X, sy_keep=keep, sy_err=err)
    print("triangulate: fixture max|dX|", np.abs(out["fx_X"] - synth.CANNED_POINTS).max(), "synthetic kept", keep.sum(), "/", len(keep))
    # the crazyhorse-style setup: K=2500, Pleft = [I|0], Pright = small motion
    K2 = np.array([[2500, 0, 512], [0, 2500, 384], [0, 0, 1]], np.float32)
    Pl2 = np.eye(3, 4, dtype=np.float32)
    Pr2 = np.zeros((3, 4), np.float32); Pr2[:, :3] = synth.euler_deg_to_rotmat(1.5, -4, 0.7).astype(np.float32); Pr2[:, 3] = (-1, 0.05, 0.1)
    Xw = np.stack([rs.uniform(-1, 1, 800), rs.uniform(-0.8, 0.8, 800), rs.uniform(4, 9, 800)], 1)
    pp = []
    for P in (Pl2, Pr2):
        Xc = Xw @ P[:, :3].astype(np.float64).T + P[:, 3]
        pp.append((Xc[:, :2] / Xc[:, 2:3] * 2500 + [512, 384] + rs.normal(0, 0.7, (800, 2))).astype(np.float32))
    X, keep, err = ref.triangulate_views(K2, Pl2, Pr2, pp[0], pp[1])
    out.update(ch_K=K2, ch_Pl=Pl2, ch_Pr=Pr2, ch_ptsL=pp[0], ch_ptsR=pp[1], ch_X=X, ch_keep=keep, ch_err=err)
    np.savez_compressed(os.path.join(OUT, "triangulate_cv2.npz"), **out)


def golden_reprojection():
    """ceres_reprojection_test (SfMUnitTests.cpp:153-189): cv::projectPoints of the 12 canned points."""
    R = synth.euler_deg_to_rotmat(5, 5, 5).astype(np.float32)
    t = np.array([-10, 0, 30], np.float32)
    rvec, _ = cv2.Rodrigues(R)
    proj, _ = cv2.projectPoints(synth.CANNED_POINTS, rvec, t, synth.TEST_K, None)
    np.savez_compressed(os.path.join(OUT, "reproj_fixture.npz"), R=R, t=t, rvec=rvec.reshape(3).astype(np.float32),
                        K=synth.TEST_K, points3d=synth.CANNED_POINTS, points2d=proj.reshape(-1, 2).astype(np.float32))
    print("reproj: rvec", rvec.ravel())


def golden_ba():
    """Independent optimum for a small synthetic BA problem: scipy TRF with the model written in numpy."""
    p = synth.make_ba_problem(n_cams=6, n_pts=150, obs_per_pt=4, seed=11)
    nc, npt = p["nc"], p["np"]
    oc, op, oxy = p["obs_cam"], p["obs_pt"], p["obs_xy"].astype(np.float64)

    def rotate(w, X):
        th = np.linalg.norm(w, axis=1, keepdims=True)
        k = w / th
        return X * np.cos(th) + np.cross(k, X) * np.sin(th) + k * (k * X).sum(1, keepdims=True) * (1 - np.cos(th))

    def fun(x):
        cams = x[:6 * nc].reshape(nc, 6); pts = x[6 * nc:6 * nc + 3 * npt].reshape(npt, 3); f = x[-1]
        P = rotate(cams[oc, :3], pts[op]) + cams[oc, 3:]
        return (f * P[:, :2] / P[:, 2:3] - oxy).ravel()

    x0 = np.concatenate([p["cams"].ravel(), p["pts"].ravel(), [p["focal"]]])
    A = lil_matrix((2 * p["nobs"], x0.size), dtype=int)
    for o in range(p["nobs"]):
        for r in (0, 1):
            A[2 * o + r, 6 * oc[o]:6 * oc[o] + 6] = 1
            A[2 * o + r, 6 * nc + 3 * op[o]:6 * nc + 3 * op[o] + 3] = 1
            A[2 * o + r, -1] = 1
    sol = least_squares(fun, x0, jac_sparsity=A, x_scale="jac", method="trf", ftol=1e-14, xtol=1e-14, gtol=1e-14, max_nfev=300)
    cost0 = 0.5 * np.sum(fun(x0) ** 2)
    print("ba: cost0", cost0, "-> scipy optimum cost", sol.cost, "nfev", sol.nfev)
    np.savez_compressed(os.path.join(OUT, "ba_scipy.npz"), cams=p["cams"], pts=p["pts"], focal=p["focal"], obs_xy=p["obs_xy"],
                        obs_cam=p["obs_cam"], pt_off=p["pt_off"], cost0=cost0, cost_opt=sol.cost,
                        res0=fun(x0))


if __name__ == "__main__":
    golden_match(); golden_triangulate(); golden_reprojection(); golden_ba()


def golden_ba_cfg2():
    """BASELINE configs[1] (cfg 2: 20 cameras / 10 k points / 80 k observations, synth.make_ba_problem(seed=0)) solved to its
    optimum by scipy's trust-region-reflective least squares with an analytic sparse Jacobian obtained by complex-step
    differentiation of the numpy model in oracle/dense_lm.py -- nothing shared with oracle/ba_oracle.c or the CUDA solver.
    Stored: the optimal cost and parameters (the inputs are regenerated from the seed)."""
    from scipy.sparse import csr_matrix
    from oracle import dense_lm
    p = synth.make_ba_problem(seed=0, **synth.BA_CONFIGS["cfg2"])
    nc, npt, nobs = p["nc"], p["np"], p["nobs"]
    oc, op, oxy = p["obs_cam"].astype(np.int64), p["obs_pt"].astype(np.int64), p["obs_xy"].astype(np.float64)
    n = 6 * nc + 3 * npt + 1

    def unpack(x):
        return x[:6 * nc].reshape(nc, 6), x[6 * nc:6 * nc + 3 * npt].reshape(npt, 3), x[-1]

    def fun(x):
        c, q, f = unpack(x)
        return dense_lm.residuals(c, q, f, oxy, oc, op).reshape(-1)

    rows = np.arange(2 * nobs).reshape(nobs, 2)

    def jac(x, h=1e-40):
        c, q, f = unpack(x)
        cc = c.astype(np.complex128); qc = q.astype(np.complex128)
        ri, ci, vv = [], [], []
        for k in range(6):
            c2 = cc.copy(); c2[:, k] += 1j * h
            d = dense_lm.residuals(c2, qc, f, oxy, oc, op).imag / h
            ri.append(rows.ravel()); ci.append(np.repeat(6 * oc + k, 2)); vv.append(d.ravel())
        for k in range(3):
            q2 = qc.copy(); q2[:, k] += 1j * h
            d = dense_lm.residuals(cc, q2, f, oxy, oc, op).imag / h
            ri.append(rows.ravel()); ci.append(np.repeat(6 * nc + 3 * op + k, 2)); vv.append(d.ravel())
        d = dense_lm.residuals(cc, qc, f + 1j * h, oxy, oc, op).imag / h
        ri.append(rows.ravel()); ci.append(np.full(2 * nobs, n - 1)); vv.append(d.ravel())
        return csr_matrix((np.concatenate(vv), (np.concatenate(ri), np.concatenate(ci))), shape=(2 * nobs, n))

    x0 = np.concatenate([p["cams"].ravel(), p["pts"].ravel(), [p["focal"]]])
    sol = least_squares(fun, x0, jac=jac, x_scale="jac", method="trf", ftol=1e-15, xtol=1e-15, gtol=1e-12, max_nfev=200, tr_solver="lsmr",
                        tr_options=dict(atol=1e-14, btol=1e-14, maxiter=4000))
    cost0 = 0.5 * np.sum(fun(x0) ** 2)
    g = jac(sol.x).T @ fun(sol.x)
    print("ba cfg2: cost0", cost0, "-> scipy optimum", sol.cost, "nfev", sol.nfev, "max|g|", np.abs(g).max(), sol.message)
    c, q, f = unpack(sol.x)
    np.savez_compressed(os.path.join(OUT, "ba_scipy_cfg2.npz"), cost0=cost0, cost_opt=sol.cost, cams=c, pts=q.astype(np.float32), focal=f,
                        grad_max=np.abs(g).max())
