"""oracle/ransac_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of how OpenCV's RANSAC scores ONE hypothesis, for the three RANSAC stages of the reference
(SfMStereoUtilities.cpp:51-72, :74-118, :208-243).  The arithmetic lives in un-vendored OpenCV (>= 3.1, CMakeLists.txt:28; this
image: cv2 4.13): RANSACPointSetRegistrator::findInliers (ptsetreg.cpp: inlier <=> err <= (float)(thresh*thresh)) over
  * HomographyEstimatorCallback::computeError (fundam.cpp): FLOAT transfer error with the model cast to float, h22 = 1;
  * EMEstimatorCallback::computeError (five-point.cpp): DOUBLE Sampson error on points normalised (x - pp)/focal, stored as float;
  * PnPRansacCallback::computeError (solvepnp.cpp): cv::projectPoints, float difference, squared L2 norm.
These callbacks are not exposed by the Python binding, so the restatement is pinned through what cv2 does expose (tests/
test_oracle_ransac.py): the mask cv2.findEssentialMat returns for the E it returns; cv2.projectPoints; cv2.perspectiveTransform.
"""
import numpy as np


def thresh2(threshold):
    return np.float32(np.float64(threshold) * np.float64(threshold))


def err_homography(H, a, b):
    H = np.asarray(H, np.float64).reshape(3, 3)
    if H[2, 2] != 0 and H[2, 2] != 1:
        H = H / H[2, 2]
    h = H.astype(np.float32).reshape(-1)
    a = np.asarray(a, np.float32).reshape(-1, 2); b = np.asarray(b, np.float32).reshape(-1, 2)
    x, y = a[:, 0], a[:, 1]
    ww = np.float32(1) / ((h[6] * x + h[7] * y) + np.float32(1))
    dx = ((h[0] * x + h[1] * y) + h[2]) * ww - b[:, 0]
    dy = ((h[3] * x + h[4] * y) + h[5]) * ww - b[:, 1]
    return (dx * dx + dy * dy).astype(np.float32)


def err_essential(E, a, b, focal, cx, cy):
    E = np.asarray(E, np.float64).reshape(3, 3)
    a = np.asarray(a, np.float32).astype(np.float64).reshape(-1, 2); b = np.asarray(b, np.float32).astype(np.float64).reshape(-1, 2)
    x1 = (a[:, 0] - cx) / focal; y1 = (a[:, 1] - cy) / focal; x2 = (b[:, 0] - cx) / focal; y2 = (b[:, 1] - cy) / focal
    e0 = (E[0, 0] * x1 + E[0, 1] * y1) + E[0, 2]; e1 = (E[1, 0] * x1 + E[1, 1] * y1) + E[1, 2]; e2 = (E[2, 0] * x1 + E[2, 1] * y1) + E[2, 2]
    t0 = (E[0, 0] * x2 + E[1, 0] * y2) + E[2, 0]; t1 = (E[0, 1] * x2 + E[1, 1] * y2) + E[2, 1]
    x2tEx1 = (x2 * e0 + y2 * e1) + e2
    den = ((e0 * e0 + e1 * e1) + t0 * t0) + t1 * t1
    return (x2tEx1 * x2tEx1 / den).astype(np.float32)


def err_pose(P, K, X, uv):
    P = np.asarray(P, np.float64).reshape(3, 4); K = np.asarray(K, np.float64).reshape(3, 3)
    X = np.asarray(X, np.float32).astype(np.float64).reshape(-1, 3); uv = np.asarray(uv, np.float32).reshape(-1, 2)
    p = X @ P[:, :3].T + P[:, 3]
    iz = np.where(p[:, 2] != 0, 1.0 / np.where(p[:, 2] != 0, p[:, 2], 1.0), 1.0)
    pu = (K[0, 0] * (p[:, 0] * iz) + K[0, 2]).astype(np.float32); pv = (K[1, 1] * (p[:, 1] * iz) + K[1, 2]).astype(np.float32)
    dx = uv[:, 0] - pu; dy = uv[:, 1] - pv
    return (dx * dx + dy * dy).astype(np.float32)


def score(model, a, b, hyps, aux, threshold):
    """-> (counts [nh], best index (most inliers, ties -> lowest), mask of the best)"""
    t2 = thresh2(threshold)
    errs = []
    for h in hyps:
        if model == 0:
            errs.append(err_homography(h, a, b))
        elif model == 1:
            errs.append(err_essential(h, a, b, aux[0], aux[1], aux[2]))
        else:
            errs.append(err_pose(h, aux, a, b))
    masks = [(e <= t2) for e in errs]
    counts = np.array([int(m.sum()) for m in masks], np.int32)
    best = int(np.argmax(counts)) if len(counts) else -1
    return counts, best, (masks[best].astype(np.uint8) if best >= 0 else np.zeros(len(np.asarray(b).reshape(-1, 2)), np.uint8))
