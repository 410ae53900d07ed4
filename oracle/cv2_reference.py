"""oracle/cv2_reference.py -- TEST INFRASTRUCTURE ONLY.

Replays the reference's hot-path functions with the very OpenCV calls the reference makes, through the cv2
Python binding (cv2 4.13 is in this image; the reference needs OpenCV >= 3.1, CMakeLists.txt:28).  This is the
closest thing to "the reference run here" for matching and triangulation (the C++ reference cannot be built:
no C++ OpenCV/Ceres/Boost).  Used to pin oracle/*.c and to generate tests/golden/*.npz.
"""
import numpy as np

RATIO_REFERENCE = float(np.float64(np.float32(0.8)))


def match_features(desc_left, desc_right, norm="hamming"):
    """SfM2DFeatureUtilities::matchFeatures, SfM2DFeatureUtilities.cpp:53-71."""
    import cv2
    if desc_right.shape[0] < 2 or desc_left.shape[0] == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32)
    name = "BruteForce-Hamming" if norm == "hamming" else "BruteForce"
    matcher = cv2.DescriptorMatcher_create(name)                                  # :59
    knn = matcher.knnMatch(desc_left, desc_right, 2)                              # :60
    q, t, d = [], [], []
    for pair in knn:                                                              # :64-68
        if np.float64(np.float32(pair[0].distance)) < RATIO_REFERENCE * np.float64(np.float32(pair[1].distance)):
            q.append(pair[0].queryIdx); t.append(pair[0].trainIdx); d.append(pair[0].distance)
    return np.asarray(q, np.int32), np.asarray(t, np.int32), np.asarray(d, np.float32)


def triangulate_views(K, Pl, Pr, ptsL, ptsR, mq=None, mt=None, max_reproj=10.0):
    """SfMStereoUtilities::triangulateViews, SfMStereoUtilities.cpp:120-206.
    Returns X [m,3] float32 for ALL matches, keep mask [m], per-view reprojection error [m,2]."""
    import cv2
    K = np.asarray(K, np.float32).reshape(3, 3); Pl = np.asarray(Pl, np.float32).reshape(3, 4)
    Pr = np.asarray(Pr, np.float32).reshape(3, 4)
    ptsL = np.asarray(ptsL, np.float32).reshape(-1, 2); ptsR = np.asarray(ptsR, np.float32).reshape(-1, 2)
    if mq is not None:                                                            # GetAlignedPointsFromMatch, SfMCommon.cpp:63-87
        ptsL = ptsL[np.asarray(mq)]; ptsR = ptsR[np.asarray(mt)]
    m = ptsL.shape[0]
    if m == 0:
        return np.zeros((0, 3), np.float32), np.zeros(0, np.uint8), np.zeros((0, 2))
    nl = cv2.undistortPoints(ptsL.reshape(-1, 1, 2), K, None)                     # :146
    nr = cv2.undistortPoints(ptsR.reshape(-1, 1, 2), K, None)                     # :147
    X4 = cv2.triangulatePoints(Pl, Pr, nl.reshape(-1, 2).T.copy(), nr.reshape(-1, 2).T.copy())   # :150
    X = cv2.convertPointsFromHomogeneous(X4.T.copy()).reshape(-1, 3)              # :153
    err = np.zeros((m, 2))
    for v, (P, pts) in enumerate(((Pl, ptsL), (Pr, ptsR))):
        rvec, _ = cv2.Rodrigues(P[:, :3].copy())                                  # :155 / :162 (float in -> float out)
        tvec = P[:, 3].copy()
        proj, _ = cv2.projectPoints(X, rvec, tvec, K, None)                       # :159 / :166
        d = proj.reshape(-1, 2).astype(np.float32) - pts
        err[:, v] = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)   # cv::norm(Point2f)
    keep = ~((err[:, 0] > max_reproj) | (err[:, 1] > max_reproj))                 # :186-187
    return X.astype(np.float32), keep.astype(np.uint8), err
