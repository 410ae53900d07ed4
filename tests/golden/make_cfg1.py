"""Generates tests/golden/cfg1_crazyhorse.npz: BASELINE configs[0] (the reference's only real-data configuration) replayed
stage by stage with the reference's own third-party calls, run HERE (the GPU box has no /root/reference):

  * ORB(5000) features of the 7 crazyhorse JPEGs, sorted by file name (SfM2DFeatureUtilities.cpp:37-51; SURVEY.md 8c:
    directory_iterator order is unspecified, so the harness fixes it) -- cv2.ORB_create(5000).detectAndCompute;
  * all 21 pairs through cv2 knnMatch + the (double)0.8f ratio test (SfM2DFeatureUtilities.cpp:53-71);
  * a full SfM::runSfM replay (sfm-toy-library_b200/runsfm.py driver, cv2 RANSAC stages, cv2 triangulation chain, the
    oracle's Ceres restatement for adjustBundle), with the inputs and outputs of EVERY triangulateViews and adjustBundle
    call recorded.

The images themselves are not committed; the descriptors/keypoints are (1.4 MB).
Run from the repo root:  python tests/golden/make_cfg1.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from oracle import cv2_stages  # noqa: E402
from sfm_toy_library_b200 import runsfm  # noqa: E402
from sfm_toy_library_b200.stages import Features  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DATASET = "/root/reference/dataset/crazyhorse"


def extract():
    files = sorted(glob.glob(os.path.join(DATASET, "*.JPG")))
    assert len(files) == 7, files
    feats, size = [], None
    for f in files:
        img = cv2.imread(f)                                           # SfM.cpp:124 (downscale factor 1.0)
        size = (img.shape[1], img.shape[0])
        orb = cv2.ORB_create(5000)                                    # SfM2DFeatureUtilities.cpp:39
        kps, desc = orb.detectAndCompute(img, None)                   # :48
        pts = np.array([k.pt for k in kps], np.float32)               # KeyPointsToPoints, SfMCommon.cpp:89-94
        feats.append(Features(points=pts, descriptors=desc))
    return files, feats, size


def main():
    files, feats, size = extract()
    trace = []
    sfm = runsfm.SfM(feats, size, matchFeatures=cv2_stages.matchFeatures, triangulateViews=cv2_stages.triangulateViews,
                     adjustBundle=cv2_stages.adjustBundle, trace=trace, verbose=True)
    sfm.runSfM()
    out = {"files": np.array([os.path.basename(f) for f in files]), "image_size": np.array(size, np.int32),
           "cv2_version": np.array(cv2.__version__)}
    for i, f in enumerate(feats):
        out[f"pts_{i}"] = f.points; out[f"desc_{i}"] = f.descriptors
    n_tri = n_ba = n_pnp = 0
    for ev in trace:
        if ev["stage"] == "match":
            out["pairs"] = np.array(ev["pairs"], np.int32)
            for p, m in enumerate(ev["matches"]):
                out[f"match_{p}_q"] = m["queryIdx"]; out[f"match_{p}_t"] = m["trainIdx"]; out[f"match_{p}_d"] = m["distance"]
        elif ev["stage"] == "triangulate":
            k = f"tri_{n_tri}_"; n_tri += 1
            out[k + "pair"] = np.array(ev["pair"], np.int32); out[k + "K"] = ev["K"]; out[k + "Pl"] = ev["Pl"]; out[k + "Pr"] = ev["Pr"]
            out[k + "mq"] = ev["matches"]["queryIdx"]; out[k + "mt"] = ev["matches"]["trainIdx"]
            out[k + "X"] = ev["X"]; out[k + "back"] = ev["back"]
        elif ev["stage"] == "bundle":
            k = f"ba_{n_ba}_"; n_ba += 1
            for name in ("cams", "pts", "obs_xy", "obs_cam", "pt_off", "used", "K_after", "poses_after", "pts_after"):
                out[k + name] = ev[name]
            out[k + "focal"] = np.array(ev["focal"])
            s = ev["summary"]
            out[k + "summary"] = np.array([s["termination_type"], s["num_iterations"], s["num_successful_steps"],
                                           s["num_unsuccessful_steps"]], np.int32)
            out[k + "cost"] = np.array([s["initial_cost"], s["final_cost"]])
        elif ev["stage"] == "pnp":
            k = f"pnp_{n_pnp}_"; n_pnp += 1
            out[k + "view"] = np.array(ev["view"], np.int32); out[k + "p2"] = ev["points2D"]; out[k + "p3"] = ev["points3D"]
            out[k + "ok"] = np.array(ev["ok"]); out[k + "pose"] = ev["pose"] if ev["pose"] is not None else np.zeros((3, 4), np.float32)
    out["n_tri"] = np.array(n_tri); out["n_ba"] = np.array(n_ba); out["n_pnp"] = np.array(n_pnp)
    out["final_cloud"] = np.array([p.p for p in sfm.mReconstructionCloud], np.float32)
    out["final_poses"] = np.array(sfm.mCameraPoses, np.float32)
    out["final_K"] = sfm.mIntrinsics.K
    np.savez_compressed(os.path.join(OUT, "cfg1_crazyhorse.npz"), **out)
    cnt = [len(out[f"match_{p}_q"]) for p in range(len(out["pairs"]))]
    print("pairs:", cnt)
    print("triangulate calls:", n_tri, "bundle calls:", n_ba, "pnp calls:", n_pnp, "final cloud:", len(sfm.mReconstructionCloud))
    print("seconds:", {k: round(v, 3) for k, v in sfm.seconds.items()})
    for ev in trace:
        if ev["stage"] == "bundle":
            s = ev["summary"]
            print("  BA nc=%d np=%d nobs=%d: %s it=%d cost %.4f -> %.4f" % (len(ev["used"]), len(ev["pts"]), len(ev["obs_cam"]),
                  s["message"], s["num_iterations"], s["initial_cost"], s["final_cost"]))


def sift_golden():
    """Real SIFT descriptors of two crazyhorse images (the reference's legacy matcher, legacy/SfMToyLib_Old/GPUSURFFeatureMatcher.cpp:
    100-124, is an L2 knn + ratio test; BASELINE.json configs[3] says SIFT-128) + cv2.BFMatcher(NORM_L2) knnMatch with the ratio test."""
    from oracle import cv2_reference as ref
    files = sorted(glob.glob(os.path.join(DATASET, "*.JPG")))
    descs = []
    for f in files[:2]:
        img = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        kps, d = cv2.SIFT_create(3000).detectAndCompute(img, None)
        assert np.array_equal(d, np.round(d)) and d.min() >= 0 and d.max() <= 255       # SIFT output is u8-valued
        descs.append(d[:3000].astype(np.float32))
    q, t = descs
    mq, mt, md = ref.match_features(q, t, norm="l2"); rq, rt, rd = ref.match_features(t, q, norm="l2")
    np.savez_compressed(os.path.join(OUT, "sift_crazyhorse.npz"), q=q.astype(np.uint8), t=t.astype(np.uint8), mq=mq, mt=mt, md=md, rq=rq, rt=rt, rd=rd)
    print("sift:", q.shape, t.shape, "survivors", len(mq), len(rq))


if __name__ == "__main__":
    main()
    sift_golden()
