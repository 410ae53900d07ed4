"""oracle/cv2_stages.py -- TEST INFRASTRUCTURE ONLY.

The three hot-path stage functions with the reference's names and argument order, computed on the CPU the way the
reference computes them: matching and triangulation by the reference's own OpenCV calls through cv2
(oracle/cv2_reference.py), bundle adjustment by the C restatement of the Ceres solve (oracle/ba_oracle.c).
They plug into the driver mirror (sfm-toy-library_b200/runsfm.py) as the CPU arm of BASELINE configs[0]
(crazyhorse replay): tests/golden/make_cfg1.py, tests/test_gpu_cfg1.py, bench.py's cfg1 cpu leg.
The host-side marshalling (flatten / write-back, SfMBundleAdjustmentUtils.cpp:111-166, :188-221) is shared with the
product's stages.py on purpose -- it is container shuffling, the arithmetic is what differs between the two arms.
"""
import numpy as np

from sfm_toy_library_b200 import stages
from sfm_toy_library_b200.stages import DMATCH, Point3DInMap

from . import cv2_reference as ref
from . import oracle

CONVERGENCE = 0


def matchFeatures(featuresLeft, featuresRight):
    """SfM2DFeatureUtilities::matchFeatures (SfM2DFeatureUtilities.cpp:53-71) by cv2 knnMatch + the (double)0.8f ratio test."""
    q, t, d = ref.match_features(featuresLeft.descriptors, featuresRight.descriptors)
    m = np.zeros(len(q), DMATCH)
    m["queryIdx"] = q; m["trainIdx"] = t; m["distance"] = d
    return m


def matchFeatures_oracle(featuresLeft, featuresRight):
    """Same, by the plain-C popcount restatement (oracle/match_oracle.c)."""
    q, t, d = oracle.match_hamming(featuresLeft.descriptors, featuresRight.descriptors)
    m = np.zeros(len(q), DMATCH)
    m["queryIdx"] = q; m["trainIdx"] = t; m["distance"] = d
    return m


def triangulateViews(intrinsics, imagePair, matches, featuresLeft, featuresRight, Pleft, Pright, pointCloud):
    """SfMStereoUtilities::triangulateViews (SfMStereoUtilities.cpp:120-206) by the reference's six OpenCV calls."""
    mq = np.ascontiguousarray(matches["queryIdx"]); mt = np.ascontiguousarray(matches["trainIdx"])
    X, keep, _ = ref.triangulate_views(intrinsics.K, Pleft, Pright, featuresLeft.points, featuresRight.points, mq, mt)
    for i in np.nonzero(keep)[0]:
        pointCloud.append(Point3DInMap(X[i].copy(), {int(imagePair.left): int(mq[i]), int(imagePair.right): int(mt[i])}))
    return True


def adjustBundle(pointCloud, cameraPoses, intrinsics, image2dFeatures, options=None, num_threads=1):
    """SfMBundleAdjustmentUtils::adjustBundle (SfMBundleAdjustmentUtils.cpp:99-222) with ceres::Solve replaced by the
    restatement in ba_oracle.c (dual-number Jacobians, LM + DENSE_SCHUR, Ceres defaults, 500 iterations / 10 s)."""
    cams, pts, focal, obs_xy, obs_cam, pt_off, used = stages.flatten_bundle(pointCloud, cameraPoses, intrinsics, image2dFeatures,
                                                                             rot2aa=oracle.rotmat_to_angle_axis_f32)
    o = options or oracle.ba_default_options(jacobian_mode=0, num_threads=num_threads)
    cams, pts, focal, summary = oracle.ba_solve(cams, pts, focal, obs_xy, obs_cam, pt_off, o)
    if summary["termination_type"] != CONVERGENCE:
        return summary
    stages.write_back_bundle(pointCloud, cameraPoses, intrinsics, cams, pts, focal, used,
                             rot2aa=oracle.rotmat_to_angle_axis_f32, aa2rot=oracle.angle_axis_to_rotmat)
    return summary
