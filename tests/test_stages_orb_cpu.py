"""Host mirror of the f-3 stage (stages.extractFeatures / extractAllFeatures, runsfm.SfM.from_images) with a stand-in context: container
shapes, KeyPointsToPoints, batching rule, driver hook.  The arithmetic is tested elsewhere (test_oracle_orb / test_gpu_orb)."""
import numpy as np

from sfm_toy_library_b200 import runsfm, stages


class FakeCtx:
    def __init__(self):
        self.calls = []

    def orb_detect_and_compute(self, images, nfeatures=5000, capacity=None):
        single = isinstance(images, np.ndarray)
        imgs = [images] if single else list(images)
        self.calls.append((len(imgs), nfeatures))
        out = []
        for im in imgs:
            n = int(im[0, 0]) if im.ndim == 2 else int(im[0, 0, 0])
            kp = np.zeros((n, 7), np.float32); kp[:, 0] = np.arange(n); kp[:, 1] = 2 * np.arange(n); kp[:, 6] = -1
            out.append((kp, np.full((n, 32), n, np.uint8)))
        return out[0] if single else out


def test_extract_features_builds_the_reference_container():
    ctx = FakeCtx()
    img = np.full((16, 16, 3), 5, np.uint8)
    f = stages.extractFeatures(img, ctx=ctx)
    assert ctx.calls == [(1, 5000)]                                        # ORB::create(5000), SfM2DFeatureUtilities.cpp:39
    assert f.points.shape == (5, 2) and f.points.dtype == np.float32 and f.descriptors.shape == (5, 32)
    assert np.array_equal(f.points, f.keyPoints[:, :2])                    # KeyPointsToPoints, SfMCommon.cpp:89-94


def test_extract_all_batches_equal_sizes_and_splits_unequal_ones():
    ctx = FakeCtx()
    same = [np.full((8, 8), i + 1, np.uint8) for i in range(4)]
    feats = stages.extractAllFeatures(same, ctx=ctx)
    assert ctx.calls == [(4, 5000)] and [len(f.points) for f in feats] == [1, 2, 3, 4]
    ctx = FakeCtx()
    mixed = [np.full((8, 8), 2, np.uint8), np.full((8, 9), 3, np.uint8)]
    feats = stages.extractAllFeatures(mixed, ctx=ctx)
    assert ctx.calls == [(1, 5000), (1, 5000)] and [len(f.points) for f in feats] == [2, 3]


def test_driver_mirror_from_images_uses_the_injected_stage():
    ctx = FakeCtx()
    imgs = [np.full((768, 1024, 3), 3, np.uint8), np.full((768, 1024, 3), 4, np.uint8)]
    sfm = runsfm.SfM.from_images(imgs, extractAllFeatures=lambda ims: stages.extractAllFeatures(ims, ctx=ctx))
    assert sfm.n == 2 and sfm.calls["extract"] == 2 and sfm.seconds["extract"] >= 0
    assert sfm.mIntrinsics.K[0, 2] == 512 and sfm.mIntrinsics.K[1, 2] == 384      # SfM.cpp:70-72 from the image size
    assert [len(f.points) for f in sfm.mImageFeatures] == [3, 4]
