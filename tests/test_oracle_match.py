"""Pins oracle/match_oracle.c against cv2 golden vectors (tests/golden/match_cv2.npz, made by make_golden.py with the
OpenCV calls of SfM2DFeatureUtilities.cpp:59-68) and, when cv2 is importable, against cv2 live."""
import numpy as np
import pytest

from sfm_toy_library_b200 import synth


@pytest.mark.parametrize("case", ["a", "b", "c", "d", "e"])
def test_match_hamming_golden(oracle, golden, case):
    g = golden("match_cv2.npz")
    q, t = g[f"{case}_q"], g[f"{case}_t"]
    idx, dist = oracle.knn2_hamming(q, t)
    np.testing.assert_array_equal(idx, g[f"{case}_knn_idx"])          # bit-exact, incl. lowest-trainIdx tie-break
    np.testing.assert_array_equal(dist.astype(np.float32), g[f"{case}_knn_dist"])
    mq, mt, md = oracle.match_hamming(q, t)
    np.testing.assert_array_equal(mq, g[f"{case}_mq"])
    np.testing.assert_array_equal(mt, g[f"{case}_mt"])
    np.testing.assert_array_equal(md, g[f"{case}_md"])


def test_ratio_constant_is_float_0p8(oracle, golden):
    """d0=4, d1=5: kept with (double)0.8f, dropped with exact 0.8."""
    g = golden("match_cv2.npz")
    assert len(oracle.match_hamming(g["c_q"], g["c_t"])[0]) == 1
    assert len(oracle.match_hamming(g["c_q"], g["c_t"], ratio=0.8)[0]) == 0


def test_match_l2_golden(oracle, golden):
    g = golden("match_cv2.npz")
    mq, mt, md = oracle.match_l2(g["l2_q"], g["l2_t"])
    np.testing.assert_array_equal(mq, g["l2_mq"]); np.testing.assert_array_equal(mt, g["l2_mt"])
    np.testing.assert_array_equal(md, g["l2_md"])


def test_match_edge_cases(oracle):
    d = synth.make_descriptors(0, 10)
    assert len(oracle.match_hamming(d, d[:1])[0]) == 0          # nt < 2: undefined in the reference -> empty
    assert len(oracle.match_hamming(d[:0], d)[0]) == 0          # empty query


def test_match_hamming_vs_cv2_live(oracle):
    ref = pytest.importorskip("cv2") and __import__("oracle.cv2_reference", fromlist=["x"])
    a = synth.make_descriptors(3, 700); b = synth.make_descriptors(4, 650, prev=a)
    for q, t in ((b, a), (a, b)):
        mq, mt, md = oracle.match_hamming(q, t)
        rq, rt, rd = ref.match_features(q, t)
        np.testing.assert_array_equal(mq, rq); np.testing.assert_array_equal(mt, rt); np.testing.assert_array_equal(md, rd)
        assert len(mq) > 20


def test_match_l2_real_sift_golden(oracle, golden):
    """Real SIFT descriptors of two crazyhorse images vs cv2.BFMatcher(NORM_L2) + ratio test (tests/golden/make_cfg1.py:sift_golden)."""
    g = golden("sift_crazyhorse.npz")
    q = g["q"].astype(np.float32); t = g["t"].astype(np.float32)
    for got, want in ((oracle.match_l2(q, t), ("mq", "mt", "md")), (oracle.match_l2(t, q), ("rq", "rt", "rd"))):
        for a, k in zip(got, want):
            np.testing.assert_array_equal(a, g[k])
