/*
 * oracle/match_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement of the reference's descriptor matching hot path:
 *   SfM2DFeatureUtilities::matchFeatures  (reference SfMToyLib/SfM2DFeatureUtilities.cpp:53-71)
 *     = cv::DescriptorMatcher("BruteForce-Hamming")->knnMatch(k=2)   (:59-60, un-vendored OpenCV >= 3.1)
 *     + ratio test  d0 < NN_MATCH_RATIO * d1  with NN_MATCH_RATIO = (double)0.8f  (:35, :65)
 *
 * OpenCV's brute-force knn (cv::batchDistance, K=2) keeps the K best in insertion-sorted
 * order with strict comparisons while scanning train rows in ascending index, i.e. ordering is
 * lexicographic on (distance, trainIdx): ties go to the lower trainIdx.
 *
 * Pinned against cv2 4.13 (the same library, Python binding) by tests/golden/match_*.npz
 * (generator: tests/golden/make_golden.py) and live in tests/test_oracle_match.py.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <float.h>

static inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

/* Hamming distance between two byte strings (any length; 8-byte chunks + tail). */
static int hamming_bytes(const uint8_t* a, const uint8_t* b, int nbytes) {
    int d = 0, i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        d += popc64(x ^ y);
    }
    for (; i < nbytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

/*
 * knn (k=2), Hamming.  best_idx/best_dist are [nq*2]; rows with fewer than 2 train rows get idx -1.
 * reference: SfM2DFeatureUtilities.cpp:60 (knnMatch(..., 2)).
 */
void sfm_oracle_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes,
                             int32_t* best_idx, int32_t* best_dist) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        int d0 = INT32_MAX, d1 = INT32_MAX, i0 = -1, i1 = -1;
        const uint8_t* qi = q + (size_t)i * desc_bytes;
        for (int j = 0; j < nt; ++j) {
            int d = hamming_bytes(qi, t + (size_t)j * desc_bytes, desc_bytes);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        best_idx[2 * i] = i0; best_idx[2 * i + 1] = i1;
        best_dist[2 * i] = i0 >= 0 ? d0 : -1;
        best_dist[2 * i + 1] = i1 >= 0 ? d1 : -1;
    }
}

/*
 * Full matchFeatures: knn2 + ratio test, output ascending queryIdx (reference :63-68).
 * ratio is passed as a double; the reference value is (double)0.8f = 0.800000011920928955.
 * nt < 2 is undefined behaviour in the reference (initialMatching[i][1] out of range); here: no matches.
 * Returns the number of surviving matches; out_* must hold nq entries.
 */
int sfm_oracle_match_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes, double ratio,
                             int32_t* out_q, int32_t* out_t, float* out_d,
                             int32_t* scratch_idx /* nq*2 */, int32_t* scratch_dist /* nq*2 */) {
    if (nt < 2 || nq <= 0) return 0;
    sfm_oracle_knn2_hamming(q, nq, t, nt, desc_bytes, scratch_idx, scratch_dist);
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        /* DMatch.distance is float; the comparison promotes to double (:65). */
        float d0 = (float)scratch_dist[2 * i], d1 = (float)scratch_dist[2 * i + 1];
        if ((double)d0 < ratio * (double)d1) {
            out_q[n] = i; out_t[n] = scratch_idx[2 * i]; out_d[n] = d0; ++n;
        }
    }
    return n;
}

/*
 * L2 variant (BASELINE.json config 4 wording, "SIFT-128"): cv::BFMatcher(NORM_L2).knnMatch(k=2).
 * OpenCV computes the squared distance in float32 (normL2Sqr_, 4-way unrolled float accumulation)
 * and takes sqrt in float.  For integer-valued descriptors (real SIFT is integer valued 0..255,
 * sums < 2^24) every partial sum is exact in float32, so any summation order gives the same bits.
 */
void sfm_oracle_knn2_l2(const float* q, int nq, const float* t, int nt, int dim,
                        int32_t* best_idx, float* best_dist /* sqrt'd, float */) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        float d0 = FLT_MAX, d1 = FLT_MAX; int i0 = -1, i1 = -1;
        const float* qi = q + (size_t)i * dim;
        for (int j = 0; j < nt; ++j) {
            const float* tj = t + (size_t)j * dim;
            float s = 0.f;
            for (int k = 0; k < dim; ++k) { float e = qi[k] - tj[k]; s += e * e; }
            if (s < d0) { d1 = d0; i1 = i0; d0 = s; i0 = j; }
            else if (s < d1) { d1 = s; i1 = j; }
        }
        best_idx[2 * i] = i0; best_idx[2 * i + 1] = i1;
        best_dist[2 * i] = i0 >= 0 ? __builtin_sqrtf(d0) : -1.f;
        best_dist[2 * i + 1] = i1 >= 0 ? __builtin_sqrtf(d1) : -1.f;
    }
}

int sfm_oracle_match_l2(const float* q, int nq, const float* t, int nt, int dim, double ratio,
                        int32_t* out_q, int32_t* out_t, float* out_d,
                        int32_t* scratch_idx, float* scratch_dist) {
    if (nt < 2 || nq <= 0) return 0;
    sfm_oracle_knn2_l2(q, nq, t, nt, dim, scratch_idx, scratch_dist);
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        float d0 = scratch_dist[2 * i], d1 = scratch_dist[2 * i + 1];
        if ((double)d0 < ratio * (double)d1) {
            out_q[n] = i; out_t[n] = scratch_idx[2 * i]; out_d[n] = d0; ++n;
        }
    }
    return n;
}
