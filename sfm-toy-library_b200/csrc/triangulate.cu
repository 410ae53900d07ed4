// triangulate.cu -- K2: batched two-view DLT triangulation with fused gather / normalise / reproject / filter.
//
// Replaces SfMStereoUtilities::triangulateViews (reference SfMToyLib/SfMStereoUtilities.cpp:120-206) and the gather of
// GetAlignedPointsFromMatch (reference SfMToyLib/SfMCommon.cpp:63-87), which the reference executes as six OpenCV calls:
//   undistortPoints x2 (:146-147) -> triangulatePoints (:150) -> convertPointsFromHomogeneous (:153)
//   -> Rodrigues + projectPoints x2 (:155-167) -> 10 px reprojection filter in either view (:184-203).
// One thread per match does all of it in registers: indexed load of the two keypoints (the descriptor-row copies of
// SfMCommon.cpp:78,80 are dead work and are dropped), x_n = (u - c) * (1/f) in double rounded to float (what
// cv::undistortPoints returns), the 4x4 DLT matrix in double, its smallest right singular vector by one-sided
// (Hestenes) Jacobi in double -- the algorithm cv::SVD uses -- float dehomogenisation with a float reciprocal,
// double projection with the float-rvec round-tripped rotation, float pixel error, keep flag.
// HBM traffic: 16 B in (+8 B indices) and 13 B out per match; everything else lives in registers.
#include "common.cuh"
#include <cfloat>
#include <cmath>

namespace {

struct TriParams {
    double PL[12], PR[12];     // float poses widened (DLT rows)
    double RtL[12], RtR[12];   // [R|t] as cv::projectPoints sees them (float rvec round trip)
    double fx, fy, cx, cy, ifx, ify;
    float max_reproj;
};

// One Hestenes rotation between columns I and J of A (4 rows), accumulating V.
#define SFM_JROT(I, J)                                                                                            \
    {                                                                                                             \
        const double p0 = A[0][I] * A[0][J] + A[1][I] * A[1][J] + A[2][I] * A[2][J] + A[3][I] * A[3][J];          \
        const double a = nrm[I], b = nrm[J];                                                                      \
        if (fabs(p0) > eps * sqrt(a * b)) {                                                                       \
            const double p = 2.0 * p0, beta = a - b, gamma = hypot(p, beta);                                      \
            double c, s;                                                                                          \
            if (beta < 0) { const double delta = (gamma - beta) * 0.5; s = sqrt(delta / gamma); c = p / (gamma * s * 2.0); } \
            else { c = sqrt((gamma + beta) / (gamma * 2.0)); s = p / (gamma * c * 2.0); }                         \
            double na = 0, nb = 0;                                                                                \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                                      \
                const double t0 = c * A[k][I] + s * A[k][J], t1 = -s * A[k][I] + c * A[k][J];                     \
                A[k][I] = t0; A[k][J] = t1; na = fma(t0, t0, na); nb = fma(t1, t1, nb);                           \
                const double v0 = c * V[k][I] + s * V[k][J], v1 = -s * V[k][I] + c * V[k][J];                     \
                V[k][I] = v0; V[k][J] = v1;                                                                       \
            }                                                                                                     \
            nrm[I] = na; nrm[J] = nb; rotated = true;                                                             \
        }                                                                                                         \
    }

__device__ __forceinline__ void null_vector_4x4(double A[4][4], double X[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double nrm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) nrm[j] = A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j] + A[3][j] * A[3][j];
    const double eps = DBL_EPSILON * 10;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        SFM_JROT(0, 1) SFM_JROT(0, 2) SFM_JROT(0, 3) SFM_JROT(1, 2) SFM_JROT(1, 3) SFM_JROT(2, 3)
        if (!rotated) break;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) nrm[j] = A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j] + A[3][j] * A[3][j];
    // smallest singular value; among equal norms the last column (cv::SVD sorts descending, takes the last row of Vt)
    double best = nrm[0];
#pragma unroll
    for (int k = 0; k < 4; ++k) X[k] = V[k][0];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        if (nrm[j] <= best) {
            best = nrm[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) X[k] = V[k][j];
        }
    }
}

// ---- fast path for the smallest right singular vector --------------------------------------------------------------
// For a well-conditioned two-view DLT the 4x4 matrix B = A^T A has one eigenvalue far below the other three.  Its
// characteristic polynomial p(x) = x^4 - c3 x^3 + c2 x^2 - c1 x + c0 has only real roots, so Newton's iteration started
// at 0 climbs monotonically to the smallest one; the eigenvector is then a column of adj(B - x I) (rank-3 matrix: every
// column of the adjugate is a multiple of the null vector; the column with the largest diagonal cofactor is used).
// ~600 flops instead of ~5000 for the Jacobi SVD.  The result is verified (residual of the eigen-equation); anything
// suspicious falls back to the Jacobi routine above, which is the reference-faithful algorithm.
__device__ __forceinline__ double det3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}
struct Sym4 { double c00, c01, c02, c03, c11, c12, c13, c22, c23, c33; };
struct Adj4 { double k00, k01, k02, k03, k11, k12, k13, k22, k23, k33; };
__device__ __forceinline__ void adjugate_sym4(const Sym4& m, Adj4& k) {
    k.k00 = det3(m.c11, m.c12, m.c13, m.c12, m.c22, m.c23, m.c13, m.c23, m.c33);
    k.k01 = -det3(m.c01, m.c12, m.c13, m.c02, m.c22, m.c23, m.c03, m.c23, m.c33);
    k.k02 = det3(m.c01, m.c11, m.c13, m.c02, m.c12, m.c23, m.c03, m.c13, m.c33);
    k.k03 = -det3(m.c01, m.c11, m.c12, m.c02, m.c12, m.c22, m.c03, m.c13, m.c23);
    k.k11 = det3(m.c00, m.c02, m.c03, m.c02, m.c22, m.c23, m.c03, m.c23, m.c33);
    k.k12 = -det3(m.c00, m.c01, m.c03, m.c02, m.c12, m.c23, m.c03, m.c13, m.c33);
    k.k13 = det3(m.c00, m.c01, m.c02, m.c02, m.c12, m.c22, m.c03, m.c13, m.c23);
    k.k22 = det3(m.c00, m.c01, m.c03, m.c01, m.c11, m.c13, m.c03, m.c13, m.c33);
    k.k23 = -det3(m.c00, m.c01, m.c02, m.c01, m.c11, m.c12, m.c03, m.c13, m.c23);
    k.k33 = det3(m.c00, m.c01, m.c02, m.c01, m.c11, m.c12, m.c02, m.c12, m.c22);
}
__device__ __forceinline__ bool null_vector_fast(const double A[4][4], double X[4]) {
    Sym4 B;
    B.c00 = A[0][0] * A[0][0] + A[1][0] * A[1][0] + A[2][0] * A[2][0] + A[3][0] * A[3][0];
    B.c01 = A[0][0] * A[0][1] + A[1][0] * A[1][1] + A[2][0] * A[2][1] + A[3][0] * A[3][1];
    B.c02 = A[0][0] * A[0][2] + A[1][0] * A[1][2] + A[2][0] * A[2][2] + A[3][0] * A[3][2];
    B.c03 = A[0][0] * A[0][3] + A[1][0] * A[1][3] + A[2][0] * A[2][3] + A[3][0] * A[3][3];
    B.c11 = A[0][1] * A[0][1] + A[1][1] * A[1][1] + A[2][1] * A[2][1] + A[3][1] * A[3][1];
    B.c12 = A[0][1] * A[0][2] + A[1][1] * A[1][2] + A[2][1] * A[2][2] + A[3][1] * A[3][2];
    B.c13 = A[0][1] * A[0][3] + A[1][1] * A[1][3] + A[2][1] * A[2][3] + A[3][1] * A[3][3];
    B.c22 = A[0][2] * A[0][2] + A[1][2] * A[1][2] + A[2][2] * A[2][2] + A[3][2] * A[3][2];
    B.c23 = A[0][2] * A[0][3] + A[1][2] * A[1][3] + A[2][2] * A[2][3] + A[3][2] * A[3][3];
    B.c33 = A[0][3] * A[0][3] + A[1][3] * A[1][3] + A[2][3] * A[2][3] + A[3][3] * A[3][3];
    const double c3 = B.c00 + B.c11 + B.c22 + B.c33;
    if (!(c3 > 0.0) || !isfinite(c3)) return false;
    Adj4 K;
    adjugate_sym4(B, K);
    const double c1 = K.k00 + K.k11 + K.k22 + K.k33;
    const double c0 = B.c00 * K.k00 + B.c01 * K.k01 + B.c02 * K.k02 + B.c03 * K.k03;
    const double c2 = (B.c00 * B.c11 - B.c01 * B.c01) + (B.c00 * B.c22 - B.c02 * B.c02) + (B.c00 * B.c33 - B.c03 * B.c03) +
                      (B.c11 * B.c22 - B.c12 * B.c12) + (B.c11 * B.c33 - B.c13 * B.c13) + (B.c22 * B.c33 - B.c23 * B.c23);
    if (!(c1 > 0.0)) return false;
    double x = 0.0;
    bool conv = false;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const double p = (((x - c3) * x + c2) * x - c1) * x + c0;
        const double dp = ((4.0 * x - 3.0 * c3) * x + 2.0 * c2) * x - c1;
        if (!(dp < 0.0)) return false;
        const double step = p / dp;                   // <= 0 while below the root
        x -= step;
        if (fabs(step) <= 1e-15 * c3) { conv = true; break; }
    }
    if (!conv || !(x >= -1e-12 * c3)) return false;
    Sym4 C = B;
    C.c00 -= x; C.c11 -= x; C.c22 -= x; C.c33 -= x;
    adjugate_sym4(C, K);
    const double a0 = fabs(K.k00), a1 = fabs(K.k11), a2 = fabs(K.k22), a3 = fabs(K.k33);
    double v0, v1, v2, v3;
    if (a0 >= a1 && a0 >= a2 && a0 >= a3) { v0 = K.k00; v1 = K.k01; v2 = K.k02; v3 = K.k03; }
    else if (a1 >= a2 && a1 >= a3) { v0 = K.k01; v1 = K.k11; v2 = K.k12; v3 = K.k13; }
    else if (a2 >= a3) { v0 = K.k02; v1 = K.k12; v2 = K.k22; v3 = K.k23; }
    else { v0 = K.k03; v1 = K.k13; v2 = K.k23; v3 = K.k33; }
    const double vn = fmax(fmax(fabs(v0), fabs(v1)), fmax(fabs(v2), fabs(v3)));
    if (!(vn > 0.0) || !isfinite(vn)) return false;
    // verify (B - x I) v = 0 and that x is well separated from the next eigenvalue: p'(x) = -prod_{i<4}(lambda_i - x)
    const double r0 = C.c00 * v0 + C.c01 * v1 + C.c02 * v2 + C.c03 * v3, r1 = C.c01 * v0 + C.c11 * v1 + C.c12 * v2 + C.c13 * v3;
    const double r2 = C.c02 * v0 + C.c12 * v1 + C.c22 * v2 + C.c23 * v3, r3 = C.c03 * v0 + C.c13 * v1 + C.c23 * v2 + C.c33 * v3;
    const double rn = fmax(fmax(fabs(r0), fabs(r1)), fmax(fabs(r2), fabs(r3)));
    const double dpx = ((4.0 * x - 3.0 * c3) * x + 2.0 * c2) * x - c1;
    // gap estimate: |p'(x)| >= (lambda_3 - x) * (c3/4)^2-ish; demand lambda_3 - x >= 1e-7 * c3 (conservative), residual tiny
    if (!(-dpx >= 1e-7 * c3 * c3 * c3 * (1.0 / 64.0)) || !(rn <= 1e-11 * c3 * vn)) return false;
    X[0] = v0; X[1] = v1; X[2] = v2; X[3] = v3;
    return true;
}

__global__ void __launch_bounds__(128)
triangulate_kernel(TriParams P, const float2* __restrict__ ptsL, const float2* __restrict__ ptsR,
                   const int32_t* __restrict__ mq, const int32_t* __restrict__ mt, int m,
                   float* __restrict__ X, uint8_t* __restrict__ keep, int32_t* __restrict__ n_keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0;
    if (i < m) {
        const int iq = mq ? mq[i] : i, it = mt ? mt[i] : i;
        const float2 l = __ldg(ptsL + iq), r = __ldg(ptsR + it);
        // cv::undistortPoints without distortion: (u - c) * (1/f) in double, returned as float
        const double xl = (double)(float)(((double)l.x - P.cx) * P.ifx), yl = (double)(float)(((double)l.y - P.cy) * P.ify);
        const double xr = (double)(float)(((double)r.x - P.cx) * P.ifx), yr = (double)(float)(((double)r.y - P.cy) * P.ify);
        double A[4][4], Xh[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            A[0][c] = xl * P.PL[8 + c] - P.PL[c];
            A[1][c] = yl * P.PL[8 + c] - P.PL[4 + c];
            A[2][c] = xr * P.PR[8 + c] - P.PR[c];
            A[3][c] = yr * P.PR[8 + c] - P.PR[4 + c];
        }
        if (!null_vector_fast(A, Xh)) null_vector_4x4(A, Xh);       // Jacobi SVD when the closed-form path declines
        // triangulatePoints returns float; convertPointsFromHomogeneous: scale = 1/w in float (w == 0 -> 1)
        const float hx = (float)Xh[0], hy = (float)Xh[1], hz = (float)Xh[2], hw = (float)Xh[3];
        const float sc = hw != 0.f ? __frcp_rn(hw) : 1.f;
        const float px = hx * sc, py = hy * sc, pz = hz * sc;
        X[3 * (size_t)i] = px; X[3 * (size_t)i + 1] = py; X[3 * (size_t)i + 2] = pz;
        // projectPoints in double, float pixel out; error norm in double of float differences
        double e2[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const double* Rt = v == 0 ? P.RtL : P.RtR;
            const double Xc = Rt[0] * px + Rt[1] * py + Rt[2] * pz + Rt[3];
            const double Yc = Rt[4] * px + Rt[5] * py + Rt[6] * pz + Rt[7];
            double Zc = Rt[8] * px + Rt[9] * py + Rt[10] * pz + Rt[11];
            Zc = Zc != 0.0 ? 1.0 / Zc : 1.0;
            const float pu = (float)(Xc * Zc * P.fx + P.cx), pv = (float)(Yc * Zc * P.fy + P.cy);
            const float du = pu - (v == 0 ? l.x : r.x), dv = pv - (v == 0 ? l.y : r.y);
            e2[v] = sqrt((double)du * du + (double)dv * dv);
        }
        k = !(e2[0] > (double)P.max_reproj || e2[1] > (double)P.max_reproj);     // :186-187 (NaN compares false -> kept)
        keep[i] = (uint8_t)k;
    }
    // survivors: warp ballot -> one atomic per warp
    const unsigned b = __ballot_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_keep, __popc(b));
}

// --- host: cv::Rodrigues(R) -> float rvec -> rotation matrix, i.e. the [R|t] projectPoints really uses (:155-167) ---
void rotmat_to_rvec(const double R[9], double r[3]) {
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double theta = std::acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = std::sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; ry = std::sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1.0 : 1.0);
        t = (R[8] + 1) * 0.5; rz = std::sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1.0 : 1.0);
        if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        const double n = theta / std::sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * n; r[1] = ry * n; r[2] = rz * n;
        return;
    }
    const double vth = theta / (2.0 * s);
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

void pose_roundtrip(const float P[12], double Rt[12]) {
    double R[9], r[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = (double)P[4 * i + j];
    rotmat_to_rvec(R, r);
    for (int i = 0; i < 3; ++i) r[i] = (double)(float)r[i];          // the rvec Mat is CV_32F
    const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    double R2[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (theta >= DBL_EPSILON) {
        const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, x = r[0] / theta, y = r[1] / theta, z = r[2] / theta;
        R2[0] = c + c1 * x * x;     R2[1] = c1 * x * y - s * z; R2[2] = c1 * x * z + s * y;
        R2[3] = c1 * x * y + s * z; R2[4] = c + c1 * y * y;     R2[5] = c1 * y * z - s * x;
        R2[6] = c1 * x * z - s * y; R2[7] = c1 * y * z + s * x; R2[8] = c + c1 * z * z;
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rt[4 * i + j] = R2[3 * i + j];
        Rt[4 * i + 3] = (double)P[4 * i + 3];
    }
}

void make_params(const float* K, const float* Pl, const float* Pr, float max_reproj, TriParams& P) {
    for (int i = 0; i < 12; ++i) { P.PL[i] = (double)Pl[i]; P.PR[i] = (double)Pr[i]; }
    pose_roundtrip(Pl, P.RtL); pose_roundtrip(Pr, P.RtR);
    P.fx = K[0]; P.fy = K[4]; P.cx = K[2]; P.cy = K[5];
    P.ifx = 1.0 / P.fx; P.ify = 1.0 / P.fy;
    P.max_reproj = max_reproj;
}

}  // namespace

static int triangulate_launch(sfmb200_ctx* ctx, const TriParams& P, const float* d_l, const float* d_r, const int32_t* d_mq,
                              const int32_t* d_mt, int m, float* d_X, uint8_t* d_keep, int32_t* d_n) {
    SFM_CUDA(ctx, cudaMemsetAsync(d_n, 0, sizeof(int32_t), ctx->stream));
    if (m == 0) return SFMB200_OK;
    triangulate_kernel<<<ceil_div(m, 128), 128, 0, ctx->stream>>>(P, (const float2*)d_l, (const float2*)d_r, d_mq, d_mt, m, d_X, d_keep, d_n);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}

extern "C" {

int sfmb200_triangulate_device(sfmb200_ctx* ctx, const float* K, const float* Pl, const float* Pr, const float* d_l, const float* d_r,
                               const int32_t* d_mq, const int32_t* d_mt, int m, float max_reproj, float* d_X, uint8_t* d_keep, int32_t* d_n) {
    if (!ctx || !K || !Pl || !Pr || m < 0 || !d_n) return SFMB200_ERR_INVALID;
    if ((d_mq == nullptr) != (d_mt == nullptr)) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "match_q and match_t must both be given or both be NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    SFM_CUDA(ctx, cudaSetDevice(ctx->device));
    TriParams P; make_params(K, Pl, Pr, max_reproj, P);
    return triangulate_launch(ctx, P, d_l, d_r, d_mq, d_mt, m, d_X, d_keep, d_n);
}

int sfmb200_triangulate(sfmb200_ctx* ctx, const float* K, const float* Pl, const float* Pr, const float* pts_left, int n_left,
                        const float* pts_right, int n_right, const int32_t* match_q, const int32_t* match_t, int m, float max_reproj,
                        float* X, uint8_t* keep, int* n_keep) {
    if (!ctx || !K || !Pl || !Pr || m < 0 || n_left < 0 || n_right < 0) return SFMB200_ERR_INVALID;
    if (n_keep) *n_keep = 0;
    if (m == 0) return SFMB200_OK;
    if (!pts_left || !pts_right || !X || !keep) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "null buffer");
    if ((match_q == nullptr) != (match_t == nullptr)) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "match_q and match_t must both be given or both be NULL");
    if (!match_q && (m > n_left || m > n_right)) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "identity alignment needs m <= min(n_left, n_right)");
    if (match_q)
        for (int i = 0; i < m; ++i)
            if (match_q[i] < 0 || match_q[i] >= n_left || match_t[i] < 0 || match_t[i] >= n_right)
                return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "match %d indexes outside the keypoint arrays", i);
    std::lock_guard<std::mutex> lk(ctx->mu);
    SFM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = Carver::pad(8 * (size_t)n_left) + Carver::pad(8 * (size_t)n_right) + 2 * Carver::pad(4 * (size_t)m) +
                         Carver::pad(12 * (size_t)m) + Carver::pad(m) + 1024;
    SFM_CUDA(ctx, ctx->scratch.reserve(bytes));
    Carver cv(ctx->scratch.p);
    float* d_l = cv.take<float>(2 * (size_t)n_left); float* d_r = cv.take<float>(2 * (size_t)n_right);
    int32_t* d_mq = cv.take<int32_t>(m); int32_t* d_mt = cv.take<int32_t>(m);
    float* d_X = cv.take<float>(3 * (size_t)m); uint8_t* d_keep = cv.take<uint8_t>(m); int32_t* d_n = cv.take<int32_t>(1);
    SFM_CUDA(ctx, cudaMemcpyAsync(d_l, pts_left, 8 * (size_t)n_left, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_r, pts_right, 8 * (size_t)n_right, cudaMemcpyHostToDevice, ctx->stream));
    if (match_q) {
        SFM_CUDA(ctx, cudaMemcpyAsync(d_mq, match_q, 4 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        SFM_CUDA(ctx, cudaMemcpyAsync(d_mt, match_t, 4 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    }
    TriParams P; make_params(K, Pl, Pr, max_reproj, P);
    int rc = triangulate_launch(ctx, P, d_l, d_r, match_q ? d_mq : nullptr, match_q ? d_mt : nullptr, m, d_X, d_keep, d_n);
    if (rc) return rc;
    int32_t hn = 0;
    SFM_CUDA(ctx, cudaMemcpyAsync(X, d_X, 12 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(keep, d_keep, (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(&hn, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_keep) *n_keep = hn;
    return SFMB200_OK;
}

}  // extern "C"
