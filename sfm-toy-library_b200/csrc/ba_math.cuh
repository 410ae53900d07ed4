// ba_math.cuh -- per-observation arithmetic of the bundle-adjustment kernels (ba.cu).
//
// Model: SimpleReprojectionError (reference SfMToyLib/SfMBundleAdjustmentUtils.cpp:58-97):
//   p = AngleAxisRotatePoint(cam[0..2], X) + cam[3..5];  r = focal * p.xy / p.z - observed
// The reference differentiates this with ceres::AutoDiffCostFunction<...,2,6,3,1> (:91-94); here the derivative
// of exactly the same expression is written out in closed form (no dual numbers), with everything that depends on
// the camera only (sin, cos, 1/theta, unit axis, rotation matrix) hoisted into CamDerived, computed once per camera
// per evaluation point instead of once per observation.
// __host__ __device__ so that tests/ can run the very same code on the CPU against the oracle's jets.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define SFM_HD __host__ __device__ __forceinline__
#else
#define SFM_HD inline
#endif

struct CamDerived {
    double R[9];      // rotation matrix (row-major) of AngleAxisRotatePoint, incl. its first-order branch
    double t[3];
    double k[3];      // unit axis (or the raw angle-axis in the small-angle branch)
    double s, c, it;  // sin(theta), cos(theta), 1/theta
    int small;        // theta^2 <= DBL_EPSILON: Ceres switches to p = X + w x X
};

SFM_HD void cam_derive(const double* cam, CamDerived& d) {
    const double w0 = cam[0], w1 = cam[1], w2 = cam[2];
    const double theta2 = w0 * w0 + w1 * w1 + w2 * w2;
    d.t[0] = cam[3]; d.t[1] = cam[4]; d.t[2] = cam[5];
    if (theta2 > DBL_EPSILON) {
        const double theta = sqrt(theta2);
        d.s = sin(theta); d.c = cos(theta); d.it = 1.0 / theta; d.small = 0;
        const double k0 = w0 * d.it, k1 = w1 * d.it, k2 = w2 * d.it, c1 = 1.0 - d.c;
        d.k[0] = k0; d.k[1] = k1; d.k[2] = k2;
        d.R[0] = d.c + c1 * k0 * k0;      d.R[1] = c1 * k0 * k1 - d.s * k2; d.R[2] = c1 * k0 * k2 + d.s * k1;
        d.R[3] = c1 * k0 * k1 + d.s * k2; d.R[4] = d.c + c1 * k1 * k1;      d.R[5] = c1 * k1 * k2 - d.s * k0;
        d.R[6] = c1 * k0 * k2 - d.s * k1; d.R[7] = c1 * k1 * k2 + d.s * k0; d.R[8] = d.c + c1 * k2 * k2;
    } else {
        d.s = 0.0; d.c = 1.0; d.it = 0.0; d.small = 1;
        d.k[0] = w0; d.k[1] = w1; d.k[2] = w2;
        d.R[0] = 1;   d.R[1] = -w2; d.R[2] = w1;
        d.R[3] = w2;  d.R[4] = 1;   d.R[5] = -w0;
        d.R[6] = -w1; d.R[7] = w0;  d.R[8] = 1;
    }
}

// residual only
SFM_HD void obs_residual(const CamDerived& d, const double* X, double f, double ox, double oy, double* r) {
    const double p0 = d.R[0] * X[0] + d.R[1] * X[1] + d.R[2] * X[2] + d.t[0];
    const double p1 = d.R[3] * X[0] + d.R[4] * X[1] + d.R[5] * X[2] + d.t[1];
    const double p2 = d.R[6] * X[0] + d.R[7] * X[1] + d.R[8] * X[2] + d.t[2];
    const double iz = 1.0 / p2;
    r[0] = f * (p0 * iz) - ox; r[1] = f * (p1 * iz) - oy;
}

// residual + Jacobian blocks (unscaled): Jc 2x6 row-major, Jp 2x3 row-major, Jf 2
SFM_HD void obs_eval(const CamDerived& d, const double* X, double f, double ox, double oy,
                     double* r, double* Jc, double* Jp, double* Jf) {
    const double x0 = X[0], x1 = X[1], x2 = X[2];
    const double q0 = d.R[0] * x0 + d.R[1] * x1 + d.R[2] * x2;     // rotated point
    const double q1 = d.R[3] * x0 + d.R[4] * x1 + d.R[5] * x2;
    const double q2 = d.R[6] * x0 + d.R[7] * x1 + d.R[8] * x2;
    const double p0 = q0 + d.t[0], p1 = q1 + d.t[1], p2 = q2 + d.t[2];
    const double iz = 1.0 / p2, xp = p0 * iz, yp = p1 * iz;
    r[0] = f * xp - ox; r[1] = f * yp - oy;
    Jf[0] = xp; Jf[1] = yp;
    // d r / d p
    const double a00 = f * iz, a02 = -f * xp * iz, a11 = f * iz, a12 = -f * yp * iz;
    // translation columns and point block
    Jc[3] = a00; Jc[4] = 0.0; Jc[5] = a02;
    Jc[9] = 0.0; Jc[10] = a11; Jc[11] = a12;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        Jp[j] = a00 * d.R[j] + a02 * d.R[6 + j];
        Jp[3 + j] = a11 * d.R[3 + j] + a12 * d.R[6 + j];
    }
    // d p / d w (3x3), column j
    double dp[3][3];
    if (!d.small) {
        const double k0 = d.k[0], k1 = d.k[1], k2 = d.k[2], s = d.s, c = d.c, c1 = 1.0 - d.c, it = d.it;
        const double kx0 = k1 * x2 - k2 * x1, kx1 = k2 * x0 - k0 * x2, kx2 = k0 * x1 - k1 * x0;   // k x X
        const double kd = k0 * x0 + k1 * x1 + k2 * x2;                                               // k . X
        // p = X c + (k x X) s + k kd (1-c);  dtheta/dw_j = k_j;  dk/dw_j = (e_j - k k_j) / theta
        // column j = k_j * A + (e_j-terms) with
        //   A = -X s + (k x X) c + k kd s   - [ (k x X) s + 2 k kd (1-c) ] / theta      (everything multiplied by k_j)
        //   e_j terms: ( e_j x X ) s/theta + e_j kd (1-c)/theta + k X_j (1-c)/theta
        const double st = s * it, ct = c1 * it;
        const double A0 = -x0 * s + kx0 * c + k0 * kd * s - kx0 * st - 2.0 * k0 * kd * ct;
        const double A1 = -x1 * s + kx1 * c + k1 * kd * s - kx1 * st - 2.0 * k1 * kd * ct;
        const double A2 = -x2 * s + kx2 * c + k2 * kd * s - kx2 * st - 2.0 * k2 * kd * ct;
        const double kdct = kd * ct;
        // e_0 x X = (0, -x2, x1); e_1 x X = (x2, 0, -x0); e_2 x X = (-x1, x0, 0)
        dp[0][0] = k0 * A0 + kdct + k0 * x0 * ct;            dp[0][1] = k1 * A0 + x2 * st + k0 * x1 * ct;         dp[0][2] = k2 * A0 - x1 * st + k0 * x2 * ct;
        dp[1][0] = k0 * A1 - x2 * st + k1 * x0 * ct;         dp[1][1] = k1 * A1 + kdct + k1 * x1 * ct;            dp[1][2] = k2 * A1 + x0 * st + k1 * x2 * ct;
        dp[2][0] = k0 * A2 + x1 * st + k2 * x0 * ct;         dp[2][1] = k1 * A2 - x0 * st + k2 * x1 * ct;         dp[2][2] = k2 * A2 + kdct + k2 * x2 * ct;
    } else {
        dp[0][0] = 0;   dp[0][1] = x2;  dp[0][2] = -x1;
        dp[1][0] = -x2; dp[1][1] = 0;   dp[1][2] = x0;
        dp[2][0] = x1;  dp[2][1] = -x0; dp[2][2] = 0;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        Jc[j] = a00 * dp[0][j] + a02 * dp[2][j];
        Jc[6 + j] = a11 * dp[1][j] + a12 * dp[2][j];
    }
}

// Cholesky U = C C^T of a symmetric 3x3 (xx xy xz yy yz zz) and M = C^-1 (lower: m00 m10 m11 m20 m21 m22).
// Returns false when U is not positive definite.
SFM_HD double sfm_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);                // one MUFU + Newton steps instead of a sqrt followed by a division
#else
    return 1.0 / sqrt(x);
#endif
}
SFM_HD bool chol3_inverse(const double* U, double* M) {
    if (!(U[0] > 0.0)) return false;
    const double i00 = sfm_rsqrt(U[0]);
    const double l10 = U[1] * i00, l20 = U[2] * i00;
    const double d1 = U[3] - l10 * l10;
    if (!(d1 > 0.0)) return false;
    const double i11 = sfm_rsqrt(d1);
    const double l21 = (U[4] - l20 * l10) * i11;
    const double d2 = U[5] - l20 * l20 - l21 * l21;
    if (!(d2 > 0.0)) return false;
    const double i22 = sfm_rsqrt(d2);
    M[0] = i00;
    M[1] = -l10 * i00 * i11; M[2] = i11;
    M[3] = -(l20 * i00 + l21 * M[1]) * i22; M[4] = -l21 * i11 * i22; M[5] = i22;
    return true;
}
