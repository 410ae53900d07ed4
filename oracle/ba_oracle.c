/*
 * oracle/ba_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement of the reference's bundle-adjustment hot path:
 *   SimpleReprojectionError::operator()       (reference SfMToyLib/SfMBundleAdjustmentUtils.cpp:58-97)
 *   SfMBundleAdjustmentUtils::adjustBundle    (reference SfMToyLib/SfMBundleAdjustmentUtils.cpp:99-222)
 *
 * The solver arithmetic lives in un-vendored Ceres Solver (version UNPINNED: find_package(Ceres REQUIRED),
 * reference CMakeLists.txt:30) which is absent from /root/reference and from this image.  This file restates
 * Ceres' published algorithm for the options the reference sets (:171-177; everything else = Ceres defaults):
 *   - cost function: AutoDiffCostFunction<SimpleReprojectionError,2,6,3,1> == forward-mode dual numbers
 *     ("jets", 10 derivative slots) through ceres::AngleAxisRotatePoint  (jacobian_mode 0, the faithful one);
 *     jacobian_mode 1 evaluates the closed-form derivative of the same expression (fast variant).
 *   - trust-region Levenberg-Marquardt minimizer with Jacobi scaling, LM diagonal clamping, the radius update
 *     rule and the four termination tests of TrustRegionMinimizer / LevenbergMarquardtStrategy;
 *   - DENSE_SCHUR: eliminate the 3D points (e-blocks), dense Cholesky of the reduced camera(+focal) system,
 *     back-substitute.
 * PARITY UNPINNED for ceres::Solve itself: the reference has no test that pins Solve()/adjustBundle output
 * (SURVEY.md section 8c).  What IS pinned: the projection model against the reference's own
 * ceres_reprojection_test fixture (SfMUnitTests.cpp:153-189, via tests/golden/reproj_fixture.npz made with
 * cv2.projectPoints), jets-vs-closed-form-vs-finite-difference Jacobians, and the optimum against
 * scipy.optimize.least_squares (independent solver) in tests/test_oracle_ba.py.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <float.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int max_num_iterations;                 /* reference :174 -> 500 */
    double max_solver_time_in_seconds;      /* reference :176 -> 10  (<= 0: no limit) */
    double function_tolerance;              /* Ceres default 1e-6  */
    double gradient_tolerance;              /* Ceres default 1e-10 */
    double parameter_tolerance;             /* Ceres default 1e-8  */
    double initial_trust_region_radius;     /* 1e4  */
    double max_trust_region_radius;         /* 1e16 */
    double min_trust_region_radius;         /* 1e-32 */
    double min_relative_decrease;           /* 1e-3 */
    double min_lm_diagonal;                 /* 1e-6 */
    double max_lm_diagonal;                 /* 1e32 */
    int jacobi_scaling;                     /* 1 */
    int max_num_consecutive_invalid_steps;  /* 5 */
    int jacobian_mode;                      /* 0 = jets (AutoDiffCostFunction), 1 = closed form */
    int num_threads;                        /* reference leaves Ceres at 1 thread (:171-177) */
    int verbose;
} sfm_oracle_ba_options;

typedef struct {
    int termination_type;                   /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
    int num_iterations;                     /* LM iterations after iteration 0 */
    int num_successful_steps, num_unsuccessful_steps;
    int num_jacobian_evals, num_residual_evals, num_linear_solves;
    double initial_cost, final_cost;
    double total_time_s, jacobian_time_s, linear_solve_time_s;
    char message[160];
} sfm_oracle_ba_summary;

void sfm_oracle_ba_default_options(sfm_oracle_ba_options* o) {
    o->max_num_iterations = 500; o->max_solver_time_in_seconds = 10.0;
    o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
    o->jacobi_scaling = 1; o->max_num_consecutive_invalid_steps = 5;
    o->jacobian_mode = 0; o->num_threads = 1; o->verbose = 0;
}

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------ */
/* Rotation conversions used at the adjustBundle boundary (:123-134 in float, :203 in double).       */
/* ------------------------------------------------------------------------------------------------ */

/* ceres::RotationMatrixToAngleAxis<float> on a COLUMN-major 3x3 (the reference passes R.t().val, :126).
 * Goes through a quaternion (Shoemake), everything in float. */
void sfm_oracle_rotmat_colmajor_to_angle_axis_f32(const float* Rc, float* aa) {
#define RM(i, j) Rc[(i) + 3 * (j)]
    float q[4];
    const float trace = RM(0, 0) + RM(1, 1) + RM(2, 2);
    if (trace >= 0.0f) {
        float t = sqrtf(trace + 1.0f);
        q[0] = 0.5f * t; t = 0.5f / t;
        q[1] = (RM(2, 1) - RM(1, 2)) * t; q[2] = (RM(0, 2) - RM(2, 0)) * t; q[3] = (RM(1, 0) - RM(0, 1)) * t;
    } else {
        int i = 0;
        if (RM(1, 1) > RM(0, 0)) i = 1;
        if (RM(2, 2) > RM(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        float t = sqrtf(RM(i, i) - RM(j, j) - RM(k, k) + 1.0f);
        q[i + 1] = 0.5f * t; t = 0.5f / t;
        q[0] = (RM(k, j) - RM(j, k)) * t; q[j + 1] = (RM(j, i) + RM(i, j)) * t; q[k + 1] = (RM(k, i) + RM(i, k)) * t;
    }
#undef RM
    const float s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (s2 > 0.0f) {
        const float s = sqrtf(s2), c = q[0];
        const float two_theta = 2.0f * ((c < 0.0f) ? atan2f(-s, -c) : atan2f(s, c));
        const float k = two_theta / s;
        aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
    } else {
        aa[0] = q[1] * 2.0f; aa[1] = q[2] * 2.0f; aa[2] = q[3] * 2.0f;
    }
}

/* ceres::AngleAxisToRotationMatrix (double), COLUMN-major output (the reference transposes on write-back, :205-209). */
void sfm_oracle_angle_axis_to_rotmat_colmajor(const double* aa, double* Rc) {
#define RM(i, j) Rc[(i) + 3 * (j)]
    const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > DBL_EPSILON) {
        const double theta = sqrt(theta2), wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
        const double c = cos(theta), s = sin(theta);
        RM(0, 0) = c + wx * wx * (1 - c);       RM(1, 0) = wz * s + wx * wy * (1 - c);  RM(2, 0) = -wy * s + wx * wz * (1 - c);
        RM(0, 1) = wx * wy * (1 - c) - wz * s;  RM(1, 1) = c + wy * wy * (1 - c);       RM(2, 1) = wx * s + wy * wz * (1 - c);
        RM(0, 2) = wy * s + wx * wz * (1 - c);  RM(1, 2) = -wx * s + wy * wz * (1 - c); RM(2, 2) = c + wz * wz * (1 - c);
    } else {
        RM(0, 0) = 1; RM(1, 0) = aa[2]; RM(2, 0) = -aa[1];
        RM(0, 1) = -aa[2]; RM(1, 1) = 1; RM(2, 1) = aa[0];
        RM(0, 2) = aa[1]; RM(1, 2) = -aa[0]; RM(2, 2) = 1;
    }
#undef RM
}

/* ------------------------------------------------------------------------------------------------ */
/* Cost functor: value only / jets / closed form                                                     */
/* ------------------------------------------------------------------------------------------------ */

/* ceres::AngleAxisRotatePoint, scalar double (also the model ceres_reprojection_test pins, in float there). */
static void rotate_point(const double* w, const double* X, double* p) {
    const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (theta2 > DBL_EPSILON) {
        const double theta = sqrt(theta2), c = cos(theta), s = sin(theta), it = 1.0 / theta;
        const double k[3] = {w[0] * it, w[1] * it, w[2] * it};
        const double kx[3] = {k[1] * X[2] - k[2] * X[1], k[2] * X[0] - k[0] * X[2], k[0] * X[1] - k[1] * X[0]};
        const double tmp = (k[0] * X[0] + k[1] * X[1] + k[2] * X[2]) * (1.0 - c);
        for (int i = 0; i < 3; ++i) p[i] = X[i] * c + kx[i] * s + k[i] * tmp;
    } else {
        p[0] = X[0] + (w[1] * X[2] - w[2] * X[1]);
        p[1] = X[1] + (w[2] * X[0] - w[0] * X[2]);
        p[2] = X[2] + (w[0] * X[1] - w[1] * X[0]);
    }
}

/* residual only: SimpleReprojectionError::operator()<double>  (:62-88) */
void sfm_oracle_ba_residual(const double* cam, const double* pt, double focal, double ox, double oy, double* r) {
    double p[3];
    rotate_point(cam, pt, p);
    p[0] += cam[3]; p[1] += cam[4]; p[2] += cam[5];
    const double xp = p[0] / p[2], yp = p[1] / p[2];
    r[0] = focal * xp - ox; r[1] = focal * yp - oy;
}

/* --- forward-mode dual numbers with 10 derivative slots: cam 0..5, point 6..8, focal 9 --- */
#define NJ 10
typedef struct { double a; double v[NJ]; } jet;
static inline jet j_const(double a) { jet r; r.a = a; for (int i = 0; i < NJ; ++i) r.v[i] = 0; return r; }
static inline jet j_var(double a, int k) { jet r = j_const(a); r.v[k] = 1.0; return r; }
static inline jet j_add(jet x, jet y) { jet r; r.a = x.a + y.a; for (int i = 0; i < NJ; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
static inline jet j_sub(jet x, jet y) { jet r; r.a = x.a - y.a; for (int i = 0; i < NJ; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
static inline jet j_mul(jet x, jet y) { jet r; r.a = x.a * y.a; for (int i = 0; i < NJ; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
static inline jet j_div(jet x, jet y) {
    jet r; const double iy = 1.0 / y.a; r.a = x.a * iy;
    for (int i = 0; i < NJ; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * iy;
    return r;
}
static inline jet j_sqrt(jet x) { jet r; r.a = sqrt(x.a); const double k = 1.0 / (2.0 * r.a); for (int i = 0; i < NJ; ++i) r.v[i] = x.v[i] * k; return r; }
static inline jet j_sin(jet x) { jet r; r.a = sin(x.a); const double c = cos(x.a); for (int i = 0; i < NJ; ++i) r.v[i] = c * x.v[i]; return r; }
static inline jet j_cos(jet x) { jet r; r.a = cos(x.a); const double s = -sin(x.a); for (int i = 0; i < NJ; ++i) r.v[i] = s * x.v[i]; return r; }
static inline jet j_scal_sub(double s, jet x) { jet r; r.a = s - x.a; for (int i = 0; i < NJ; ++i) r.v[i] = -x.v[i]; return r; }

static void residual_jets(const double* cam, const double* pt, double focal, double ox, double oy,
                          double* r, double* Jc /*2x6*/, double* Jp /*2x3*/, double* Jf /*2*/) {
    jet w[3], t[3], X[3], f, p[3];
    for (int i = 0; i < 3; ++i) { w[i] = j_var(cam[i], i); t[i] = j_var(cam[3 + i], 3 + i); X[i] = j_var(pt[i], 6 + i); }
    f = j_var(focal, 9);
    const jet theta2 = j_add(j_add(j_mul(w[0], w[0]), j_mul(w[1], w[1])), j_mul(w[2], w[2]));
    if (theta2.a > DBL_EPSILON) {
        const jet theta = j_sqrt(theta2), c = j_cos(theta), s = j_sin(theta), it = j_div(j_const(1.0), theta);
        const jet k[3] = {j_mul(w[0], it), j_mul(w[1], it), j_mul(w[2], it)};
        const jet kx[3] = {j_sub(j_mul(k[1], X[2]), j_mul(k[2], X[1])), j_sub(j_mul(k[2], X[0]), j_mul(k[0], X[2])),
                           j_sub(j_mul(k[0], X[1]), j_mul(k[1], X[0]))};
        const jet tmp = j_mul(j_add(j_add(j_mul(k[0], X[0]), j_mul(k[1], X[1])), j_mul(k[2], X[2])), j_scal_sub(1.0, c));
        for (int i = 0; i < 3; ++i) p[i] = j_add(j_add(j_mul(X[i], c), j_mul(kx[i], s)), j_mul(k[i], tmp));
    } else {
        p[0] = j_add(X[0], j_sub(j_mul(w[1], X[2]), j_mul(w[2], X[1])));
        p[1] = j_add(X[1], j_sub(j_mul(w[2], X[0]), j_mul(w[0], X[2])));
        p[2] = j_add(X[2], j_sub(j_mul(w[0], X[1]), j_mul(w[1], X[0])));
    }
    for (int i = 0; i < 3; ++i) p[i] = j_add(p[i], t[i]);
    const jet xp = j_div(p[0], p[2]), yp = j_div(p[1], p[2]);
    const jet r0 = j_sub(j_mul(f, xp), j_const(ox)), r1 = j_sub(j_mul(f, yp), j_const(oy));
    r[0] = r0.a; r[1] = r1.a;
    for (int i = 0; i < 6; ++i) { Jc[i] = r0.v[i]; Jc[6 + i] = r1.v[i]; }
    for (int i = 0; i < 3; ++i) { Jp[i] = r0.v[6 + i]; Jp[3 + i] = r1.v[6 + i]; }
    Jf[0] = r0.v[9]; Jf[1] = r1.v[9];
}

/* closed-form derivative of exactly the same expression */
static void residual_closed_form(const double* cam, const double* pt, double focal, double ox, double oy,
                                 double* r, double* Jc, double* Jp, double* Jf) {
    const double* w = cam; const double* X = pt;
    double p[3], dpdw[3][3], R[3][3];
    const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (theta2 > DBL_EPSILON) {
        const double theta = sqrt(theta2), c = cos(theta), s = sin(theta), it = 1.0 / theta, c1 = 1.0 - c;
        const double k[3] = {w[0] * it, w[1] * it, w[2] * it};
        const double kx[3] = {k[1] * X[2] - k[2] * X[1], k[2] * X[0] - k[0] * X[2], k[0] * X[1] - k[1] * X[0]};
        const double kd = k[0] * X[0] + k[1] * X[1] + k[2] * X[2];
        for (int i = 0; i < 3; ++i) p[i] = X[i] * c + kx[i] * s + k[i] * kd * c1;
        /* R = c I + s [k]x + (1-c) k k^T */
        R[0][0] = c + c1 * k[0] * k[0];        R[0][1] = c1 * k[0] * k[1] - s * k[2]; R[0][2] = c1 * k[0] * k[2] + s * k[1];
        R[1][0] = c1 * k[0] * k[1] + s * k[2]; R[1][1] = c + c1 * k[1] * k[1];        R[1][2] = c1 * k[1] * k[2] - s * k[0];
        R[2][0] = c1 * k[0] * k[2] - s * k[1]; R[2][1] = c1 * k[1] * k[2] + s * k[0]; R[2][2] = c + c1 * k[2] * k[2];
        for (int j = 0; j < 3; ++j) {
            /* dk/dw_j = (e_j - k k_j)/theta ; dtheta/dw_j = k_j */
            double dk[3] = {-k[0] * k[j] * it, -k[1] * k[j] * it, -k[2] * k[j] * it};
            dk[j] += it;
            const double dkx[3] = {dk[1] * X[2] - dk[2] * X[1], dk[2] * X[0] - dk[0] * X[2], dk[0] * X[1] - dk[1] * X[0]};
            const double dkd = dk[0] * X[0] + dk[1] * X[1] + dk[2] * X[2];
            for (int i = 0; i < 3; ++i)
                dpdw[i][j] = -X[i] * s * k[j] + kx[i] * c * k[j] + dkx[i] * s
                             + dk[i] * kd * c1 + k[i] * dkd * c1 + k[i] * kd * s * k[j];
        }
    } else {
        p[0] = X[0] + (w[1] * X[2] - w[2] * X[1]);
        p[1] = X[1] + (w[2] * X[0] - w[0] * X[2]);
        p[2] = X[2] + (w[0] * X[1] - w[1] * X[0]);
        R[0][0] = 1; R[0][1] = -w[2]; R[0][2] = w[1];
        R[1][0] = w[2]; R[1][1] = 1; R[1][2] = -w[0];
        R[2][0] = -w[1]; R[2][1] = w[0]; R[2][2] = 1;
        /* d(w x X)/dw_j = e_j x X */
        dpdw[0][0] = 0;     dpdw[0][1] = X[2];  dpdw[0][2] = -X[1];
        dpdw[1][0] = -X[2]; dpdw[1][1] = 0;     dpdw[1][2] = X[0];
        dpdw[2][0] = X[1];  dpdw[2][1] = -X[0]; dpdw[2][2] = 0;
    }
    p[0] += cam[3]; p[1] += cam[4]; p[2] += cam[5];
    const double iz = 1.0 / p[2], xp = p[0] * iz, yp = p[1] * iz;
    r[0] = focal * xp - ox; r[1] = focal * yp - oy;
    /* d r / d P */
    const double a0[3] = {focal * iz, 0.0, -focal * xp * iz}, a1[3] = {0.0, focal * iz, -focal * yp * iz};
    for (int j = 0; j < 3; ++j) {
        Jc[j] = a0[0] * dpdw[0][j] + a0[1] * dpdw[1][j] + a0[2] * dpdw[2][j];
        Jc[6 + j] = a1[0] * dpdw[0][j] + a1[1] * dpdw[1][j] + a1[2] * dpdw[2][j];
        Jc[3 + j] = a0[j]; Jc[9 + j] = a1[j];
        Jp[j] = a0[0] * R[0][j] + a0[1] * R[1][j] + a0[2] * R[2][j];
        Jp[3 + j] = a1[0] * R[0][j] + a1[1] * R[1][j] + a1[2] * R[2][j];
    }
    Jf[0] = xp; Jf[1] = yp;
}

void sfm_oracle_ba_residual_jacobian(const double* cam, const double* pt, double focal, double ox, double oy, int mode,
                                     double* r, double* Jc, double* Jp, double* Jf) {
    if (mode == 0) residual_jets(cam, pt, focal, ox, oy, r, Jc, Jp, Jf);
    else residual_closed_form(cam, pt, focal, ox, oy, r, Jc, Jp, Jf);
}

/* ------------------------------------------------------------------------------------------------ */
/* Problem = flat arrays in the order the reference adds residual blocks (:142-166):                 */
/* for each point i (cloud order), for each (view, feature) in its std::map (ascending view).        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int nc, np, nobs;
    const float* obs_xy;     /* [nobs*2], principal point already subtracted IN FLOAT (:149-153) */
    const int32_t* obs_cam;  /* [nobs] */
    const int32_t* pt_off;   /* [np+1] CSR */
    int32_t* obs_pt;         /* [nobs] derived */
} ba_problem;

/* cost = 1/2 sum r^2 ; returns DBL_MAX when any residual is not finite (Ceres: evaluation failure) */
static double eval_cost(const ba_problem* P, const double* x, int nthreads) {
    const double* cams = x; const double* pts = x + 6 * P->nc; const double focal = x[6 * P->nc + 3 * P->np];
    double sum = 0; int bad = 0;
    (void)nthreads;
#pragma omp parallel for schedule(static) reduction(+ : sum) reduction(| : bad) num_threads(nthreads)
    for (int o = 0; o < P->nobs; ++o) {
        double r[2];
        sfm_oracle_ba_residual(cams + 6 * P->obs_cam[o], pts + 3 * P->obs_pt[o], focal,
                               (double)P->obs_xy[2 * o], (double)P->obs_xy[2 * o + 1], r);
        if (!isfinite(r[0]) || !isfinite(r[1])) bad |= 1;
        sum += r[0] * r[0] + r[1] * r[1];
    }
    if (bad || !isfinite(sum)) return DBL_MAX;
    return 0.5 * sum;
}

/* residuals + unscaled Jacobian blocks; returns 0 on evaluation failure */
static int eval_jacobian(const ba_problem* P, const double* x, int mode, int nthreads,
                         double* res /*2*nobs*/, double* Jc /*12*nobs*/, double* Jp /*6*nobs*/, double* Jf /*2*nobs*/,
                         double* cost) {
    const double* cams = x; const double* pts = x + 6 * P->nc; const double focal = x[6 * P->nc + 3 * P->np];
    double sum = 0; int bad = 0;
    (void)nthreads;
#pragma omp parallel for schedule(static) reduction(+ : sum) reduction(| : bad) num_threads(nthreads)
    for (int o = 0; o < P->nobs; ++o) {
        sfm_oracle_ba_residual_jacobian(cams + 6 * P->obs_cam[o], pts + 3 * P->obs_pt[o], focal,
                                        (double)P->obs_xy[2 * o], (double)P->obs_xy[2 * o + 1], mode,
                                        res + 2 * o, Jc + 12 * o, Jp + 6 * o, Jf + 2 * o);
        if (!isfinite(res[2 * o]) || !isfinite(res[2 * o + 1])) bad |= 1;
        sum += res[2 * o] * res[2 * o] + res[2 * o + 1] * res[2 * o + 1];
    }
    *cost = 0.5 * sum;
    return !(bad || !isfinite(sum));
}

/* gradient g = J^T r (unscaled J); layout like x.  Parallel over points (observations of a point are contiguous);
 * camera/focal entries go through per-thread accumulators. */
static void eval_gradient(const ba_problem* P, const double* res, const double* Jc, const double* Jp, const double* Jf,
                          double* g, int nthreads) {
    const int n = 6 * P->nc + 3 * P->np + 1, ncf = 6 * P->nc + 1;
    memset(g, 0, sizeof(double) * n);
    double* gp = g + 6 * P->nc;
    (void)nthreads;
#pragma omp parallel num_threads(nthreads)
    {
        double* loc = (double*)calloc(ncf, sizeof(double));
#pragma omp for schedule(static)
        for (int p = 0; p < P->np; ++p) {
            for (int o = P->pt_off[p]; o < P->pt_off[p + 1]; ++o) {
                const double r0 = res[2 * o], r1 = res[2 * o + 1];
                double* c = loc + 6 * P->obs_cam[o];
                for (int k = 0; k < 6; ++k) c[k] += Jc[12 * o + k] * r0 + Jc[12 * o + 6 + k] * r1;
                for (int k = 0; k < 3; ++k) gp[3 * p + k] += Jp[6 * o + k] * r0 + Jp[6 * o + 3 + k] * r1;
                loc[ncf - 1] += Jf[2 * o] * r0 + Jf[2 * o + 1] * r1;
            }
        }
#pragma omp critical
        { for (int i = 0; i < 6 * P->nc; ++i) g[i] += loc[i]; g[n - 1] += loc[ncf - 1]; }
        free(loc);
    }
}

/* squared column norms of the (already scaled) Jacobian; layout like x */
static void squared_column_norms(const ba_problem* P, const double* Jc, const double* Jp, const double* Jf, double* d, int nthreads) {
    const int n = 6 * P->nc + 3 * P->np + 1, ncf = 6 * P->nc + 1;
    memset(d, 0, sizeof(double) * n);
    double* dp = d + 6 * P->nc;
    (void)nthreads;
#pragma omp parallel num_threads(nthreads)
    {
        double* loc = (double*)calloc(ncf, sizeof(double));
#pragma omp for schedule(static)
        for (int p = 0; p < P->np; ++p) {
            for (int o = P->pt_off[p]; o < P->pt_off[p + 1]; ++o) {
                double* c = loc + 6 * P->obs_cam[o];
                for (int k = 0; k < 6; ++k) c[k] += Jc[12 * o + k] * Jc[12 * o + k] + Jc[12 * o + 6 + k] * Jc[12 * o + 6 + k];
                for (int k = 0; k < 3; ++k) dp[3 * p + k] += Jp[6 * o + k] * Jp[6 * o + k] + Jp[6 * o + 3 + k] * Jp[6 * o + 3 + k];
                loc[ncf - 1] += Jf[2 * o] * Jf[2 * o] + Jf[2 * o + 1] * Jf[2 * o + 1];
            }
        }
#pragma omp critical
        { for (int i = 0; i < 6 * P->nc; ++i) d[i] += loc[i]; d[n - 1] += loc[ncf - 1]; }
        free(loc);
    }
}

static void scale_columns(const ba_problem* P, const double* scale, double* Jc, double* Jp, double* Jf, int nthreads) {
    const double* sc = scale; const double* sp = scale + 6 * P->nc; const double sf = scale[6 * P->nc + 3 * P->np];
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int o = 0; o < P->nobs; ++o) {
        const double* c = sc + 6 * P->obs_cam[o]; const double* p = sp + 3 * P->obs_pt[o];
        for (int k = 0; k < 6; ++k) { Jc[12 * o + k] *= c[k]; Jc[12 * o + 6 + k] *= c[k]; }
        for (int k = 0; k < 3; ++k) { Jp[6 * o + k] *= p[k]; Jp[6 * o + 3 + k] *= p[k]; }
        Jf[2 * o] *= sf; Jf[2 * o + 1] *= sf;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* DENSE_SCHUR                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* inverse of a symmetric positive definite 3x3 through its Cholesky factor; returns 0 if not SPD */
static int inv_spd3(const double U[6] /* xx xy xz yy yz zz */, double Ui[6]) {
    const double l00 = sqrt(U[0]); if (!(U[0] > 0)) return 0;
    const double l10 = U[1] / l00, l20 = U[2] / l00;
    const double d1 = U[3] - l10 * l10; if (!(d1 > 0)) return 0;
    const double l11 = sqrt(d1), l21 = (U[4] - l20 * l10) / l11;
    const double d2 = U[5] - l20 * l20 - l21 * l21; if (!(d2 > 0)) return 0;
    const double l22 = sqrt(d2);
    /* M = L^-1 (lower) */
    const double m00 = 1 / l00, m11 = 1 / l11, m22 = 1 / l22;
    const double m10 = -l10 * m00 * m11, m21 = -l21 * m11 * m22, m20 = -(l20 * m00 + l21 * m10) * m22;
    /* U^-1 = M^T M */
    Ui[0] = m00 * m00 + m10 * m10 + m20 * m20; Ui[1] = m10 * m11 + m20 * m21; Ui[2] = m20 * m22;
    Ui[3] = m11 * m11 + m21 * m21;             Ui[4] = m21 * m22;             Ui[5] = m22 * m22;
    return 1;
}

/* in-place dense Cholesky A = L L^T on the lower triangle of row-major n x n; returns 0 if not SPD */
static int dense_cholesky(double* A, int n, int nthreads) {
    (void)nthreads;
    for (int j = 0; j < n; ++j) {
        double* Aj = A + (size_t)j * n;
        double d = Aj[j];
        for (int k = 0; k < j; ++k) d -= Aj[k] * Aj[k];
        if (!(d > 0) || !isfinite(d)) return 0;
        const double ljj = sqrt(d); Aj[j] = ljj;
        const double inv = 1.0 / ljj;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (n - j > 256)
        for (int i = j + 1; i < n; ++i) {
            double* Ai = A + (size_t)i * n;
            double s = Ai[j];
            for (int k = 0; k < j; ++k) s -= Ai[k] * Aj[k];
            Ai[j] = s * inv;
        }
    }
    return 1;
}
static void cholesky_solve(const double* L, int n, double* b) {
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k]; b[i] = s / L[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k]; b[i] = s / L[(size_t)i * n + i]; }
}

/*
 * Build the reduced camera(+focal) system for scaled Jacobian blocks, residuals and LM diagonal D (layout like x).
 *   S   [(6nc+1)^2] row-major, full symmetric;  rhs [6nc+1]
 *   Uinv [6*np], gpt [3*np]  (kept for back substitution)
 * Returns 0 when a point block is not SPD.
 */
static int build_reduced_system(const ba_problem* P, const double* res, const double* Jc, const double* Jp, const double* Jf,
                                const double* D, double* S, double* rhs, double* Uinv, double* gpt, int nthreads,
                                double* pool /* nthreads * (n*n + n) doubles of per-thread accumulators, or NULL */) {
    const int nc = P->nc, np = P->np, n = 6 * nc + 1, fidx = 6 * nc;
    const double* Dp = D + 6 * nc;
    int ok = 1;
    memset(S, 0, sizeof(double) * (size_t)n * n);
    memset(rhs, 0, sizeof(double) * n);
    (void)nthreads;
    const size_t stride = (size_t)n * n + n;
    int used_threads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double* Sl = S; double* rl = rhs; int own = 0;
#ifdef _OPENMP
        if (omp_get_num_threads() > 1 && pool) {
            Sl = pool + stride * omp_get_thread_num(); rl = Sl + (size_t)n * n; own = 1;
            memset(Sl, 0, sizeof(double) * stride);
#pragma omp single
            used_threads = omp_get_num_threads();
        } else if (omp_get_num_threads() > 1) {
            Sl = (double*)calloc((size_t)n * n, sizeof(double)); rl = (double*)calloc(n, sizeof(double)); own = 2;
        }
#endif
        double (*W)[21] = NULL; int wcap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int p = 0; p < np; ++p) {
            const int o0 = P->pt_off[p], o1 = P->pt_off[p + 1], k = o1 - o0;
            if (k <= 0) { for (int a = 0; a < 6; ++a) Uinv[6 * p + a] = 0; for (int a = 0; a < 3; ++a) gpt[3 * p + a] = 0; continue; }
            if (k > wcap) { free(W); wcap = k; W = (double (*)[21])malloc(sizeof(double[21]) * wcap); }
            double U[6] = {Dp[3 * p] * Dp[3 * p], 0, 0, Dp[3 * p + 1] * Dp[3 * p + 1], 0, Dp[3 * p + 2] * Dp[3 * p + 2]};
            double g[3] = {0, 0, 0};
            for (int o = o0; o < o1; ++o) {
                const double* e = Jp + 6 * o; const double r0 = res[2 * o], r1 = res[2 * o + 1];
                U[0] += e[0] * e[0] + e[3] * e[3]; U[1] += e[0] * e[1] + e[3] * e[4]; U[2] += e[0] * e[2] + e[3] * e[5];
                U[3] += e[1] * e[1] + e[4] * e[4]; U[4] += e[1] * e[2] + e[4] * e[5]; U[5] += e[2] * e[2] + e[5] * e[5];
                for (int a = 0; a < 3; ++a) g[a] += e[a] * r0 + e[3 + a] * r1;
                /* F^T F and F^T r for this row block (camera 6 + focal 1) */
                const int c = P->obs_cam[o];
                double F[2][7];
                for (int a = 0; a < 6; ++a) { F[0][a] = Jc[12 * o + a]; F[1][a] = Jc[12 * o + 6 + a]; }
                F[0][6] = Jf[2 * o]; F[1][6] = Jf[2 * o + 1];
                int idx[7]; for (int a = 0; a < 6; ++a) idx[a] = 6 * c + a; idx[6] = fidx;
                for (int a = 0; a < 7; ++a) {
                    rl[idx[a]] += F[0][a] * r0 + F[1][a] * r1;
                    for (int b = 0; b < 7; ++b) Sl[(size_t)idx[a] * n + idx[b]] += F[0][a] * F[0][b] + F[1][a] * F[1][b];
                    for (int b = 0; b < 3; ++b) W[o - o0][3 * a + b] = F[0][a] * e[b] + F[1][a] * e[3 + b];
                }
            }
            double Ui[6];
            if (!inv_spd3(U, Ui)) { ok = 0; continue; }
            for (int a = 0; a < 6; ++a) Uinv[6 * p + a] = Ui[a];
            for (int a = 0; a < 3; ++a) gpt[3 * p + a] = g[a];
            const double Uf[3][3] = {{Ui[0], Ui[1], Ui[2]}, {Ui[1], Ui[3], Ui[4]}, {Ui[2], Ui[4], Ui[5]}};
            const double ug[3] = {Uf[0][0] * g[0] + Uf[0][1] * g[1] + Uf[0][2] * g[2],
                                  Uf[1][0] * g[0] + Uf[1][1] * g[1] + Uf[1][2] * g[2],
                                  Uf[2][0] * g[0] + Uf[2][1] * g[1] + Uf[2][2] * g[2]};
            for (int i = 0; i < k; ++i) {
                const int ci = P->obs_cam[o0 + i];
                double T[21]; /* W_i U^-1 */
                for (int a = 0; a < 7; ++a) for (int b = 0; b < 3; ++b)
                    T[3 * a + b] = W[i][3 * a] * Uf[0][b] + W[i][3 * a + 1] * Uf[1][b] + W[i][3 * a + 2] * Uf[2][b];
                for (int a = 0; a < 7; ++a) {
                    const int ia = a < 6 ? 6 * ci + a : fidx;
                    rl[ia] -= W[i][3 * a] * ug[0] + W[i][3 * a + 1] * ug[1] + W[i][3 * a + 2] * ug[2];
                    for (int j = 0; j < k; ++j) {
                        const int cj = P->obs_cam[o0 + j];
                        for (int b = 0; b < 7; ++b) {
                            const int jb = b < 6 ? 6 * cj + b : fidx;
                            Sl[(size_t)ia * n + jb] -= T[3 * a] * W[j][3 * b] + T[3 * a + 1] * W[j][3 * b + 1] + T[3 * a + 2] * W[j][3 * b + 2];
                        }
                    }
                }
            }
        }
        free(W);
        if (own == 2) {
#pragma omp critical
            { for (size_t i = 0; i < (size_t)n * n; ++i) S[i] += Sl[i]; for (int i = 0; i < n; ++i) rhs[i] += rl[i]; }
            free(Sl); free(rl);
        }
    }
    if (pool && used_threads > 1) {      /* parallel merge of the per-thread accumulators */
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (long i = 0; i < (long)stride; ++i) {
            double a = 0;
            for (int t = 0; t < used_threads; ++t) a += pool[stride * t + i];
            if (i < (long)n * n) S[i] += a; else rhs[i - (long)n * n] += a;
        }
    }
    for (int i = 0; i < n; ++i) { const double d = i < 6 * nc ? D[i] : D[6 * nc + 3 * np]; S[(size_t)i * n + i] += d * d; }
    return ok;
}

/* exported for kernel-level parity tests: reduced system at x for a given radius (D from clamped column norms) */
int sfm_oracle_ba_reduced_system(int nc, int np, int nobs, const double* cams, const double* pts, double focal,
                                 const float* obs_xy, const int32_t* obs_cam, const int32_t* pt_off,
                                 int jacobi_scaling, const double* scale_in /* NULL: compute at x */, double radius,
                                 double min_diag, double max_diag, int mode,
                                 double* S, double* rhs, double* scale_out /* may be NULL */, double* grad_out /* may be NULL */,
                                 double* cost_out) {
    ba_problem P = {nc, np, nobs, obs_xy, obs_cam, pt_off, NULL};
    const int n = 6 * nc + 3 * np + 1;
    P.obs_pt = (int32_t*)malloc(sizeof(int32_t) * (nobs > 0 ? nobs : 1));
    for (int p = 0; p < np; ++p) for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) P.obs_pt[o] = p;
    double* x = (double*)malloc(sizeof(double) * n);
    memcpy(x, cams, sizeof(double) * 6 * nc); memcpy(x + 6 * nc, pts, sizeof(double) * 3 * np); x[n - 1] = focal;
    double* res = (double*)malloc(sizeof(double) * 2 * nobs); double* Jc = (double*)malloc(sizeof(double) * 12 * nobs);
    double* Jp = (double*)malloc(sizeof(double) * 6 * nobs); double* Jf = (double*)malloc(sizeof(double) * 2 * nobs);
    double* scale = (double*)malloc(sizeof(double) * n); double* D = (double*)malloc(sizeof(double) * n);
    double* Uinv = (double*)malloc(sizeof(double) * 6 * np); double* gpt = (double*)malloc(sizeof(double) * 3 * np);
    double cost;
    int ok = eval_jacobian(&P, x, mode, 1, res, Jc, Jp, Jf, &cost);
    if (grad_out) eval_gradient(&P, res, Jc, Jp, Jf, grad_out, 1);
    if (scale_in) memcpy(scale, scale_in, sizeof(double) * n);
    else if (jacobi_scaling) { squared_column_norms(&P, Jc, Jp, Jf, scale, 1); for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + sqrt(scale[i])); }
    else for (int i = 0; i < n; ++i) scale[i] = 1.0;
    scale_columns(&P, scale, Jc, Jp, Jf, 1);
    squared_column_norms(&P, Jc, Jp, Jf, D, 1);
    for (int i = 0; i < n; ++i) { double d = D[i]; d = d < min_diag ? min_diag : d; d = d > max_diag ? max_diag : d; D[i] = sqrt(d / radius); }
    ok = build_reduced_system(&P, res, Jc, Jp, Jf, D, S, rhs, Uinv, gpt, 1, NULL) && ok;
    if (scale_out) memcpy(scale_out, scale, sizeof(double) * n);
    if (cost_out) *cost_out = cost;
    free(P.obs_pt); free(x); free(res); free(Jc); free(Jp); free(Jf); free(scale); free(D); free(Uinv); free(gpt);
    return ok;
}

double sfm_oracle_ba_cost(int nc, int np, int nobs, const double* cams, const double* pts, double focal,
                          const float* obs_xy, const int32_t* obs_cam, const int32_t* pt_off, int nthreads) {
    ba_problem P = {nc, np, nobs, obs_xy, obs_cam, pt_off, NULL};
    const int n = 6 * nc + 3 * np + 1;
    P.obs_pt = (int32_t*)malloc(sizeof(int32_t) * (nobs > 0 ? nobs : 1));
    for (int p = 0; p < np; ++p) for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) P.obs_pt[o] = p;
    double* x = (double*)malloc(sizeof(double) * n);
    memcpy(x, cams, sizeof(double) * 6 * nc); memcpy(x + 6 * nc, pts, sizeof(double) * 3 * np); x[n - 1] = focal;
    const double c = eval_cost(&P, x, nthreads > 0 ? nthreads : 1);
    free(P.obs_pt); free(x);
    return c;
}

/* ------------------------------------------------------------------------------------------------ */
/* Trust-region Levenberg-Marquardt (Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy)        */
/* ------------------------------------------------------------------------------------------------ */
/*
 * cams [6*nc] (angle-axis, t), pts [3*np], focal[1] are updated in place with the final iterate x
 * (Ceres always leaves the last accepted point in the parameter blocks; the CONVERGENCE-only write-back
 * rule of the reference (:182-185) concerns copying them back into cv::Matx34f/Point3f and lives in the caller).
 * trace (optional) [ (max_num_iterations+1) * 4 ] : cost, trust-region radius after the iteration,
 * step accepted flag (-1 for iteration 0), gradient max norm.
 */
int sfm_oracle_ba_solve(const sfm_oracle_ba_options* opt, int nc, int np, int nobs,
                        double* cams, double* pts, double* focal,
                        const float* obs_xy, const int32_t* obs_cam, const int32_t* pt_off,
                        sfm_oracle_ba_summary* sum, double* trace) {
    const double t_start = now_s();
    const int nthreads = opt->num_threads > 0 ? opt->num_threads : 1;
    ba_problem P = {nc, np, nobs, obs_xy, obs_cam, pt_off, NULL};
    const int n = 6 * nc + 3 * np + 1, nr = 6 * nc + 1;
    memset(sum, 0, sizeof(*sum));
    P.obs_pt = (int32_t*)malloc(sizeof(int32_t) * (nobs > 0 ? nobs : 1));
    for (int p = 0; p < np; ++p) for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) P.obs_pt[o] = p;

    double* x = (double*)malloc(sizeof(double) * n); double* xc = (double*)malloc(sizeof(double) * n);
    double* res = (double*)malloc(sizeof(double) * 2 * (size_t)nobs); double* Jc = (double*)malloc(sizeof(double) * 12 * (size_t)nobs);
    double* Jp = (double*)malloc(sizeof(double) * 6 * (size_t)nobs); double* Jf = (double*)malloc(sizeof(double) * 2 * (size_t)nobs);
    double* scale = (double*)malloc(sizeof(double) * n); double* diag = (double*)malloc(sizeof(double) * n);
    double* D = (double*)malloc(sizeof(double) * n); double* g = (double*)malloc(sizeof(double) * n);
    double* step = (double*)malloc(sizeof(double) * n);
    double* S = (double*)malloc(sizeof(double) * (size_t)nr * nr); double* rhs = (double*)malloc(sizeof(double) * nr);
    double* Uinv = (double*)malloc(sizeof(double) * 6 * (size_t)np); double* gpt = (double*)malloc(sizeof(double) * 3 * (size_t)np);
    double* pool = nthreads > 1 ? (double*)malloc(sizeof(double) * (size_t)nthreads * ((size_t)nr * nr + nr)) : NULL;
    memcpy(x, cams, sizeof(double) * 6 * nc); memcpy(x + 6 * nc, pts, sizeof(double) * 3 * np); x[n - 1] = *focal;

    double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
    int reuse_diagonal = 0, invalid_steps = 0, iter = 0;
    double x_cost = 0, x_norm = 0, gmax = 0;
    sum->termination_type = 1;

#define EVAL_JAC()                                                                              \
    do {                                                                                        \
        const double t0_ = now_s();                                                             \
        jac_ok = eval_jacobian(&P, x, opt->jacobian_mode, nthreads, res, Jc, Jp, Jf, &x_cost);  \
        sum->num_jacobian_evals++; sum->num_residual_evals++;                                   \
        eval_gradient(&P, res, Jc, Jp, Jf, g, nthreads);                                        \
        gmax = 0; for (int i_ = 0; i_ < n; ++i_) { const double a_ = fabs(g[i_]); if (a_ > gmax) gmax = a_; } \
        sum->jacobian_time_s += now_s() - t0_;                                                  \
    } while (0)

    int jac_ok;
    EVAL_JAC();
    if (!jac_ok) { sum->termination_type = 2; snprintf(sum->message, sizeof sum->message, "Residual and Jacobian evaluation failed."); goto done; }
    if (opt->jacobi_scaling) {
        squared_column_norms(&P, Jc, Jp, Jf, scale, nthreads);
        for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + sqrt(scale[i]));
    } else for (int i = 0; i < n; ++i) scale[i] = 1.0;
    scale_columns(&P, scale, Jc, Jp, Jf, nthreads);
    x_norm = 0; for (int i = 0; i < n; ++i) x_norm += x[i] * x[i]; x_norm = sqrt(x_norm);
    sum->initial_cost = x_cost;
    if (trace) { trace[0] = x_cost; trace[1] = radius; trace[2] = -1; trace[3] = gmax; }
    if (opt->verbose) printf("iter %3d cost %.9e |g|max %.3e radius %.3e\n", 0, x_cost, gmax, radius);
    if (gmax <= opt->gradient_tolerance) {
        sum->termination_type = 0; snprintf(sum->message, sizeof sum->message, "Gradient tolerance reached."); goto done;
    }

    for (;;) {
        /* FinalizeIterationAndCheckIfMinimizerCanContinue (time / iterations / gradient / radius) */
        if (opt->max_solver_time_in_seconds > 0 && now_s() - t_start >= opt->max_solver_time_in_seconds) {
            sum->termination_type = 1; snprintf(sum->message, sizeof sum->message, "Maximum solver time reached."); break;
        }
        if (iter >= opt->max_num_iterations) {
            sum->termination_type = 1; snprintf(sum->message, sizeof sum->message, "Maximum number of iterations reached."); break;
        }
        if (iter > 0 && gmax <= opt->gradient_tolerance) {
            sum->termination_type = 0; snprintf(sum->message, sizeof sum->message, "Gradient tolerance reached."); break;
        }
        if (radius <= opt->min_trust_region_radius) {
            sum->termination_type = 0; snprintf(sum->message, sizeof sum->message, "Minimum trust region radius reached."); break;
        }
        ++iter;
        int accepted = 0;

        /* LevenbergMarquardtStrategy::ComputeStep */
        const double t0 = now_s();
        if (!reuse_diagonal) {
            squared_column_norms(&P, Jc, Jp, Jf, diag, nthreads);
            for (int i = 0; i < n; ++i) { double d = diag[i]; d = d < opt->min_lm_diagonal ? opt->min_lm_diagonal : d; d = d > opt->max_lm_diagonal ? opt->max_lm_diagonal : d; diag[i] = d; }
        }
        for (int i = 0; i < n; ++i) D[i] = sqrt(diag[i] / radius);
        int lin_ok = build_reduced_system(&P, res, Jc, Jp, Jf, D, S, rhs, Uinv, gpt, nthreads, pool);
        if (lin_ok) lin_ok = dense_cholesky(S, nr, nthreads);
        if (lin_ok) {
            cholesky_solve(S, nr, rhs);               /* rhs := y_f */
            const double* yc = rhs; const double yf = rhs[6 * nc];
            for (int i = 0; i < 6 * nc; ++i) step[i] = -yc[i];
            step[n - 1] = -yf;
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int p = 0; p < np; ++p) {
                double t[3] = {gpt[3 * p], gpt[3 * p + 1], gpt[3 * p + 2]};
                for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) {
                    const double* c = yc + 6 * obs_cam[o];
                    double m0 = Jf[2 * o] * yf, m1 = Jf[2 * o + 1] * yf;
                    for (int k = 0; k < 6; ++k) { m0 += Jc[12 * o + k] * c[k]; m1 += Jc[12 * o + 6 + k] * c[k]; }
                    for (int k = 0; k < 3; ++k) t[k] -= Jp[6 * o + k] * m0 + Jp[6 * o + 3 + k] * m1;
                }
                const double* Ui = Uinv + 6 * p;
                step[6 * nc + 3 * p + 0] = -(Ui[0] * t[0] + Ui[1] * t[1] + Ui[2] * t[2]);
                step[6 * nc + 3 * p + 1] = -(Ui[1] * t[0] + Ui[3] * t[1] + Ui[4] * t[2]);
                step[6 * nc + 3 * p + 2] = -(Ui[2] * t[0] + Ui[4] * t[1] + Ui[5] * t[2]);
            }
            for (int i = 0; i < n; ++i) if (!isfinite(step[i])) { lin_ok = 0; break; }
        }
        reuse_diagonal = 1;
        sum->num_linear_solves++; sum->linear_solve_time_s += now_s() - t0;

        /* model cost change = -(J s)^T (r + J s / 2), J the scaled Jacobian, s the scaled step */
        double model_cost_change = 0; int step_valid = 0;
        if (lin_ok) {
            const double* sc = step; const double* sp = step + 6 * nc; const double sf = step[n - 1];
            double acc = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc) num_threads(nthreads)
            for (int o = 0; o < nobs; ++o) {
                const double* c = sc + 6 * obs_cam[o]; const double* p = sp + 3 * P.obs_pt[o];
                double m0 = Jf[2 * o] * sf, m1 = Jf[2 * o + 1] * sf;
                for (int k = 0; k < 6; ++k) { m0 += Jc[12 * o + k] * c[k]; m1 += Jc[12 * o + 6 + k] * c[k]; }
                for (int k = 0; k < 3; ++k) { m0 += Jp[6 * o + k] * p[k]; m1 += Jp[6 * o + 3 + k] * p[k]; }
                acc += m0 * (res[2 * o] + 0.5 * m0) + m1 * (res[2 * o + 1] + 0.5 * m1);
            }
            model_cost_change = -acc;
            step_valid = model_cost_change > 0.0;
        }
        if (!step_valid) {
            /* HandleInvalidStep */
            if (++invalid_steps >= opt->max_num_consecutive_invalid_steps) {
                sum->termination_type = 2;
                snprintf(sum->message, sizeof sum->message, "Number of consecutive invalid steps more than max_num_consecutive_invalid_steps: %d", opt->max_num_consecutive_invalid_steps);
                sum->num_unsuccessful_steps++;
                if (trace) { trace[4 * iter] = x_cost; trace[4 * iter + 1] = radius; trace[4 * iter + 2] = 0; trace[4 * iter + 3] = gmax; }
                break;
            }
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
            sum->num_unsuccessful_steps++;
            if (trace) { trace[4 * iter] = x_cost; trace[4 * iter + 1] = radius; trace[4 * iter + 2] = 0; trace[4 * iter + 3] = gmax; }
            if (opt->verbose) printf("iter %3d INVALID step radius %.3e\n", iter, radius);
            continue;
        }
        invalid_steps = 0;

        /* candidate = x + step .* scale ; evaluate cost */
        double step_norm = 0;
        for (int i = 0; i < n; ++i) { const double d = step[i] * scale[i]; xc[i] = x[i] + d; }
        for (int i = 0; i < n; ++i) { const double d = x[i] - xc[i]; step_norm += d * d; }
        step_norm = sqrt(step_norm);
        const double cand_cost = eval_cost(&P, xc, nthreads);
        sum->num_residual_evals++;

        /* ParameterToleranceReached */
        if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
            sum->termination_type = 0; snprintf(sum->message, sizeof sum->message, "Parameter tolerance reached. Relative step_norm: %e <= %e.", step_norm / (x_norm + opt->parameter_tolerance), opt->parameter_tolerance);
            if (trace) { trace[4 * iter] = x_cost; trace[4 * iter + 1] = radius; trace[4 * iter + 2] = 0; trace[4 * iter + 3] = gmax; }
            break;
        }
        /* FunctionToleranceReached */
        const double cost_change = x_cost - cand_cost;
        if (fabs(cost_change) <= opt->function_tolerance * x_cost) {
            sum->termination_type = 0; snprintf(sum->message, sizeof sum->message, "Function tolerance reached. |cost_change|/cost: %e <= %e", fabs(cost_change) / x_cost, opt->function_tolerance);
            if (trace) { trace[4 * iter] = x_cost; trace[4 * iter + 1] = radius; trace[4 * iter + 2] = 0; trace[4 * iter + 3] = gmax; }
            break;
        }
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > opt->min_relative_decrease) {
            /* HandleSuccessfulStep */
            memcpy(x, xc, sizeof(double) * n);
            x_norm = 0; for (int i = 0; i < n; ++i) x_norm += x[i] * x[i]; x_norm = sqrt(x_norm);
            EVAL_JAC();
            if (!jac_ok) { sum->termination_type = 2; snprintf(sum->message, sizeof sum->message, "Residual and Jacobian evaluation failed."); break; }
            scale_columns(&P, scale, Jc, Jp, Jf, nthreads);
            const double q = 2.0 * relative_decrease - 1.0;
            double denom = 1.0 - q * q * q; if (denom < 1.0 / 3.0) denom = 1.0 / 3.0;
            radius = radius / denom; if (radius > opt->max_trust_region_radius) radius = opt->max_trust_region_radius;
            decrease_factor = 2.0; reuse_diagonal = 0;
            sum->num_successful_steps++; accepted = 1;
        } else {
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
            sum->num_unsuccessful_steps++;
        }
        if (trace) { trace[4 * iter] = x_cost; trace[4 * iter + 1] = radius; trace[4 * iter + 2] = accepted; trace[4 * iter + 3] = gmax; }
        if (opt->verbose) printf("iter %3d cost %.9e |g|max %.3e radius %.3e rho %.3e %s\n", iter, x_cost, gmax, radius, relative_decrease, accepted ? "ok" : "rejected");
    }
done:
    sum->num_iterations = iter; sum->final_cost = x_cost;
    memcpy(cams, x, sizeof(double) * 6 * nc); memcpy(pts, x + 6 * nc, sizeof(double) * 3 * np); *focal = x[n - 1];
    free(P.obs_pt); free(x); free(xc); free(res); free(Jc); free(Jp); free(Jf); free(scale); free(diag); free(D); free(g);
    free(step); free(S); free(rhs); free(Uinv); free(gpt); free(pool);
    sum->total_time_s = now_s() - t_start;
    return sum->termination_type;
#undef EVAL_JAC
}
