/*
 * oracle/triangulate_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement of the reference's two-view triangulation hot path:
 *   SfMStereoUtilities::triangulateViews   (reference SfMToyLib/SfMStereoUtilities.cpp:120-206)
 *   GetAlignedPointsFromMatch (gather)     (reference SfMToyLib/SfMCommon.cpp:63-87)
 *
 * The arithmetic lives in un-vendored OpenCV (>= 3.1, CMakeLists.txt:28).  Call chain restated:
 *   :146-147  cv::undistortPoints(pts, K, no-dist)        -> x_n = (u - cx) * (1/fx)   (double, stored float)
 *   :150      cv::triangulatePoints(Pl, Pr, nl, nr)       -> per point A(4x4, double) rows x*P[2]-P[0], y*P[2]-P[1]
 *                                                            per view; X = right singular vector of the smallest
 *                                                            singular value (one-sided Jacobi SVD in double); float out
 *   :153      cv::convertPointsFromHomogeneous            -> float: scale = (w != 0) ? 1/w : 1 ; X*scale
 *   :155-167  cv::Rodrigues(R)->rvec (float) ; cv::projectPoints(X, rvec, t, K, no-dist)  (double inside, float out)
 *   :184-203  drop if ||proj - x|| > 10 in EITHER view  (norm in double of float differences)
 *
 * Pinned against cv2 4.13 by tests/golden/triangulate_*.npz (made by tests/golden/make_golden.py) and the
 * reference's own unit-test fixture triangulate_from_2_views (SfMUnitTests.cpp:221-251, tolerance 0.01).
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <float.h>

/* --- Rodrigues: rotation matrix -> vector -> matrix round trip (through a float rvec, as at :155-167) --- */

static void rotmat_to_rvec(const double R[9], double r[3]) {
    /* standard log map; OpenCV additionally re-orthonormalises R with an SVD first (difference ~1e-7 for the
       float32 rotation matrices the reference holds, i.e. below the float rvec rounding that follows). */
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1.0 : 1.0);
        t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1.0 : 1.0);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        double n = theta / sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * n; r[1] = ry * n; r[2] = rz * n;
        return;
    }
    double vth = 1.0 / (2.0 * s) * theta;
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

static void rvec_to_rotmat(const double r[3], double R[9]) {
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

/* P (3x4 float, row-major) -> the double [R|t] that cv::projectPoints effectively uses (float rvec round trip). */
void sfm_oracle_pose_roundtrip(const float P[12], double Rt[12]) {
    double R[9], r[3], R2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = (double)P[4 * i + j];
    rotmat_to_rvec(R, r);
    for (int i = 0; i < 3; ++i) r[i] = (double)(float)r[i];   /* rvec Mat is CV_32F */
    rvec_to_rotmat(r, R2);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rt[4 * i + j] = R2[3 * i + j];
        Rt[4 * i + 3] = (double)P[4 * i + 3];
    }
}

/* --- smallest right singular vector of a 4x4 (Hestenes one-sided Jacobi on columns, double) --- */
static void null_vector_4x4(double A[4][4] /* destroyed */, double X[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double nrm[4];
    for (int j = 0; j < 4; ++j) {
        double s = 0; for (int k = 0; k < 4; ++k) s += A[k][j] * A[k][j];
        nrm[j] = s;
    }
    const double eps = DBL_EPSILON * 10;
    for (int sweep = 0; sweep < 30; ++sweep) {
        int rotated = 0;
        for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 4; ++j) {
            double a = nrm[i], b = nrm[j], p = 0;
            for (int k = 0; k < 4; ++k) p += A[k][i] * A[k][j];
            if (fabs(p) <= eps * sqrt(a * b)) continue;
            p *= 2;
            double beta = a - b, gamma = hypot(p, beta), c, s;
            if (beta < 0) { double delta = (gamma - beta) * 0.5; s = sqrt(delta / gamma); c = p / (gamma * s * 2); }
            else { c = sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
            a = b = 0;
            for (int k = 0; k < 4; ++k) {
                double t0 = c * A[k][i] + s * A[k][j], t1 = -s * A[k][i] + c * A[k][j];
                A[k][i] = t0; A[k][j] = t1; a += t0 * t0; b += t1 * t1;
                double v0 = c * V[k][i] + s * V[k][j], v1 = -s * V[k][i] + c * V[k][j];
                V[k][i] = v0; V[k][j] = v1;
            }
            nrm[i] = a; nrm[j] = b; rotated = 1;
        }
        if (!rotated) break;
    }
    int m = 0;
    for (int j = 0; j < 4; ++j) {
        double s = 0; for (int k = 0; k < 4; ++k) s += A[k][j] * A[k][j];
        nrm[j] = s;
    }
    /* OpenCV sorts descending and takes the last row of Vt: with equal norms the LAST such column survives. */
    for (int j = 1; j < 4; ++j) if (nrm[j] <= nrm[m]) m = j;
    for (int k = 0; k < 4; ++k) X[k] = V[k][m];
}

/*
 * Full triangulateViews on flat arrays.
 *   K[9] float row-major; Pl/Pr[12] float row-major; ptsL [nl*2], ptsR [nr*2] float pixel coords;
 *   mq/mt [m] match query/train indices (NULL => identity, i.e. GetAlignedMatching);
 *   out X [m*3] float (all m points, filtered or not), keep [m] (1 = passes the reprojection filter at :186-187),
 *   err [m*2] double reprojection error in each view (may be NULL).
 * Returns the number of kept points.  The reference appends kept points in match order with back references
 * mq[i]/mt[i] (SfMCommon.cpp:81-82, SfMStereoUtilities.cpp:199-202); callers compact with `keep`.
 */
int sfm_oracle_triangulate(const float* K, const float* Pl, const float* Pr,
                           const float* ptsL, const float* ptsR, const int32_t* mq, const int32_t* mt, int m,
                           float max_reproj, float* X, uint8_t* keep, double* err) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    double RtL[12], RtR[12], PL[12], PR[12];
    sfm_oracle_pose_roundtrip(Pl, RtL);
    sfm_oracle_pose_roundtrip(Pr, RtR);
    for (int i = 0; i < 12; ++i) { PL[i] = (double)Pl[i]; PR[i] = (double)Pr[i]; }
    int nkeep = 0;
    for (int i = 0; i < m; ++i) {
        const int iq = mq ? mq[i] : i, it = mt ? mt[i] : i;
        const float ul = ptsL[2 * iq], vl = ptsL[2 * iq + 1], ur = ptsR[2 * it], vr = ptsR[2 * it + 1];
        /* undistortPoints, no distortion: double arithmetic, float result */
        const double xl = (double)(float)(((double)ul - cx) * ifx), yl = (double)(float)(((double)vl - cy) * ify);
        const double xr = (double)(float)(((double)ur - cx) * ifx), yr = (double)(float)(((double)vr - cy) * ify);
        double A[4][4], Xh[4];
        for (int k = 0; k < 4; ++k) {
            A[0][k] = xl * PL[8 + k] - PL[k];
            A[1][k] = yl * PL[8 + k] - PL[4 + k];
            A[2][k] = xr * PR[8 + k] - PR[k];
            A[3][k] = yr * PR[8 + k] - PR[4 + k];
        }
        null_vector_4x4(A, Xh);
        /* triangulatePoints output is float 4xM; convertPointsFromHomogeneous in float */
        const float hx = (float)Xh[0], hy = (float)Xh[1], hz = (float)Xh[2], hw = (float)Xh[3];
        const float sc = hw != 0.f ? 1.f / hw : 1.f;
        const float px = hx * sc, py = hy * sc, pz = hz * sc;
        X[3 * i] = px; X[3 * i + 1] = py; X[3 * i + 2] = pz;
        /* projectPoints (double inside, float out) in both views */
        double e[2];
        for (int v = 0; v < 2; ++v) {
            const double* Rt = v == 0 ? RtL : RtR;
            const double Xc = Rt[0] * px + Rt[1] * py + Rt[2] * pz + Rt[3];
            const double Yc = Rt[4] * px + Rt[5] * py + Rt[6] * pz + Rt[7];
            double Zc = Rt[8] * px + Rt[9] * py + Rt[10] * pz + Rt[11];
            Zc = Zc ? 1.0 / Zc : 1.0;
            const float pu = (float)(Xc * Zc * fx + cx), pv = (float)(Yc * Zc * fy + cy);
            const float du = pu - (v == 0 ? ul : ur), dv = pv - (v == 0 ? vl : vr);
            e[v] = sqrt((double)du * du + (double)dv * dv);
        }
        if (err) { err[2 * i] = e[0]; err[2 * i + 1] = e[1]; }
        const int k = !(e[0] > (double)max_reproj || e[1] > (double)max_reproj);
        keep[i] = (uint8_t)k; nkeep += k;
    }
    return nkeep;
}
