// orb_math.cuh -- the per-pixel / per-key-point arithmetic of the ORB extraction stage, written once as
// __host__ __device__ functions: orb.cu's kernels call them on the GPU, tests/host_orb_math.cpp compiles the same functions for the
// CPU so that `-m "not gpu"` tests can compare them with the oracle (oracle/orb_oracle.py) without a device.
//
// Reference: `mDetector->detectAndCompute(...)` with `ORB::create(5000)`, SfM2DFeatureUtilities.cpp:39, 46-51.  The arithmetic is OpenCV's
// (un-vendored dependency); every function names the OpenCV routine whose result it has to reproduce BIT FOR BIT.  Floating-point
// expressions are spelled with explicit round-to-nearest operations so that neither nvcc nor the host compiler contracts them
// differently from the binary the parity tests compare with: OpenCV's baseline code (orb.cpp, mathfuncs) is built without fused
// multiply-adds, its AVX2-dispatched separable filter with them.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define ORB_HD __host__ __device__ __forceinline__
#else
#define ORB_HD inline
#endif

namespace orbm {

#if defined(__CUDA_ARCH__)
ORB_HD float mul(float a, float b) { return __fmul_rn(a, b); }
ORB_HD float add(float a, float b) { return __fadd_rn(a, b); }
ORB_HD float sub(float a, float b) { return __fsub_rn(a, b); }
ORB_HD float div(float a, float b) { return __fdiv_rn(a, b); }
ORB_HD float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
ORB_HD int round_even(float a) { return __float2int_rn(a); }
#else
// host build: the translation unit is compiled with -ffp-contract=off / without FMA code generation; volatile keeps every
// intermediate a rounded float even under -ffast-math-free -O3 inlining.
ORB_HD float mul(float a, float b) { volatile float r = a * b; return r; }
ORB_HD float add(float a, float b) { volatile float r = a + b; return r; }
ORB_HD float sub(float a, float b) { volatile float r = a - b; return r; }
ORB_HD float div(float a, float b) { volatile float r = a / b; return r; }
ORB_HD float fma(float a, float b, float c) { return std::fmaf(a, b, c); }
ORB_HD int round_even(float a) { return (int)std::nearbyintf(a); }       // cvRound: ties to even (default rounding mode)
#endif

// cvtColor(COLOR_BGR2GRAY), 8-bit: 15-bit fixed point (imgproc color_rgb: B2Y 3735, G2Y 19235, R2Y 9798).
ORB_HD int gray_from_bgr(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15; }

// resize(INTER_LINEAR_EXACT), 8-bit: horizontal pass in 8.8 fixed point, vertical pass 16.16, rounded.  ax / ay = weight of the right /
// lower tap in 1/256 (0 at the clamped borders).
ORB_HD int resize_linear_exact(int s00, int s01, int s10, int s11, int ax, int ay) {
    const int h0 = (256 - ax) * s00 + ax * s01;
    const int h1 = (256 - ax) * s10 + ax * s11;
    return (int)(((unsigned)(256 - ay) * (unsigned)h0 + (unsigned)ay * (unsigned)h1 + 32768u) >> 16);
}

// FAST-9/16 (fast.cpp + cornerScore<16>, fast_score.cpp).  v = centre, p[16] = the circle of radius 3 in OpenCV's order.
// Returns 0 when the pixel is no corner for `threshold`, else the corner score (>= threshold, <= 254): the largest t' for which the
// pixel is still a corner = max over the 16 arcs of 9 of min(v - p) resp. min(p - v), minus 1.  Branch-free: the minimum (maximum) over
// every window of 9 comes from a doubling network (2, 4, 8, +1), all indices compile-time constants so the values stay in registers.
ORB_HD int fast9_score(int v, const int* p, int threshold) {
    int d[16], lo[16], hi[16], t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - p[k];
#pragma unroll
    for (int k = 0; k < 16; k++) { lo[k] = d[k] < d[(k + 1) & 15] ? d[k] : d[(k + 1) & 15]; hi[k] = d[k] > d[(k + 1) & 15] ? d[k] : d[(k + 1) & 15]; }
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = lo[k] < lo[(k + 2) & 15] ? lo[k] : lo[(k + 2) & 15];
#pragma unroll
    for (int k = 0; k < 16; k++) lo[k] = t[k] < t[(k + 4) & 15] ? t[k] : t[(k + 4) & 15];          // min of d[k .. k+7]
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = hi[k] > hi[(k + 2) & 15] ? hi[k] : hi[(k + 2) & 15];
#pragma unroll
    for (int k = 0; k < 16; k++) hi[k] = t[k] > t[(k + 4) & 15] ? t[k] : t[(k + 4) & 15];          // max of d[k .. k+7]
    int a = -256, b = 256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int e = d[(k + 8) & 15];
        const int mn = lo[k] < e ? lo[k] : e, mx = hi[k] > e ? hi[k] : e;                            // over the arc k .. k+8
        a = mn > a ? mn : a; b = mx < b ? mx : b;
    }
    const int s = (a > -b ? a : -b) - 1;
    return s >= threshold ? s : 0;
}
// Cheap necessary condition on the four compass points of the circle (p[0], p[4], p[8], p[12]): any arc of 9 holds at least two of them.
ORB_HD bool fast9_may_be_corner(int v, int p0, int p4, int p8, int p12, int threshold) {
    const int nd = (v - p0 > threshold) + (v - p4 > threshold) + (v - p8 > threshold) + (v - p12 > threshold);
    const int nb = (p0 - v > threshold) + (p4 - v > threshold) + (p8 - v > threshold) + (p12 - v > threshold);
    return nd >= 2 || nb >= 2;
}

// orb.cpp HarrisResponses (blockSize 7, k = 0.04f): a, b, c = integer sums of Ix*Ix, Iy*Iy, Ix*Iy over the 7x7 window;
// response = ((float)a * b - (float)c * c - k * ((float)a + b) * ((float)a + b)) * scale^4, evaluated left to right in float.
ORB_HD float harris_response(int a, int b, int c) {
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float scale = div(1.0f, mul(28.0f, 255.0f));                       // 1.f / ((1 << 2) * blockSize * 255.f)
    const float ssq = mul(mul(mul(scale, scale), scale), scale);
    const float s = add(fa, fb);
    return mul(sub(sub(mul(fa, fb), mul(fc, fc)), mul(mul(0.04f, s), s)), ssq);
}

// cv::fastAtan2 (mathfuncs_core atan_f32): degrees in [0, 360).  Constants are the float products 0.99978784f * (float)(180 / CV_PI), ...
ORB_HD float fast_atan2(float y, float x) {
    const float p1 = 57.2836266f, p3 = -18.6674461f, p5 = 8.91400051f, p7 = -2.53972459f;   // bit patterns checked in tests
    const float ax = fabsf(x), ay = fabsf(y), eps = 2.220446049250313e-16f;
    float a;
    if (ax >= ay) {
        const float c = div(ay, add(ax, eps)), c2 = mul(c, c);
        a = mul(add(mul(add(mul(add(mul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = div(ax, add(ay, eps)), c2 = mul(c, c);
        a = sub(90.0f, mul(add(mul(add(mul(add(mul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = sub(180.0f, a);
    if (y < 0) a = sub(360.0f, a);
    return a;
}

// The separable Gaussian cv::GaussianBlur(level, 7x7, sigma 2) runs INSIDE ORB (the level is a sub-matrix, so the 8-bit fixed-point
// path is skipped and sepFilter2D works with the float kernel getGaussianKernel(7, 2, CV_32F)):
//   row pass     s = k0*p0; s = fma(k_i, p_i, s), i = 1..6            (RowFilter<uchar,float>, AVX2 build: fused)
//   column pass  s = k3*r0; s = fma(k_{3+j}, r_{+j} + r_{-j}, s)      (SymmColumnFilter<Cast<float,uchar>>), then cvRound + saturate
ORB_HD float gauss_k(int i) {           // i = 0..3 : taps at distance 3, 2, 1, 0 from the centre
    const float k[4] = {0.07015932351350784f, 0.13107487559318542f, 0.1907128244638443f, 0.21610593795776367f};
    return k[i];
}
ORB_HD float blur_row(const int* p /* 7 pixels */) {
    float s = mul(gauss_k(0), (float)p[0]);
    s = fma(gauss_k(1), (float)p[1], s); s = fma(gauss_k(2), (float)p[2], s); s = fma(gauss_k(3), (float)p[3], s);
    s = fma(gauss_k(2), (float)p[4], s); s = fma(gauss_k(1), (float)p[5], s); s = fma(gauss_k(0), (float)p[6], s);
    return s;
}
ORB_HD int blur_col(const float* r /* 7 row-filtered values, r[3] = centre */) {
    float s = mul(gauss_k(3), r[3]);
    s = fma(gauss_k(2), add(r[4], r[2]), s); s = fma(gauss_k(1), add(r[5], r[1]), s); s = fma(gauss_k(0), add(r[6], r[0]), s);
    const int v = round_even(s);
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// orb.cpp computeOrbDescriptors: angle (degrees) -> (cos, sin) as floats of the double functions; a test point (px, py) of the
// pattern goes to (cvRound(px*a - py*b), cvRound(px*b + py*a)), products and sums rounded separately.
ORB_HD void angle_to_cs(float angle_deg, float* a, float* b) {
    const float rad = mul(angle_deg, 0.017453292519943295f);                 // (float)(CV_PI / 180.f)
    *a = (float)cos((double)rad); *b = (float)sin((double)rad);
}
ORB_HD int rot_x(int px, int py, float a, float b) { return round_even(sub(mul((float)px, a), mul((float)py, b))); }
ORB_HD int rot_y(int px, int py, float a, float b) { return round_even(add(mul((float)px, b), mul((float)py, a))); }

// half-width of the circular patch of radius 15 per row (orb.cpp umax)
ORB_HD int umax15(int v) {
    const int u[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    return u[v];
}

}  // namespace orbm
