// tests/host_orb_math.cpp -- csrc/orb_math.cuh (__host__ __device__, the arithmetic of the ORB kernels in csrc/orb.cu) compiled for the
// CPU.  Each function walks an image the way the corresponding kernel does and calls the SAME per-pixel / per-key-point routines, so
// that tests/test_orb_math_host.py can compare them with oracle/orb_oracle.py (and through it with cv2) without a GPU.
#include <cstdint>
#include <cstring>

#include "../sfm-toy-library_b200/csrc/orb_math.cuh"
#include "../sfm-toy-library_b200/csrc/orb_pattern.h"

static inline int reflect101(int i, int n) { i = i < 0 ? -i : i; return i >= n ? 2 * n - 2 - i : i; }

extern "C" {

void orbh_gray(const uint8_t* bgr, int n, uint8_t* out) {
    for (int i = 0; i < n; i++) out[i] = (uint8_t)orbm::gray_from_bgr(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}

void orbh_resize(const uint8_t* s, int sw, const int32_t* x0, const int32_t* x1, const int32_t* ax, const int32_t* y0, const int32_t* y1,
                 const int32_t* ay, int dw, int dh, uint8_t* out) {
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++)
            out[y * dw + x] = (uint8_t)orbm::resize_linear_exact(s[y0[y] * sw + x0[x]], s[y0[y] * sw + x1[x]], s[y1[y] * sw + x0[x]], s[y1[y] * sw + x1[x]],
                                                                 ax[x], ay[y]);
}

void orbh_fast_score_map(const uint8_t* I, int w, int h, int threshold, uint8_t* out) {
    memset(out, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* c = I + y * w + x;
            int p[16] = {c[3 * w], c[3 * w + 1], c[2 * w + 2], c[w + 3], c[3], c[-w + 3], c[-2 * w + 2], c[-3 * w + 1],
                         c[-3 * w], c[-3 * w - 1], c[-2 * w - 2], c[-w - 3], c[-3], c[w - 3], c[2 * w - 2], c[3 * w - 1]};
            out[y * w + x] = orbm::fast9_may_be_corner(c[0], p[0], p[4], p[8], p[12], threshold) ? (uint8_t)orbm::fast9_score(c[0], p, threshold) : 0;
        }
}

void orbh_harris(const uint8_t* I, int w, const int32_t* xs, const int32_t* ys, int n, float* out) {
    for (int i = 0; i < n; i++) {
        int a = 0, b = 0, c = 0;
        for (int k = 0; k < 49; k++) {
            const uint8_t* p = I + (ys[i] - 3 + k / 7) * w + (xs[i] - 3 + k % 7);
            const int ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-w + 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[w - 1]);
            const int iy = ((int)p[w] - (int)p[-w]) * 2 + ((int)p[w - 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[-w + 1]);
            a += ix * ix; b += iy * iy; c += ix * iy;
        }
        out[i] = orbm::harris_response(a, b, c);
    }
}

void orbh_atan2(const float* y, const float* x, int n, float* out) { for (int i = 0; i < n; i++) out[i] = orbm::fast_atan2(y[i], x[i]); }

void orbh_angles(const uint8_t* I, int w, const int32_t* xs, const int32_t* ys, int n, float* out) {
    for (int i = 0; i < n; i++) {
        int m10 = 0, m01 = 0;
        for (int v = -15; v <= 15; v++) {
            const int u = orbm::umax15(v < 0 ? -v : v);
            const uint8_t* row = I + (ys[i] + v) * w + xs[i];
            int sum = 0;
            for (int k = -u; k <= u; k++) { sum += row[k]; m10 += k * row[k]; }
            m01 += v * sum;
        }
        out[i] = orbm::fast_atan2((float)m01, (float)m10);
    }
}

void orbh_blur(const uint8_t* I, int w, int h, uint8_t* out) {
    if (w < 4 || h < 4) { memcpy(out, I, (size_t)w * h); return; }
    float* R = new float[(size_t)w * (h + 6)];
    for (int yy = -3; yy < h + 3; yy++) {
        const uint8_t* row = I + reflect101(yy, h) * w;
        for (int x = 0; x < w; x++) {
            int p[7];
            for (int k = 0; k < 7; k++) p[k] = row[reflect101(x + k - 3, w)];
            R[(size_t)(yy + 3) * w + x] = orbm::blur_row(p);
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float r[7];
            for (int k = 0; k < 7; k++) r[k] = R[(size_t)(y + k) * w + x];
            out[y * w + x] = (uint8_t)orbm::blur_col(r);
        }
    delete[] R;
}

// the describe kernel's per-key-point work: pt = (x, y) * scale, centre = cvRound(pt * inv_scale), 32 descriptor bytes
void orbh_describe(const uint8_t* B, int w, const int32_t* xs, const int32_t* ys, const float* angle, int n, float scale, float inv_scale,
                   float* pt /* [n][2] */, uint8_t* desc /* [n][32] */) {
    for (int i = 0; i < n; i++) {
        const float px = orbm::mul((float)xs[i], scale), py = orbm::mul((float)ys[i], scale);
        pt[2 * i] = px; pt[2 * i + 1] = py;
        const int cx = orbm::round_even(orbm::mul(px, inv_scale)), cy = orbm::round_even(orbm::mul(py, inv_scale));
        float a, b;
        orbm::angle_to_cs(angle[i], &a, &b);
        const uint8_t* C = B + cy * w + cx;
        for (int lane = 0; lane < 32; lane++) {
            unsigned byte = 0;
            for (int bit = 0; bit < 8; bit++) {
                const signed char* q = ORB_BIT_PATTERN_31 + (lane * 8 + bit) * 4;
                const int t0 = C[orbm::rot_y(q[0], q[1], a, b) * w + orbm::rot_x(q[0], q[1], a, b)];
                const int t1 = C[orbm::rot_y(q[2], q[3], a, b) * w + orbm::rot_x(q[2], q[3], a, b)];
                byte |= (unsigned)(t0 < t1) << bit;
            }
            desc[32 * i + lane] = (uint8_t)byte;
        }
    }
}

const signed char* orbh_pattern() { return ORB_BIT_PATTERN_31; }

}  // extern "C"
