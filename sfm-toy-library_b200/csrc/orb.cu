// orb.cu -- SURVEY.md section 8 row f-3: ORB extraction, the step before matching.
//
// Replaces SfM2DFeatureUtilities::extractFeatures (SfM2DFeatureUtilities.cpp:46-51) = `ORB::create(5000)->detectAndCompute(image,
// noArray(), keyPoints, descriptors)` (:39, :48) for one image or for all images of SfM::extractFeatures (SfM.cpp:141-154) in one
// call.  Results are bit-identical to OpenCV's: key points (pt, size, angle, response, octave) in OpenCV's output order and the
// 256-bit descriptors (tests/test_gpu_orb.py compares with cv2 on real and synthetic images and with oracle/orb_oracle.py stage by stage).
//
// Work split.  Everything that touches pixels runs on the GPU: grey conversion, the 8-level INTER_LINEAR_EXACT pyramid, FAST-9/16 scores
// of every pyramid pixel, 3x3 non-maximum suppression + border filter + ORDERED compaction (raster order inside a level, the order
// cv::FAST emits), Harris responses, intensity-centroid angles, the float separable Gaussian, steered BRIEF.  The two
// `KeyPointsFilter::retainBest` selections stay on the host ON PURPOSE: their result ORDER is whatever libstdc++'s std::nth_element +
// std::partition leave behind, and every later stage (and the matcher's trainIdx) inherits that order -- running the same two
// algorithms on 8-byte (response, index) records reproduces it exactly and costs ~0.1 ms per image.
// All kernels carry the image index in blockIdx.z / the record, so a batch of equally sized images shares every launch and
// every host<->device round trip (3 per batch).
//
// Integer / byte work, HBM- and latency-bound (2.6 MB of pyramid per 1024x768 image): no tensor cores involved.
#include <algorithm>

#include "common.cuh"
#include "orb_math.cuh"
#define ORB_PATTERN_DECL __device__
#include "orb_pattern.h"

namespace {

constexpr int ORB_LEVELS = 8;
constexpr int ORB_EDGE = 31;          // edgeThreshold
constexpr int ORB_FAST_T = 20;        // fastThreshold
constexpr int CHUNK = 256;            // pixels of one row handled by one CTA of the row-chunk kernels

struct OrbLevel { int w, h, off, row0; float scale, inv_scale; int quota, pad; };
struct OrbLayout {
    OrbLevel lv[ORB_LEVELS];
    int total_rows;      // rows of all levels
    int slab;            // bytes of one image's pyramid (levels back to back, each 16-byte aligned)
    int chunks;          // CTAs per row = ceil(level-0 width / CHUNK)
    int ncnt;            // total_rows * chunks
};

// ---------------------------------------------------------------------------------------------------- host-side layout (pure functions)
// ORB_Impl::detectAndCompute: layerScale[l] = (float)pow((double)1.2f, l); layer size = cvRound(cols / scale) x cvRound(rows / scale);
// computeKeyPoints: nfeaturesPerLevel from the geometric series with factor 1 / 1.2f.
int make_layout(int w, int h, int nfeatures, OrbLayout& L) {
    const double sf = (double)1.2f;
    int off = 0, row = 0;
    const float factor = (float)(1.0 / sf);
    float ndesired = (float)nfeatures * (1.0f - factor) / (1.0f - (float)std::pow((double)factor, (double)ORB_LEVELS));
    int sum = 0;
    for (int l = 0; l < ORB_LEVELS; l++) {
        OrbLevel& v = L.lv[l];
        v.scale = (float)std::pow(sf, (double)l);
        v.inv_scale = 1.0f / v.scale;
        v.w = (int)std::nearbyintf((float)w / v.scale); v.h = (int)std::nearbyintf((float)h / v.scale);
        if (v.w < 0) v.w = 0;
        if (v.h < 0) v.h = 0;
        if (v.w == 0 || v.h == 0) v.w = v.h = 0;
        v.off = off; v.row0 = row; v.pad = 0;
        off += (v.w * v.h + 15) & ~15; row += v.h;
        if (l < ORB_LEVELS - 1) { v.quota = (int)std::nearbyintf(ndesired); sum += v.quota; ndesired *= factor; }
        else v.quota = std::max(nfeatures - sum, 0);
    }
    L.total_rows = row; L.slab = (off + 255) & ~255; L.chunks = ceil_div(w, CHUNK); L.ncnt = L.total_rows * L.chunks;
    return 0;
}

// resize(INTER_LINEAR_EXACT) tap tables: for destination index d, the two source indices and the weight (1/256) of the second one.
void linear_exact_taps(int src, int dst, int32_t* i0, int32_t* i1, int32_t* a) {
    const double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; d++) {
        const double f = scale * ((double)d + 0.5) - 0.5;
        const int i = (int)std::floor(f);
        int w = (int)std::nearbyint((f - (double)i) * 256.0);
        int lo = i, hi = i + 1;
        if (i < 0) { lo = hi = 0; w = 0; }
        if (i >= src - 1) { lo = hi = src - 1; w = 0; }
        i0[d] = lo; i1[d] = hi; a[d] = w;
    }
}

// KeyPointsFilter::retainBest on (response, index) records: same std algorithms, same comparator results, hence the same permutation.
struct Rec { float response; int32_t idx; };
struct RecGreater { bool operator()(const Rec& a, const Rec& b) const { return a.response > b.response; } };
struct RecAtLeast { float v; bool operator()(const Rec& k) const { return k.response >= v; } };
void retain_best(std::vector<Rec>& k, int n_points) {
    if (n_points >= 0 && k.size() > (size_t)n_points) {
        if (n_points == 0) { k.clear(); return; }
        std::nth_element(k.begin(), k.begin() + n_points - 1, k.end(), RecGreater());
        const float ambiguous = k[n_points - 1].response;
        auto new_end = std::partition(k.begin() + n_points, k.end(), RecAtLeast{ambiguous});
        k.resize(new_end - k.begin());
    }
}

// ---------------------------------------------------------------------------------------------------- kernels
__device__ __forceinline__ int level_of_row(const OrbLayout& L, int row) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_LEVELS; i++) l += (row >= L.lv[i].row0 && L.lv[i].h > 0) ? 1 : 0;   // levels are non-increasing, empty ones at the end
    return l;
}

__global__ void __launch_bounds__(256) orb_gray_kernel(const uint8_t* __restrict__ raw, size_t raw_stride_img, int row_stride, int w, int h,
                                                        uint8_t* __restrict__ pyr, int slab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= w) return;
    const uint8_t* p = raw + (size_t)img * raw_stride_img + (size_t)y * row_stride + 3 * x;
    pyr[(size_t)img * slab + (size_t)y * w + x] = (uint8_t)orbm::gray_from_bgr(p[0], p[1], p[2]);
}

__global__ void __launch_bounds__(256) orb_resize_kernel(uint8_t* pyr, int slab, int src_off, int sw, int dst_off, int dw, int dh,
                                                          const int32_t* __restrict__ taps /* x: i0,i1,a [dw each]; y: i0,i1,a [dh each] */) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= dw) return;
    const int x0 = taps[x], x1 = taps[dw + x], ax = taps[2 * dw + x];
    const int32_t* ty = taps + 3 * dw;
    const int y0 = ty[y], y1 = ty[dh + y], ay = ty[2 * dh + y];
    const uint8_t* s = pyr + (size_t)img * slab + src_off;
    const int v = orbm::resize_linear_exact(s[(size_t)y0 * sw + x0], s[(size_t)y0 * sw + x1], s[(size_t)y1 * sw + x0], s[(size_t)y1 * sw + x1], ax, ay);
    pyr[(size_t)img * slab + dst_off + (size_t)y * dw + x] = (uint8_t)v;
}

// FAST-9/16 score of every pyramid pixel (0 = no corner).  grid = (chunks, total_rows, images).
__global__ void __launch_bounds__(CHUNK) orb_fast_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score, OrbLayout L) {
    const int row = blockIdx.y, img = blockIdx.z;
    const int l = level_of_row(L, row);
    const int w = L.lv[l].w, h = L.lv[l].h, y = row - L.lv[l].row0, x = blockIdx.x * CHUNK + threadIdx.x;
    if (x >= w) return;
    const size_t base = (size_t)img * L.slab + L.lv[l].off;
    int s = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const uint8_t* c = pyr + base + (size_t)y * w + x;
        const int v = c[0];
        int p[16];
        p[0] = c[3 * w];          p[1] = c[3 * w + 1];   p[2] = c[2 * w + 2];   p[3] = c[w + 3];
        p[4] = c[3];              p[5] = c[-w + 3];      p[6] = c[-2 * w + 2];  p[7] = c[-3 * w + 1];
        p[8] = c[-3 * w];         p[9] = c[-3 * w - 1];  p[10] = c[-2 * w - 2]; p[11] = c[-w - 3];
        p[12] = c[-3];            p[13] = c[w - 3];      p[14] = c[2 * w - 2];  p[15] = c[3 * w - 1];
        s = orbm::fast9_score(v, p, ORB_FAST_T);
    }
    score[base + (size_t)y * w + x] = (uint8_t)s;
}

// fast.cpp non-maximum suppression (strictly greater than the 8 neighbours) + KeyPointsFilter::runByImageBorder(edgeThreshold)
__device__ __forceinline__ int nms_keep(const uint8_t* __restrict__ sc, int w, int h, int x, int y) {
    if (x < ORB_EDGE || x >= w - ORB_EDGE || y < ORB_EDGE || y >= h - ORB_EDGE) return 0;     // the border filter implies 1 <= x < w-1, ...
    const uint8_t* c = sc + (size_t)y * w + x;
    const int s = c[0];
    if (s == 0) return 0;
    return (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] && s > c[w] && s > c[w + 1]) ? s : 0;
}

// pass 1: survivors per (row, chunk).  grid = (chunks, total_rows, images)
__global__ void __launch_bounds__(CHUNK) orb_nms_count_kernel(const uint8_t* __restrict__ score, OrbLayout L, int32_t* __restrict__ cnt) {
    const int row = blockIdx.y, img = blockIdx.z;
    const int l = level_of_row(L, row);
    const int w = L.lv[l].w, h = L.lv[l].h, y = row - L.lv[l].row0, x = blockIdx.x * CHUNK + threadIdx.x;
    const int keep = (x < w) ? (nms_keep(score + (size_t)img * L.slab + L.lv[l].off, w, h, x, y) != 0) : 0;
    const int n = __syncthreads_count(keep);
    if (threadIdx.x == 0) cnt[(size_t)img * L.ncnt + (size_t)row * L.chunks + blockIdx.x] = n;
}

// pass 2: exclusive scan of the counts of one image (one CTA per image); also the start of every level and the total.
__global__ void __launch_bounds__(1024) orb_scan_kernel(const int32_t* __restrict__ cnt, int32_t* __restrict__ off, int32_t* __restrict__ lvl_start, OrbLayout L) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t* c = cnt + (size_t)img * L.ncnt; int32_t* o = off + (size_t)img * L.ncnt;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < L.ncnt; base += 1024) {
        const int i = base + tid;
        const int v = i < L.ncnt ? c[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += t; }
        if (lane == 31) warp_sum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int t = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += u; }
            warp_sum[lane] = t;
        }
        __syncthreads();
        const int before = carry + (wid ? warp_sum[wid - 1] : 0) + s - v;
        if (i < L.ncnt) o[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid < ORB_LEVELS) lvl_start[img * 16 + tid] = (L.lv[tid].h > 0 && L.lv[tid].row0 * L.chunks < L.ncnt) ? o[(size_t)L.lv[tid].row0 * L.chunks] : carry;
    if (tid == ORB_LEVELS) lvl_start[img * 16 + ORB_LEVELS] = carry;
}

// pass 3: ordered scatter -> candidates (x | y << 16, score) in raster order per level, levels back to back.
__global__ void __launch_bounds__(CHUNK) orb_nms_scatter_kernel(const uint8_t* __restrict__ score, OrbLayout L, const int32_t* __restrict__ off,
                                                                 uint2* __restrict__ cand, int cand_cap) {
    __shared__ int warp_cnt[CHUNK / 32];
    const int row = blockIdx.y, img = blockIdx.z, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int l = level_of_row(L, row);
    const int w = L.lv[l].w, h = L.lv[l].h, y = row - L.lv[l].row0, x = blockIdx.x * CHUNK + threadIdx.x;
    const int s = (x < w) ? nms_keep(score + (size_t)img * L.slab + L.lv[l].off, w, h, x, y) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, s != 0);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    if (s) {
        int pos = off[(size_t)img * L.ncnt + (size_t)row * L.chunks + blockIdx.x] + __popc(m & ((1u << lane) - 1u));
        for (int i = 0; i < wid; i++) pos += warp_cnt[i];
        if (pos < cand_cap) cand[(size_t)img * cand_cap + pos] = make_uint2((unsigned)x | ((unsigned)y << 16), (unsigned)s | ((unsigned)l << 16));
    }
}

// Harris response of selected candidates: one warp per record {x | y << 16, level | img << 8}.
__global__ void __launch_bounds__(256) orb_harris_kernel(const uint8_t* __restrict__ pyr, OrbLayout L, const uint2* __restrict__ sel, int n,
                                                          float* __restrict__ response) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint2 r = sel[i];
    const int x = r.x & 0xFFFF, y = r.x >> 16, l = r.y & 0xFF, img = r.y >> 8;
    const int w = L.lv[l].w;
    const uint8_t* I = pyr + (size_t)img * L.slab + L.lv[l].off;
    int a = 0, b = 0, c = 0;
    for (int k = lane; k < 49; k += 32) {
        const uint8_t* p = I + (size_t)(y - 3 + k / 7) * w + (x - 3 + k % 7);
        const int ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-w + 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[w - 1]);
        const int iy = ((int)p[w] - (int)p[-w]) * 2 + ((int)p[w - 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[-w + 1]);
        a += ix * ix; b += iy * iy; c += ix * iy;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); c += __shfl_xor_sync(0xffffffffu, c, d); }
    if (lane == 0) response[i] = orbm::harris_response(a, b, c);
}

// Float separable Gaussian of every level (BORDER_REFLECT_101).  Tile 64 x 16 per CTA; grid = (ceil(w0/64), ceil(h_l/16) summed over levels, images)
constexpr int BT_W = 64, BT_H = 16;
__device__ __forceinline__ int reflect101(int i, int n) { i = i < 0 ? -i : i; return i >= n ? 2 * n - 2 - i : i; }
__global__ void __launch_bounds__(256) orb_blur_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, OrbLayout L,
                                                        const int32_t* __restrict__ tile_lvl /* [grid.y] level of the tile row */,
                                                        const int32_t* __restrict__ tile_y0 /* [grid.y] first row of the tile */) {
    __shared__ float R[BT_H + 6][BT_W];
    const int l = tile_lvl[blockIdx.y], y0 = tile_y0[blockIdx.y], x0 = blockIdx.x * BT_W, img = blockIdx.z;
    const int w = L.lv[l].w, h = L.lv[l].h;
    if (x0 >= w) return;
    const uint8_t* I = pyr + (size_t)img * L.slab + L.lv[l].off;
    uint8_t* O = blur + (size_t)img * L.slab + L.lv[l].off;
    if (w < 4 || h < 4) {                               // too small to reflect: copy (such a level cannot hold a key point anyway)
        for (int t = threadIdx.x; t < BT_W * BT_H; t += 256) {
            const int x = x0 + t % BT_W, y = y0 + t / BT_W;
            if (x < w && y < h) O[(size_t)y * w + x] = I[(size_t)y * w + x];
        }
        return;
    }
    for (int t = threadIdx.x; t < BT_W * (BT_H + 6); t += 256) {
        const int tx = t % BT_W, ty = t / BT_W, x = x0 + tx, y = reflect101(y0 + ty - 3, h);
        float v = 0.f;
        if (x < w && y0 + ty - 3 < h + 3) {
            const uint8_t* row = I + (size_t)y * w;
            int p[7];
#pragma unroll
            for (int k = 0; k < 7; k++) p[k] = row[reflect101(x + k - 3, w)];
            v = orbm::blur_row(p);
        }
        R[ty][tx] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < BT_W * BT_H; t += 256) {
        const int tx = t % BT_W, ty = t / BT_W, x = x0 + tx, y = y0 + ty;
        if (x < w && y < h) {
            float r[7];
#pragma unroll
            for (int k = 0; k < 7; k++) r[k] = R[ty + k][tx];
            O[(size_t)y * w + x] = (uint8_t)orbm::blur_col(r);
        }
    }
}

// cv::KeyPoint, 28 bytes: the shim can copy these straight into std::vector<cv::KeyPoint>
struct KeyPointOut { float x, y, size, angle, response; int32_t octave, class_id; };

// ICAngles + computeOrbDescriptors + the final `pt *= scale`: one warp per key point record {x | y << 16, level | img << 8}.
__global__ void __launch_bounds__(256) orb_describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, OrbLayout L,
                                                            const uint2* __restrict__ sel, const float* __restrict__ response, int n,
                                                            KeyPointOut* __restrict__ kp, uint8_t* __restrict__ desc) {
    __shared__ signed char pat[1024];
    for (int t = threadIdx.x; t < 1024; t += blockDim.x) pat[t] = ORB_BIT_PATTERN_31[t];
    __syncthreads();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint2 r = sel[i];
    const int x = r.x & 0xFFFF, y = r.x >> 16, l = r.y & 0xFF, img = r.y >> 8;
    const int w = L.lv[l].w;
    const uint8_t* I = pyr + (size_t)img * L.slab + L.lv[l].off;
    // intensity centroid over the circular patch of radius 15: lane = row v + 15
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15, u = orbm::umax15(v < 0 ? -v : v);
        const uint8_t* row = I + (size_t)(y + v) * w + x;
        int sum = 0;
        for (int k = -u; k <= u; k++) { const int p = row[k]; sum += p; m10 += k * p; }
        m01 = v * sum;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, d); m01 += __shfl_xor_sync(0xffffffffu, m01, d); }
    const float angle = orbm::fast_atan2((float)m01, (float)m10);
    const float scale = L.lv[l].scale, inv = L.lv[l].inv_scale;
    const float px = orbm::mul((float)x, scale), py = orbm::mul((float)y, scale);            // keypoints[i].pt *= scale
    if (lane == 0) {
        KeyPointOut o; o.x = px; o.y = py; o.size = orbm::mul(31.0f, scale); o.angle = angle; o.response = response[i]; o.octave = l; o.class_id = -1;
        kp[i] = o;
    }
    // steered BRIEF on the blurred level, centre = (cvRound(pt.x * (1/scale)), cvRound(pt.y * (1/scale))); lane = descriptor byte
    const int cx = orbm::round_even(orbm::mul(px, inv)), cy = orbm::round_even(orbm::mul(py, inv));
    float a, b;
    orbm::angle_to_cs(angle, &a, &b);
    const uint8_t* C = blur + (size_t)img * L.slab + L.lv[l].off + (size_t)cy * w + cx;
    unsigned byte = 0;
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
        const signed char* q = pat + (lane * 8 + bit) * 4;
        const int t0 = C[orbm::rot_y(q[0], q[1], a, b) * w + orbm::rot_x(q[0], q[1], a, b)];
        const int t1 = C[orbm::rot_y(q[2], q[3], a, b) * w + orbm::rot_x(q[2], q[3], a, b)];
        byte |= (unsigned)(t0 < t1) << bit;
    }
    desc[(size_t)i * 32 + lane] = (uint8_t)byte;
}


// ---------------------------------------------------------------------------------------------------- host driver
struct OrbPlan {
    OrbLayout L;
    std::vector<int32_t> taps;        // per level l >= 1: x taps (3 * w_l) then y taps (3 * h_l)
    int taps_off[ORB_LEVELS];
    std::vector<int32_t> tile_lvl, tile_y0;   // blur kernel: level and first row of every tile row
};

void make_plan(int w, int h, int nfeatures, OrbPlan& P) {
    make_layout(w, h, nfeatures, P.L);
    P.taps.clear();
    for (int l = 0; l < ORB_LEVELS; l++) {
        P.taps_off[l] = (int)P.taps.size();
        const OrbLevel& d = P.L.lv[l];
        if (l == 0 || d.w == 0) continue;
        const OrbLevel& s = P.L.lv[l - 1];
        const size_t at = P.taps.size();
        P.taps.resize(at + 3 * (size_t)d.w + 3 * (size_t)d.h);
        int32_t* t = P.taps.data() + at;
        linear_exact_taps(s.w, d.w, t, t + d.w, t + 2 * d.w);
        t += 3 * d.w;
        linear_exact_taps(s.h, d.h, t, t + d.h, t + 2 * d.h);
    }
    P.tile_lvl.clear(); P.tile_y0.clear();
    for (int l = 0; l < ORB_LEVELS; l++)
        for (int y = 0; y < P.L.lv[l].h; y += BT_H) { P.tile_lvl.push_back(l); P.tile_y0.push_back(y); }
}

int orb_run(sfmb200_ctx* ctx, const uint8_t* const* images, int n_images, int w, int h, int channels, size_t row_stride, int nfeatures,
            int cap, KeyPointOut* kp_out, uint8_t* desc_out, int32_t* n_out) {
    OrbPlan P;
    make_plan(w, h, nfeatures, P);
    const OrbLayout& L = P.L;
    if (L.total_rows > 65535) return sfmb200_fail(ctx, SFMB200_ERR_UNSUPPORTED, "image too tall for one launch (%d pyramid rows)", L.total_rows);
    cudaStream_t st = ctx->stream;
    if (!ctx->orb_stream) {
        SFM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->orb_stream, cudaStreamNonBlocking));
        for (auto& ev : ctx->orb_ev) SFM_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    const size_t raw_img = channels == 3 ? (size_t)3 * w * h : 0;
    const int cand_cap = L.slab / 4 + 64;                    // 3x3 non-maximum suppression keeps at most one pixel of every 2x2 block
    const size_t per_img = Carver::pad(raw_img) + 3 * Carver::pad(L.slab) + 2 * Carver::pad(4 * (size_t)L.ncnt) + Carver::pad(64) +
                           Carver::pad(8 * (size_t)cand_cap);
    const int slots = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_images, ((size_t)3 << 30) / per_img));
    const size_t fixed = Carver::pad(4 * P.taps.size() + 16) + 2 * Carver::pad(4 * P.tile_lvl.size() + 16) + 4096;
    SFM_CUDA(ctx, ctx->orb_dev.reserve(fixed + per_img * slots));
    Carver cv(ctx->orb_dev.p);
    int32_t* d_taps = cv.take<int32_t>(P.taps.size() + 4);
    int32_t* d_tile_lvl = cv.take<int32_t>(P.tile_lvl.size() + 4); int32_t* d_tile_y0 = cv.take<int32_t>(P.tile_y0.size() + 4);
    uint8_t* d_raw = cv.take<uint8_t>(raw_img * slots);
    uint8_t* d_pyr = cv.take<uint8_t>((size_t)L.slab * slots); uint8_t* d_blur = cv.take<uint8_t>((size_t)L.slab * slots);
    uint8_t* d_score = cv.take<uint8_t>((size_t)L.slab * slots);
    int32_t* d_cnt = cv.take<int32_t>((size_t)L.ncnt * slots); int32_t* d_off = cv.take<int32_t>((size_t)L.ncnt * slots);
    int32_t* d_lvl = cv.take<int32_t>(16 * (size_t)slots);
    uint2* d_cand = cv.take<uint2>((size_t)cand_cap * slots);
    ctx->orb_last = OrbLast{d_pyr, d_blur, d_score, L.slab, 0, w, h, nfeatures};
    if (!P.taps.empty()) SFM_CUDA(ctx, cudaMemcpyAsync(d_taps, P.taps.data(), 4 * P.taps.size(), cudaMemcpyHostToDevice, st));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_tile_lvl, P.tile_lvl.data(), 4 * P.tile_lvl.size(), cudaMemcpyHostToDevice, st));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_tile_y0, P.tile_y0.data(), 4 * P.tile_y0.size(), cudaMemcpyHostToDevice, st));

    std::vector<Rec> recs;
    std::vector<uint2> sel;                 // records handed to the Harris / describe kernels
    std::vector<int> seg;                   // per (slot, level): number of records
    for (int i0 = 0; i0 < n_images; i0 += slots) {
        const int nb = std::min(slots, n_images - i0);
        // ---- upload, grey, pyramid, FAST, ordered compaction, blur: one stream, no host involvement
        for (int s = 0; s < nb; s++) {
            if (channels == 1)
                SFM_CUDA(ctx, cudaMemcpy2DAsync(d_pyr + (size_t)s * L.slab, w, images[i0 + s], row_stride, w, h, cudaMemcpyHostToDevice, st));
            else
                SFM_CUDA(ctx, cudaMemcpy2DAsync(d_raw + (size_t)s * raw_img, (size_t)3 * w, images[i0 + s], row_stride, (size_t)3 * w, h, cudaMemcpyHostToDevice, st));
        }
        if (channels == 3) {
            orb_gray_kernel<<<dim3(ceil_div(w, 256), h, nb), 256, 0, st>>>(d_raw, raw_img, 3 * w, w, h, d_pyr, L.slab);
            SFM_LAUNCH_CHECK(ctx);
        }
        for (int l = 1; l < ORB_LEVELS; l++) {
            const OrbLevel& d = L.lv[l]; const OrbLevel& s = L.lv[l - 1];
            if (d.w == 0) break;
            orb_resize_kernel<<<dim3(ceil_div(d.w, 256), d.h, nb), 256, 0, st>>>(d_pyr, L.slab, s.off, s.w, d.off, d.w, d.h, d_taps + P.taps_off[l]);
            SFM_LAUNCH_CHECK(ctx);
        }
        // the blur only needs the pyramid: second stream, joined again before the descriptors
        SFM_CUDA(ctx, cudaEventRecord(ctx->orb_ev[0], st));
        SFM_CUDA(ctx, cudaStreamWaitEvent(ctx->orb_stream, ctx->orb_ev[0], 0));
        orb_blur_kernel<<<dim3(ceil_div(w, BT_W), (unsigned)P.tile_lvl.size(), nb), 256, 0, ctx->orb_stream>>>(d_pyr, d_blur, L, d_tile_lvl, d_tile_y0);
        SFM_LAUNCH_CHECK(ctx);
        SFM_CUDA(ctx, cudaEventRecord(ctx->orb_ev[1], ctx->orb_stream));
        const dim3 grid_rows(L.chunks, L.total_rows, nb);
        orb_fast_kernel<<<grid_rows, CHUNK, 0, st>>>(d_pyr, d_score, L); SFM_LAUNCH_CHECK(ctx);
        orb_nms_count_kernel<<<grid_rows, CHUNK, 0, st>>>(d_score, L, d_cnt); SFM_LAUNCH_CHECK(ctx);
        orb_scan_kernel<<<nb, 1024, 0, st>>>(d_cnt, d_off, d_lvl, L); SFM_LAUNCH_CHECK(ctx);
        orb_nms_scatter_kernel<<<grid_rows, CHUNK, 0, st>>>(d_score, L, d_off, d_cand, cand_cap); SFM_LAUNCH_CHECK(ctx);
        SFM_CUDA(ctx, ctx->orb_pin.reserve(64 * (size_t)nb + 64));
        int32_t* h_lvl = (int32_t*)ctx->orb_pin.p;
        SFM_CUDA(ctx, cudaMemcpyAsync(h_lvl, d_lvl, 64 * (size_t)nb, cudaMemcpyDeviceToHost, st));
        SFM_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->orb_last.nimg = nb;
        // ---- round trip 1: candidates of every image (raster order per level)
        std::vector<int32_t> lvl(h_lvl, h_lvl + 16 * (size_t)nb);      // the pinned buffer is re-carved below
        size_t total_cand = 0;
        for (int s = 0; s < nb; s++) {
            if (lvl[16 * s + ORB_LEVELS] > cand_cap) return sfmb200_fail(ctx, SFMB200_ERR_CUDA, "candidate buffer overflow (internal)");
            total_cand += lvl[16 * s + ORB_LEVELS];
        }
        SFM_CUDA(ctx, ctx->orb_pin.reserve(8 * total_cand + 4 * total_cand + 64 + (size_t)nb * 64));
        uint2* h_cand = (uint2*)ctx->orb_pin.p;
        {
            size_t at = 0;
            for (int s = 0; s < nb; s++) {
                const int n = lvl[16 * s + ORB_LEVELS];
                if (n) SFM_CUDA(ctx, cudaMemcpyAsync(h_cand + at, d_cand + (size_t)s * cand_cap, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
                at += n;
            }
            SFM_CUDA(ctx, cudaStreamSynchronize(st));
        }
        // ---- first selection: best 2 * quota FAST scores per level (KeyPointsFilter::retainBest keeps ties)
        sel.clear(); seg.assign((size_t)nb * ORB_LEVELS, 0);
        {
            size_t at = 0;
            for (int s = 0; s < nb; s++) {
                for (int l = 0; l < ORB_LEVELS; l++) {
                    const int b = lvl[16 * s + l], e = lvl[16 * s + l + 1];
                    recs.resize(e - b);
                    for (int k = b; k < e; k++) recs[k - b] = Rec{(float)(h_cand[at + k].y & 0xFFFF), k};
                    retain_best(recs, 2 * L.lv[l].quota);
                    for (const Rec& r : recs) sel.push_back(make_uint2(h_cand[at + r.idx].x, (unsigned)l | ((unsigned)s << 8)));
                    seg[(size_t)s * ORB_LEVELS + l] = (int)recs.size();
                }
                at += lvl[16 * s + ORB_LEVELS];
            }
        }
        const size_t nsel = sel.size();
        SFM_CUDA(ctx, ctx->orb_lists.reserve(Carver::pad(8 * nsel + 16) * 2 + Carver::pad(4 * nsel + 16) * 2 + Carver::pad(28 * nsel + 32) + Carver::pad(32 * nsel + 32) + 4096));
        Carver lc(ctx->orb_lists.p);
        uint2* d_sel = lc.take<uint2>(nsel + 2); float* d_resp = lc.take<float>(nsel + 4);
        uint2* d_fin = lc.take<uint2>(nsel + 2); float* d_fresp = lc.take<float>(nsel + 4);
        KeyPointOut* d_kp = lc.take<KeyPointOut>(nsel + 1); uint8_t* d_desc = lc.take<uint8_t>(32 * nsel + 32);
        std::vector<float> resp(nsel);
        if (nsel) {
            SFM_CUDA(ctx, cudaMemcpyAsync(d_sel, sel.data(), 8 * nsel, cudaMemcpyHostToDevice, st));
            orb_harris_kernel<<<(unsigned)ceil_div64((int64_t)nsel * 32, 256), 256, 0, st>>>(d_pyr, L, d_sel, (int)nsel, d_resp);
            SFM_LAUNCH_CHECK(ctx);
            SFM_CUDA(ctx, cudaMemcpyAsync(resp.data(), d_resp, 4 * nsel, cudaMemcpyDeviceToHost, st));
            SFM_CUDA(ctx, cudaStreamSynchronize(st));                                              // round trip 2
        }
        // ---- second selection: best quota Harris responses per level
        std::vector<uint2> fin; std::vector<float> fresp; std::vector<int> n_img(nb, 0);
        fin.reserve(nsel); fresp.reserve(nsel);
        {
            size_t at = 0;
            for (int s = 0; s < nb; s++)
                for (int l = 0; l < ORB_LEVELS; l++) {
                    const int n = seg[(size_t)s * ORB_LEVELS + l];
                    recs.resize(n);
                    for (int k = 0; k < n; k++) recs[k] = Rec{resp[at + k], k};
                    retain_best(recs, L.lv[l].quota);
                    for (const Rec& r : recs) { fin.push_back(sel[at + r.idx]); fresp.push_back(r.response); }
                    n_img[s] += (int)recs.size();
                    at += n;
                }
        }
        const size_t nfin = fin.size();
        if (nfin) {
            SFM_CUDA(ctx, ctx->orb_pin.reserve((28 + 32) * nfin + 256));
            KeyPointOut* h_kp = (KeyPointOut*)ctx->orb_pin.p; uint8_t* h_desc = (uint8_t*)(h_kp + nfin);
            SFM_CUDA(ctx, cudaMemcpyAsync(d_fin, fin.data(), 8 * nfin, cudaMemcpyHostToDevice, st));
            SFM_CUDA(ctx, cudaMemcpyAsync(d_fresp, fresp.data(), 4 * nfin, cudaMemcpyHostToDevice, st));
            SFM_CUDA(ctx, cudaStreamWaitEvent(st, ctx->orb_ev[1], 0));
            orb_describe_kernel<<<(unsigned)ceil_div64((int64_t)nfin * 32, 256), 256, 0, st>>>(d_pyr, d_blur, L, d_fin, d_fresp, (int)nfin, d_kp, d_desc);
            SFM_LAUNCH_CHECK(ctx);
            SFM_CUDA(ctx, cudaMemcpyAsync(h_kp, d_kp, 28 * nfin, cudaMemcpyDeviceToHost, st));
            SFM_CUDA(ctx, cudaMemcpyAsync(h_desc, d_desc, 32 * nfin, cudaMemcpyDeviceToHost, st));
            SFM_CUDA(ctx, cudaStreamSynchronize(st));                                              // round trip 3
            size_t at = 0;
            for (int s = 0; s < nb; s++) {
                const int n = std::min(n_img[s], cap);
                if (n > 0) {
                    memcpy(kp_out + (size_t)(i0 + s) * cap, h_kp + at, 28 * (size_t)n);
                    memcpy(desc_out + (size_t)(i0 + s) * cap * 32, h_desc + 32 * at, 32 * (size_t)n);
                }
                at += n_img[s];
            }
        }
        for (int s = 0; s < nb; s++) n_out[i0 + s] = n_img[s];
        SFM_CUDA(ctx, cudaStreamSynchronize(ctx->orb_stream));        // the next batch (and a download for inspection) reuses the blur buffer
    }
    return SFMB200_OK;
}

}  // namespace

extern "C" {

int sfmb200_orb_detect_and_compute_batch(sfmb200_ctx* ctx, const uint8_t* const* images, int n_images, int width, int height, int channels,
                                         size_t row_stride, int nfeatures, int max_keypoints, sfmb200_keypoint* keypoints,
                                         uint8_t* descriptors, int32_t* n_keypoints) {
    static_assert(sizeof(sfmb200_keypoint) == 28 && sizeof(KeyPointOut) == 28, "cv::KeyPoint layout");
    if (!ctx || n_images < 0 || !n_keypoints) return SFMB200_ERR_INVALID;
    if (n_images == 0) return SFMB200_OK;
    if (!images || !keypoints || !descriptors || max_keypoints < 0) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "null buffer");
    if (channels != 1 && channels != 3) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "channels must be 1 (grey) or 3 (BGR), got %d", channels);
    if (width < 8 || height < 8 || width > 65535 || height > 65535)
        return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "image size %dx%d outside [8, 65535]", width, height);
    if (row_stride == 0) row_stride = (size_t)width * channels;
    if (row_stride < (size_t)width * channels) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "row_stride smaller than a row");
    if (nfeatures < 0) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "nfeatures < 0");
    for (int i = 0; i < n_images; i++) if (!images[i]) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "image %d is NULL", i);
    std::lock_guard<std::mutex> lk(ctx->mu);
    SFM_CUDA(ctx, cudaSetDevice(ctx->device));
    return orb_run(ctx, images, n_images, width, height, channels, row_stride, nfeatures, max_keypoints, (KeyPointOut*)keypoints, descriptors, n_keypoints);
}

int sfmb200_orb_detect_and_compute(sfmb200_ctx* ctx, const uint8_t* image, int width, int height, int channels, size_t row_stride, int nfeatures,
                                   int max_keypoints, sfmb200_keypoint* keypoints, uint8_t* descriptors, int32_t* n_keypoints) {
    const uint8_t* one[1] = {image};
    return sfmb200_orb_detect_and_compute_batch(ctx, one, 1, width, height, channels, row_stride, nfeatures, max_keypoints, keypoints, descriptors, n_keypoints);
}

int sfmb200_orb_layout(int width, int height, int nfeatures, int32_t* level_w, int32_t* level_h, float* level_scale, int32_t* level_quota) {
    if (width < 1 || height < 1 || nfeatures < 0) return SFMB200_ERR_INVALID;
    OrbLayout L; make_layout(width, height, nfeatures, L);
    for (int l = 0; l < ORB_LEVELS; l++) {
        if (level_w) level_w[l] = L.lv[l].w;
        if (level_h) level_h[l] = L.lv[l].h;
        if (level_scale) level_scale[l] = L.lv[l].scale;
        if (level_quota) level_quota[l] = L.lv[l].quota;
    }
    return SFMB200_OK;
}

int sfmb200_orb_linear_exact_taps(int src, int dst, int32_t* i0, int32_t* i1, int32_t* weight) {
    if (src < 1 || dst < 1 || !i0 || !i1 || !weight) return SFMB200_ERR_INVALID;
    linear_exact_taps(src, dst, i0, i1, weight);
    return SFMB200_OK;
}

int sfmb200_orb_retain_best(const float* response, int n, int n_points, int32_t* order) {
    if (n < 0 || (n && (!response || !order))) return -1;
    std::vector<Rec> k(n);
    for (int i = 0; i < n; i++) k[i] = Rec{response[i], i};
    retain_best(k, n_points);
    for (size_t i = 0; i < k.size(); i++) order[i] = k[i].idx;
    return (int)k.size();
}

int sfmb200_orb_download_level(sfmb200_ctx* ctx, int stage, int image, int level, uint8_t* out) {
    if (!ctx || !out || level < 0 || level >= ORB_LEVELS) return SFMB200_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const OrbLast& o = ctx->orb_last;
    if (!o.pyr || image < 0 || image >= o.nimg) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "no extraction to inspect (image %d of %d)", image, o.nimg);
    OrbLayout L; make_layout(o.w, o.h, o.nfeatures, L);
    const uint8_t* base = stage == 0 ? o.pyr : stage == 1 ? o.blur : stage == 2 ? o.score : nullptr;
    if (!base) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "stage must be 0 (pyramid), 1 (blurred) or 2 (FAST score)");
    SFM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (L.lv[level].w == 0) return SFMB200_OK;
    SFM_CUDA(ctx, cudaMemcpyAsync(out, base + (size_t)image * o.slab + L.lv[level].off, (size_t)L.lv[level].w * L.lv[level].h, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFMB200_OK;
}

}  // extern "C"
