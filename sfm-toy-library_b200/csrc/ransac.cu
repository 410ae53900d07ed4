// ransac.cu -- SURVEY.md 8 row f-2: batched hypothesis scoring for the three RANSAC stages that sit on either side of
// triangulation in the reference driver:
//   findHomographyInliers        SfMToyLib/SfMStereoUtilities.cpp:51-72   cv::findHomography(RANSAC, 10 px)   -> inlier count
//   findCameraMatricesFromMatch  SfMToyLib/SfMStereoUtilities.cpp:74-118  cv::findEssentialMat(RANSAC, 0.999, 1 px) + recoverPose
//   findCameraPoseFrom2D3DMatch  SfMToyLib/SfMStereoUtilities.cpp:208-243 cv::solvePnPRansac(100 iterations, 10 px, 0.99)
// OpenCV's RANSAC evaluates one hypothesis at a time over all correspondences (RANSACPointSetRegistrator::run -> computeError
// -> findInliers); here ALL hypotheses of a run are scored against ALL correspondences in one launch: grid (point chunks,
// hypotheses), a block counts the inliers of its chunk, integer atomics collect the counts (deterministic), a second kernel
// picks the best hypothesis (most inliers, ties -> lowest index = the first one OpenCV would have kept) and writes its mask.
// The hypotheses themselves come from the host (minimal solvers); the error formulas and their arithmetic types follow
// OpenCV's computeError callbacks so that a hypothesis gets the inlier set cv:: gives it:
//   homography  fundam.cpp  HomographyEstimatorCallback::computeError: FLOAT  ww = 1/(h6 x + h7 y + 1), dx, dy, err = dx^2 + dy^2
//   essential   five-point.cpp EMEstimatorCallback::computeError: DOUBLE Sampson error (x2^T E x1)^2 / (|Ex1|_xy^2 + |E^T x2|_xy^2), stored as float
//   pose        solvepnp.cpp PnPRansacCallback::computeError: projection (double), float difference, squared norm
// inlier  <=>  err <= (float)(threshold^2)   (RANSACPointSetRegistrator::findInliers).
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;

struct Model { double m[12]; };      // 3x3 (homography / essential) or 3x4 pose, row-major

__device__ __forceinline__ float err_homography(const double* H, float x, float y, float X, float Y) {
    // no FMA contraction: the CPU code is plain float mul/add in this order
    const float h0 = (float)H[0], h1 = (float)H[1], h2 = (float)H[2], h3 = (float)H[3], h4 = (float)H[4], h5 = (float)H[5], h6 = (float)H[6], h7 = (float)H[7];
    const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(h6, x), __fmul_rn(h7, y)), 1.f));
    const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(h0, x), __fmul_rn(h1, y)), h2), ww), X);
    const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(h3, x), __fmul_rn(h4, y)), h5), ww), Y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ float err_essential(const double* E, double x1, double y1, double x2, double y2) {
    // Ex1 = E * (x1, y1, 1), Etx2 = E^T * (x2, y2, 1)
    const double e0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x1), __dmul_rn(E[1], y1)), E[2]);
    const double e1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], x1), __dmul_rn(E[4], y1)), E[5]);
    const double e2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], x1), __dmul_rn(E[7], y1)), E[8]);
    const double t0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x2), __dmul_rn(E[3], y2)), E[6]);
    const double t1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], x2), __dmul_rn(E[4], y2)), E[7]);
    const double x2tEx1 = __dadd_rn(__dadd_rn(__dmul_rn(x2, e0), __dmul_rn(y2, e1)), e2);
    const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(t0, t0)), __dmul_rn(t1, t1));
    return (float)(__dmul_rn(x2tEx1, x2tEx1) / den);
}
__device__ __forceinline__ float err_pose(const double* P, const double* K, float X, float Y, float Z, float u, float v) {
    const double x = P[0] * X + P[1] * Y + P[2] * Z + P[3], y = P[4] * X + P[5] * Y + P[6] * Z + P[7], z = P[8] * X + P[9] * Y + P[10] * Z + P[11];
    const double iz = z != 0.0 ? 1.0 / z : 1.0;          // cv::projectPoints: z = z ? 1./z : 1
    const float pu = (float)(K[0] * (x * iz) + K[2]), pv = (float)(K[4] * (y * iz) + K[5]);
    const float dx = __fsub_rn(u, pu), dy = __fsub_rn(v, pv);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// model: 0 homography, 1 essential (a, b already normalised doubles? no: float pixels, normalised here with f, cx, cy in aux), 2 pose
template <int MODEL>
__device__ __forceinline__ bool is_inlier(const double* M, const double* aux, const float* __restrict__ a, const float* __restrict__ b, int i, float t2) {
    if (MODEL == 0) return err_homography(M, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]) <= t2;
    if (MODEL == 1) {
        // cv::findEssentialMat(points, focal, pp): points converted to double and normalised (x - cx) / f before the estimator
        const double f = aux[0], cx = aux[1], cy = aux[2];
        return err_essential(M, ((double)a[2 * i] - cx) / f, ((double)a[2 * i + 1] - cy) / f, ((double)b[2 * i] - cx) / f, ((double)b[2 * i + 1] - cy) / f) <= t2;
    }
    return err_pose(M, aux, a[3 * i], a[3 * i + 1], a[3 * i + 2], b[2 * i], b[2 * i + 1]) <= t2;
}

template <int MODEL>
__global__ void __launch_bounds__(RS_THREADS) ransac_score_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, const Model* __restrict__ hyp,
                                                                  const double* __restrict__ aux, float t2, int per_block, int32_t* __restrict__ counts) {
    __shared__ double M[12], A[9];
    __shared__ int wsum[RS_THREADS / 32];
    if (threadIdx.x < 12) M[threadIdx.x] = hyp[blockIdx.y].m[threadIdx.x];
    if (threadIdx.x < 9) A[threadIdx.x] = aux[threadIdx.x];
    __syncthreads();
    const int begin = blockIdx.x * per_block, end = min(n, begin + per_block);
    int c = 0;
    for (int i = begin + threadIdx.x; i < end; i += RS_THREADS) c += is_inlier<MODEL>(M, A, a, b, i, t2) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < RS_THREADS / 32; ++w) s += wsum[w];
        if (s) atomicAdd(counts + blockIdx.y, s);
    }
}

// best hypothesis = most inliers, ties -> lowest index; single CTA
__global__ void __launch_bounds__(1024) ransac_best_kernel(const int32_t* __restrict__ counts, int nh, int32_t* __restrict__ best) {
    __shared__ long long red[32];
    long long key = -1;                                   // (count << 32) | (0x7fffffff - index): max = most inliers, lowest index
    for (int h = threadIdx.x; h < nh; h += 1024) { const long long k = ((long long)counts[h] << 32) | (long long)(0x7fffffff - h); key = k > key ? k : key; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const long long other = __shfl_xor_sync(0xffffffffu, key, o); key = other > key ? other : key; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) key = red[w] > key ? red[w] : key;
        best[0] = nh > 0 ? 0x7fffffff - (int)(key & 0xffffffffLL) : -1;
        best[1] = nh > 0 ? (int)(key >> 32) : 0;
    }
}
template <int MODEL>
__global__ void __launch_bounds__(RS_THREADS) ransac_mask_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, const Model* __restrict__ hyp,
                                                                 const double* __restrict__ aux, float t2, const int32_t* __restrict__ best, uint8_t* __restrict__ mask) {
    const int h = best[0];
    const int i = blockIdx.x * RS_THREADS + threadIdx.x;
    if (i >= n) return;
    if (h < 0) { mask[i] = 0; return; }
    double M[12], A[9];
#pragma unroll
    for (int k = 0; k < 12; ++k) M[k] = hyp[h].m[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = aux[k];
    mask[i] = is_inlier<MODEL>(M, A, a, b, i, t2) ? 1 : 0;
}

}  // namespace

extern "C" int sfmb200_ransac_score(sfmb200_ctx* ctx, int model, const float* a, const float* b, int n, const double* hyp, int nh, const double* aux9,
                                    double threshold, int32_t* inlier_counts, int32_t* best_index, uint8_t* best_mask) {
    if (!ctx || model < 0 || model > 2 || n < 0 || nh < 0) return SFMB200_ERR_INVALID;
    if (best_index) *best_index = -1;
    if (n == 0 || nh == 0) { if (inlier_counts) for (int h = 0; h < nh; ++h) inlier_counts[h] = 0; if (best_mask) memset(best_mask, 0, (size_t)n); return SFMB200_OK; }
    if (!a || !b || !hyp) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "null buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    SFM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int adim = model == 2 ? 3 : 2, hdim = model == 2 ? 12 : 9;
    std::vector<Model> hm(nh);
    for (int h = 0; h < nh; ++h) {
        memset(hm[h].m, 0, sizeof hm[h].m);
        for (int k = 0; k < hdim; ++k) hm[h].m[k] = hyp[(size_t)h * hdim + k];
        if (model == 0) {                                   // cv:: keeps homographies normalised to h22 = 1 (the error formula assumes it)
            const double s = hm[h].m[8];
            if (s != 0.0 && s != 1.0) for (int k = 0; k < 9; ++k) hm[h].m[k] /= s;
        }
    }
    double aux[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (aux9) memcpy(aux, aux9, sizeof aux);
    const size_t bytes = Carver::pad(4 * (size_t)n * adim) + Carver::pad(8 * (size_t)n) + Carver::pad(sizeof(Model) * nh) + Carver::pad(72) + Carver::pad(4 * (size_t)nh) + Carver::pad(n) + 1024;
    SFM_CUDA(ctx, ctx->scratch.reserve(bytes));
    Carver cv(ctx->scratch.p);
    float* d_a = cv.take<float>((size_t)n * adim); float* d_b = cv.take<float>((size_t)n * 2); Model* d_h = cv.take<Model>(nh); double* d_aux = cv.take<double>(9);
    int32_t* d_cnt = cv.take<int32_t>(nh); int32_t* d_best = cv.take<int32_t>(2); uint8_t* d_mask = cv.take<uint8_t>(n);
    SFM_CUDA(ctx, cudaMemcpyAsync(d_a, a, 4 * (size_t)n * adim, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_b, b, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_h, hm.data(), sizeof(Model) * nh, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(d_aux, aux, 72, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 4 * (size_t)nh, ctx->stream));
    const float t2 = (float)(threshold * threshold);
    // enough blocks to fill the machine: hypotheses x point chunks
    const int chunks = std::max(1, std::min(ceil_div(n, RS_THREADS), ceil_div(4 * ctx->sm_count, nh)));
    const int per_block = ceil_div(n, chunks);
    dim3 grid(ceil_div(n, per_block), nh);
    if (model == 0) ransac_score_kernel<0><<<grid, RS_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_h, d_aux, t2, per_block, d_cnt);
    else if (model == 1) ransac_score_kernel<1><<<grid, RS_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_h, d_aux, t2, per_block, d_cnt);
    else ransac_score_kernel<2><<<grid, RS_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_h, d_aux, t2, per_block, d_cnt);
    SFM_LAUNCH_CHECK(ctx);
    ransac_best_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt, nh, d_best);
    SFM_LAUNCH_CHECK(ctx);
    if (best_mask) {
        const int mb = ceil_div(n, RS_THREADS);
        if (model == 0) ransac_mask_kernel<0><<<mb, RS_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_h, d_aux, t2, d_best, d_mask);
        else if (model == 1) ransac_mask_kernel<1><<<mb, RS_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_h, d_aux, t2, d_best, d_mask);
        else ransac_mask_kernel<2><<<mb, RS_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_h, d_aux, t2, d_best, d_mask);
        SFM_LAUNCH_CHECK(ctx);
        SFM_CUDA(ctx, cudaMemcpyAsync(best_mask, d_mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    int32_t hb[2] = {-1, 0};
    if (inlier_counts) SFM_CUDA(ctx, cudaMemcpyAsync(inlier_counts, d_cnt, 4 * (size_t)nh, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaMemcpyAsync(hb, d_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (best_index) *best_index = hb[0];
    return SFMB200_OK;
}
