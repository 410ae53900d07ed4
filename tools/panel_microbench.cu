// tools/panel_microbench.cu -- phase-level cycle counts of chol_panel_kernel (generated from csrc/ba.cu by tools/make_panel_microbench.py)
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
constexpr int NB = 32;
constexpr int PANEL_WARPS = 4;
__global__ void __launch_bounds__(PANEL_WARPS * 32) chol_panel_kernel(double* __restrict__ A, int npad, int n, int k, int nbk,
                                                                      double* __restrict__ dinv, int* __restrict__ fail, long long* __restrict__ clk) {
    const long long t0 = clock64();
    __shared__ double Ls[NB][NB + 1];          // diagonal tile, then its factor (lower)
    __shared__ double colbuf[2][NB];
    __shared__ double invd[NB];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = k + 1 + blockIdx.x * PANEL_WARPS + w;
    const bool has_tile = i < nbk;
    double b[NB];                               // this warp's sub-diagonal tile, row `lane`
    if (has_tile) {
        const double* src = A + (size_t)(i * NB + lane) * npad + k * NB;
#pragma unroll
        for (int c = 0; c < NB; c += 2) { const double2 v = *reinterpret_cast<const double2*>(src + c); b[c] = v.x; b[c + 1] = v.y; }
    }
    {   // diagonal tile -> shared memory; all 8 loads of a thread are issued before the first store
        double v[NB * NB / (PANEL_WARPS * 32)];
#pragma unroll
        for (int u = 0; u < NB * NB / (PANEL_WARPS * 32); ++u) { const int e = threadIdx.x + u * PANEL_WARPS * 32; v[u] = A[(size_t)(k * NB + (e >> 5)) * npad + k * NB + (e & 31)]; }
#pragma unroll
        for (int u = 0; u < NB * NB / (PANEL_WARPS * 32); ++u) { const int e = threadIdx.x + u * PANEL_WARPS * 32; Ls[e >> 5][e & 31] = v[u]; }
    }
    __syncthreads();
    const long long t1 = clock64();
    // ---- phase 1: col[q] = column 4*(q + rot) + w of row `lane`  (rot = number of rotations so far)
    double col[NB / PANEL_WARPS];
#pragma unroll
    for (int q = 0; q < NB / PANEL_WARPS; ++q) col[q] = Ls[lane][PANEL_WARPS * q + w];
    bool bad = false;
#pragma unroll 1
    for (int jb = 0; jb < NB; jb += PANEL_WARPS) {
#pragma unroll
        for (int ow = 0; ow < PANEL_WARPS; ++ow) {                    // pivot j = jb + ow is column col[0] of warp ow
            const int j = jb + ow, gj = k * NB + j;
            if (w == ow) {
                const double d = __shfl_sync(0xffffffffu, col[0], j);
                double ljj, inv;
                if (gj >= n) { ljj = 1.0; inv = 0.0; }
                else if (!(d > 0.0) || !isfinite(d)) { bad = true; ljj = 1.0; inv = 1.0; }
                else { inv = rsqrt(d); ljj = d * inv; }
                const double lrj = lane == j ? ljj : (lane > j ? col[0] * inv : 0.0);
                col[0] = lrj;
                colbuf[j & 1][lane] = lrj;
                if (lane == j) invd[j] = inv;
            }
            __syncthreads();
            const double lrj = colbuf[j & 1][lane];
#pragma unroll
            for (int q = 0; q < NB / PANEL_WARPS; ++q) {
                const int c = jb + PANEL_WARPS * q + w;               // column held in col[q]; >= NB means wrapped (finished)
                if (c > j && c < NB) col[q] = fma(-lrj, colbuf[j & 1][c], col[q]);
            }
        }
        // the pivot columns of this group are final: store them, rotate the register set by one
        if (lane >= jb + w) Ls[lane][jb + w] = col[0]; else Ls[lane][jb + w] = 0.0;
        const double t = col[0];
#pragma unroll
        for (int q = 0; q < NB / PANEL_WARPS - 1; ++q) col[q] = col[q + 1];
        col[NB / PANEL_WARPS - 1] = t;
    }
    const long long t2 = clock64();
    if (bad && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(fail, 1);
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < NB * NB; e += PANEL_WARPS * 32) { const int r = e >> 5, c = e & 31; if (c <= r) A[(size_t)(k * NB + r) * npad + k * NB + c] = Ls[r][c]; }
        if (threadIdx.x < NB) dinv[k * NB + threadIdx.x] = invd[threadIdx.x];
    }
    const long long t3 = clock64();
    if (!has_tile) return;
    // ---- phase 2: X L^T = B for row `lane`, columns in registers.  The register row is rotated by PANEL_WARPS every
    // PANEL_WARPS pivots so that the loop stays rolled with static indices; L[c][j] arrives as a broadcast LDS.
#pragma unroll 1
    for (int jb = 0; jb < NB; jb += PANEL_WARPS) {
#pragma unroll
        for (int u = 0; u < PANEL_WARPS; ++u) {
            const int j = jb + u;
            const double xj = b[u] * invd[j];
            b[u] = xj;
#pragma unroll
            for (int p2 = u + 1; p2 < NB; ++p2)
                if (jb + p2 < NB) b[p2] = fma(-xj, Ls[jb + p2][j], b[p2]);
        }
        double t[PANEL_WARPS];
#pragma unroll
        for (int u = 0; u < PANEL_WARPS; ++u) t[u] = b[u];
#pragma unroll
        for (int p2 = 0; p2 < NB - PANEL_WARPS; ++p2) b[p2] = b[p2 + PANEL_WARPS];
#pragma unroll
        for (int u = 0; u < PANEL_WARPS; ++u) b[NB - PANEL_WARPS + u] = t[u];
    }
    const long long t4 = clock64();
    double* dst = A + (size_t)(i * NB + lane) * npad + k * NB;
#pragma unroll
    for (int c = 0; c < NB; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(b[c], b[c + 1]);
    const long long t5 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t5 - t4; }
}


int main() {
    const int n = 601, npad = 608, nbk = 19;
    std::vector<double> h((size_t)npad * npad, 0.0);
    for (int i = 0; i < npad; ++i) for (int j = 0; j <= i; ++j) h[(size_t)i * npad + j] = (i == j) ? 700.0 + i : 1.0 / (1 + i - j);
    double *A, *dinv; int* fail; long long* clk;
    cudaMalloc(&A, h.size() * 8); cudaMalloc(&dinv, npad * 8); cudaMalloc(&fail, 8); cudaMalloc(&clk, 64);
    cudaMemset(fail, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        chol_panel_kernel<<<5, 128>>>(A, npad, n, 0, nbk, dinv, fail, clk);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c[5]; cudaMemcpy(c, clk, 40, cudaMemcpyDeviceToHost);
        printf("rep %d: %.1f us  cycles: load %lld  phase1 %lld  writeback %lld  solve %lld  store %lld  (%s)\n", rep, ms * 1e3, c[0], c[1], c[2], c[3], c[4], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
