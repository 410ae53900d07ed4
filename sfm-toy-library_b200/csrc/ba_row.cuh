// ba_row.cuh -- K3b: the camera-row kernel of the bundle adjustment ("row" mode, default), its combine kernel and the stable
// camera-major sort.  Included by ba.cu inside its anonymous namespace (uses BAView, LMX, ObsJ, eval_scaled, dmma_m8n8k4).
//
// What it replaces: the camera-major kernel (diagonal blocks, camera-focal column, rhs, gradient, J^T J diagonal: 47 register
// accumulators per thread, 252 registers) AND the per-camera-pair gather kernel (one index pair + two scattered 144-byte
// records per off-diagonal update, 45 MB of entry lists built per problem).  Round-1 profile: 95 + 221 us of a 780 us step,
// pair kernel L1-throughput bound, 1.6 GB of record gathers per pass.
//
// Formulation.  The camera-major copy of the observation list is STABLE (ascending point inside a camera), and inside a point
// the observations are stored in ascending camera order (std::map order, reference SfMBundleAdjustmentUtils.cpp:146).  Hence
// for an observation o of camera ci the observations of the same point by cameras cj > ci are exactly the records that FOLLOW
// it in the point-major Z buffer: Z[o+1 .. o+npart].  A CTA walks a contiguous slice of the camera-major list; for the camera
// it is in it keeps the block row S[ci, ci+1 ..] (36 doubles per block, 28.8 KB at 100 cameras) in shared memory and adds
// Z_i Z_j^T for every follower j -- one mma.sync.m8n8k4.f64 per (i, j), operand b a coalesced 144-byte read of a record that
// is contiguous with its neighbours.  Per point with k observations that is k(k+1)/2 record reads (36 at k = 8) instead of
// k(k-1) (56) plus 8 bytes of index per pair, and no index lists at all.
//   * Block (ci,cj) is owned by warp (cj mod 4) of the CTA: plain read-modify-write, no atomics, fixed order.
//   * The diagonal terms of a batch of 128 observations are two 8x8 fp64 tensor-core products per warp:
//       X^T X   with X = [Jc | Jf | r]  (2 rows per observation)      -> Jc^T Jc, Jc^T Jf, Jc^T r, Jf^T Jf, Jf^T r, diag
//       Zh Zh^T with Zh = [Z ; zf^T ; zg^T] (8 x 3 per observation)    -> Z Z^T, Z zf, Z zg
//     i.e. 4 accumulator registers instead of 47.
//   * Every (slice, camera) segment writes ONE partial record; ba_combine_kernel sums the records of a camera in slice order.
//     No floating-point atomics anywhere: two runs give bitwise identical results.
#pragma once

constexpr int ROW_THREADS = 128;
constexpr int ROW_WARPS = ROW_THREADS / 32;
constexpr int ROW_HDR = 64;                               // header doubles of a partial record (62 used)
constexpr int ROW_ZSTRIDE = 3 * ROW_THREADS + 4;          // Zh row stride in doubles: = 4 (mod 32) -> conflict-free fragment loads
constexpr int ROW_PLIST = 256;                            // per-warp list of owned (i, j) updates of a batch
constexpr int ROW_UNROLL = 8;                             // updates in flight per warp

struct RowArgs {
    const int32_t* cm_obs;        // [nobs] point-major index o of each camera-major entry
    const uint8_t* cm_np;         // [nobs] number of followers of o inside its point (observations by cameras > ci)
    int per_cta;                  // camera-major entries per CTA slice
    double* part;                 // [(grid + nc)] partial records of rec_stride doubles; record of (slice s, camera c) = s + c
    int rec_stride;               // ROW_HDR + 36 * nc
};

static inline size_t row_smem_bytes(int nc) {
    return sizeof(double) * ((size_t)36 * nc + (size_t)ROW_THREADS * 2 * 8 + (size_t)8 * ROW_ZSTRIDE + (size_t)ROW_WARPS * 2 * 64) +
           sizeof(int) * ((size_t)ROW_THREADS * 2 + (size_t)ROW_WARPS * ROW_PLIST * 2);
}

template <bool NORM_ONLY>
__global__ void __launch_bounds__(ROW_THREADS) ba_row_kernel(BAView v, RowArgs ra) {
    const LMX x = lm_x(v);
    if (!x.run) return;
    const int start = blockIdx.x * ra.per_cta, stop = min(v.nobs, start + ra.per_cta);
    if (start >= stop) return;
    extern __shared__ __align__(16) unsigned char row_smem[];
    double* rowS = reinterpret_cast<double*>(row_smem);                     // [nc][36]: sum Z_i Z_j^T of blocks (c, cj)
    double* Xs = rowS + (size_t)36 * v.nc;                                  // [2*ROW_THREADS][8]
    double* Zh = Xs + (size_t)ROW_THREADS * 2 * 8;                          // [8][ROW_ZSTRIDE]
    double* fragred = Zh + (size_t)8 * ROW_ZSTRIDE;                         // [ROW_WARPS][2][64]
    int* meta_o = reinterpret_cast<int*>(fragred + ROW_WARPS * 2 * 64);     // [ROW_THREADS]
    int* meta_np = meta_o + ROW_THREADS;                                    // [ROW_THREADS]
    int2* plist = reinterpret_cast<int2*>(meta_np + ROW_THREADS);           // [ROW_WARPS][ROW_PLIST]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const bool zvalid = fr < 6 && fc < 3;                                   // this lane holds an element of a 6x3 Z record
    int c = 0;
    {   // last camera whose list begins at or before `start`
        int lo = 0, hi = v.nc - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (v.cm_off[mid] <= start) lo = mid; else hi = mid - 1; }
        c = lo;
    }
    const double f = *x.focal, sf = NORM_ONLY ? 1.0 : v.scale_cf[6 * v.nc];
    for (; c < v.nc && v.cm_off[c] < stop; ++c) {
        const int begin = max(start, v.cm_off[c]), end = min(stop, v.cm_off[c + 1]);
        if (begin >= end) continue;
        const CamDerived d = x.camd[c];
        double sc[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sc[a] = NORM_ONLY ? 1.0 : v.scale_cf[6 * c + a];
        if (!NORM_ONLY) for (int i = 36 * (c + 1) + tid; i < 36 * v.nc; i += ROW_THREADS) rowS[i] = 0.0;
        double dX0 = 0.0, dX1 = 0.0, dZ0 = 0.0, dZ1 = 0.0;                   // this warp's X^T X and Zh Zh^T accumulator fragments
        for (int b0 = begin; b0 < end; b0 += ROW_THREADS) {
            __syncthreads();                                               // previous batch fully consumed (and rowS zeroed)
            // ---- step 1: one observation per thread: Jacobian blocks, Z_i, operands into shared memory
            const int i = b0 + tid;
            const bool valid = i < end;
            int o = 0, npart = 0;
            double xr0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xr1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double Z[18], zfg[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < 18; ++q) Z[q] = 0.0;
            if (valid) {
                const int p = v.cm_pt[i];
                o = ra.cm_obs[i]; npart = ra.cm_np[i];
                const double X[3] = {x.pts[3 * p], x.pts[3 * p + 1], x.pts[3 * p + 2]};
                ObsJ J;
                if (NORM_ONLY) {
                    const double one[3] = {1.0, 1.0, 1.0};
                    eval_scaled(d, X, f, v.cm_xy[i], sc, one, 1.0, J);
                } else {
                    const double sp[3] = {v.scale_pt[3 * p], v.scale_pt[3 * p + 1], v.scale_pt[3 * p + 2]};
                    eval_scaled(d, X, f, v.cm_xy[i], sc, sp, sf, J);
                    const double* pb = v.ptblk + (size_t)p * PTB;
                    const double M[6] = {pb[0], pb[1], pb[2], pb[3], pb[4], pb[5]};
#pragma unroll
                    for (int q = 0; q < 6; ++q) zfg[q] = pb[6 + q];          // zg (3), zf (3)
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double w0 = J.Jc[a] * J.Jp[0] + J.Jc[6 + a] * J.Jp[3], w1 = J.Jc[a] * J.Jp[1] + J.Jc[6 + a] * J.Jp[4],
                                     w2 = J.Jc[a] * J.Jp[2] + J.Jc[6 + a] * J.Jp[5];
                        Z[a * 3] = w0 * M[0]; Z[a * 3 + 1] = w0 * M[1] + w1 * M[2]; Z[a * 3 + 2] = w0 * M[3] + w1 * M[4] + w2 * M[5];
                    }
                }
#pragma unroll
                for (int a = 0; a < 6; ++a) { xr0[a] = J.Jc[a]; xr1[a] = J.Jc[6 + a]; }
                xr0[6] = J.Jf[0]; xr1[6] = J.Jf[1]; xr0[7] = J.r[0]; xr1[7] = J.r[1];
            }
            {
                double2* xd = reinterpret_cast<double2*>(Xs + (size_t)tid * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) { xd[q] = make_double2(xr0[2 * q], xr0[2 * q + 1]); xd[4 + q] = make_double2(xr1[2 * q], xr1[2 * q + 1]); }
            }
            if (!NORM_ONLY) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) Zh[a * ROW_ZSTRIDE + 3 * tid + b] = Z[a * 3 + b];
#pragma unroll
                for (int b = 0; b < 3; ++b) { Zh[6 * ROW_ZSTRIDE + 3 * tid + b] = zfg[3 + b]; Zh[7 * ROW_ZSTRIDE + 3 * tid + b] = zfg[b]; }
                meta_o[tid] = o; meta_np[tid] = npart;
            }
            __syncthreads();
            // ---- step 2: diagonal terms of this warp's 32 observations on the FP64 tensor pipe.  A (m8 x k4, row) wants lane l
            // to hold A[l>>2][l&3], B (k4 x n8, col) B[l&3][l>>2]: for X^T X and Zh Zh^T both are the SAME value.
            {
                const double* xw = Xs + (size_t)warp * 64 * 8;
#pragma unroll
                for (int t = 0; t < 16; ++t) { const double xv = xw[(4 * t + fc) * 8 + fr]; dmma_m8n8k4(dX0, dX1, xv, xv); }
                if (!NORM_ONLY) {
                    const double* zw = Zh + (size_t)fr * ROW_ZSTRIDE + 96 * warp + fc;
#pragma unroll
                    for (int t = 0; t < 24; ++t) { const double zv = zw[4 * t]; dmma_m8n8k4(dZ0, dZ1, zv, zv); }
                }
            }
            if (NORM_ONLY) continue;
            // ---- step 3: off-diagonal row blocks.  Every warp scans the batch's 128 observations and collects the updates whose
            // partner camera it owns (cj mod 4 == warp) in (group, follower index, lane) order -- a fixed order; then processes
            // the list ROW_UNROLL at a time: operand loads first (a: Z_i from shared memory, b: Z_j from the point-major buffer,
            // 18 lanes x 8 bytes of one contiguous record), then the DMMAs, then the read-modify-writes of the owned blocks.
            int2* mylist = plist + warp * ROW_PLIST;
            int nlist = 0;
            auto flush = [&]() {
                for (int s = 0; s < nlist; s += ROW_UNROLL) {
                    double a[ROW_UNROLL], b[ROW_UNROLL];
                    int cjs[ROW_UNROLL];
#pragma unroll
                    for (int u = 0; u < ROW_UNROLL; ++u) {
                        const bool live = s + u < nlist;
                        const int2 en = mylist[live ? s + u : s];
                        const int e = en.y & 0xff;
                        cjs[u] = live ? (en.y >> 8) : -1;
                        const double va = zvalid ? Zh[fr * ROW_ZSTRIDE + 3 * e + fc] : 0.0;
                        const double vb = zvalid ? __ldg(v.Zbuf + (size_t)en.x * 18 + fr * 3 + fc) : 0.0;
                        a[u] = live ? va : 0.0; b[u] = live ? vb : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < ROW_UNROLL; ++u) {
                        double c0 = 0.0, c1 = 0.0;
                        dmma_m8n8k4(c0, c1, a[u], b[u]);
                        if (cjs[u] >= 0 && fr < 6 && fc < 3) {             // accumulator fragment: row fr, columns 2*fc, 2*fc+1
                            double2* dst = reinterpret_cast<double2*>(rowS + (size_t)cjs[u] * 36 + fr * 6 + 2 * fc);
                            double2 cur = *dst; cur.x += c0; cur.y += c1; *dst = cur;
                        }
                    }
                }
                nlist = 0;
                __syncwarp();
            };
#pragma unroll 1
            for (int g = 0; g < ROW_WARPS; ++g) {
                const int oe = meta_o[32 * g + lane], ne = meta_np[32 * g + lane];
                const int maxq = (int)__reduce_max_sync(0xffffffffu, (unsigned)ne);
#pragma unroll 1
                for (int q = 0; q < maxq; ++q) {
                    const int cq = q < ne ? __ldg(v.obs_cam + oe + 1 + q) : -1;
                    const bool own = cq >= 0 && (cq & (ROW_WARPS - 1)) == warp;
                    const unsigned m = __ballot_sync(0xffffffffu, own);
                    if (m == 0u) continue;
                    if (nlist + 32 > ROW_PLIST) flush();
                    if (own) mylist[nlist + __popc(m & ((1u << lane) - 1u))] = make_int2(oe + 1 + q, (cq << 8) | (32 * g + lane));
                    nlist += __popc(m);
                    __syncwarp();
                }
            }
            flush();
        }
        // ---- end of this (slice, camera) segment: combine the warps' diagonal fragments in warp order, write the partial record
        __syncthreads();
        {
            double* fw = fragred + (size_t)warp * 128;
            // accumulator fragment: row fr, columns 2*fc and 2*fc+1
            fw[fr * 8 + 2 * fc] = dX0; fw[fr * 8 + 2 * fc + 1] = dX1;
            fw[64 + fr * 8 + 2 * fc] = dZ0; fw[64 + fr * 8 + 2 * fc + 1] = dZ1;
        }
        __syncthreads();
        double* rec = ra.part + (size_t)(blockIdx.x + c) * ra.rec_stride;
        if (tid < 64) {
            double xx = 0.0, zz = 0.0;
#pragma unroll
            for (int w = 0; w < ROW_WARPS; ++w) { xx += fragred[w * 128 + tid]; zz += fragred[w * 128 + 64 + tid]; }
            const int r = tid >> 3, q = tid & 7;
            if (r < 6 && q < 6) rec[r * 6 + q] = xx - zz;                    // diagonal block  Jc^T Jc - Z Z^T
            else if (r < 6 && q == 6) rec[36 + r] = xx - zz;                // camera-focal    Jc^T Jf - Z zf
            else if (r < 6 && q == 7) { rec[42 + r] = xx - zz; rec[48 + r] = xx; }   // rhs  Jc^T r - Z zg ; gradient Jc^T r
            else if (r == 6 && q == 6) rec[60] = xx;                        // Jf^T Jf
            else if (r == 6 && q == 7) rec[61] = xx;                        // Jf^T r
            if (r < 6 && q == r) rec[54 + r] = xx;                          // diag(Jc^T Jc)
        }
        if (!NORM_ONLY) for (int i = 36 * (c + 1) + tid; i < 36 * v.nc; i += ROW_THREADS) rec[ROW_HDR + i] = rowS[i];
    }
}

// Sums the partial records of every camera in slice order and writes the reduced system in the layout the rest of the solver
// (rank exchange, ba_assemble_kernel, ba_cam_update_kernel) reads: Sblk | Scf | Sff | rhs | gcf | dcf.  CTA c = camera c.
// The focal terms need a sum over cameras: the last CTA to finish adds them up in camera order (fixed order whoever is last).
// norm_only: only the squared column norms of the unscaled Jacobian (Jacobi scaling at x0) -> colnorm [6 nc + 1].
__global__ void __launch_bounds__(256) ba_combine_kernel(BAView v, RowArgs ra, int norm_only, double* __restrict__ colnorm,
                                                         double* __restrict__ fpart /* [nc][2] */, unsigned* __restrict__ counter) {
    if (v.st && v.st->status != LM_RUNNING) return;
    const int c = blockIdx.x, nc = v.nc;
    int first = 0, last = -1;
    if (v.cm_off[c + 1] > v.cm_off[c]) { first = v.cm_off[c] / ra.per_cta; last = (v.cm_off[c + 1] - 1) / ra.per_cta; }
    auto sum = [&](int idx) {
        double s = 0.0;
        for (int sl = first; sl <= last; ++sl) s += ra.part[(size_t)(sl + c) * ra.rec_stride + idx];
        return s;
    };
    if (norm_only) {
        if (threadIdx.x < 6) colnorm[6 * c + threadIdx.x] = sum(54 + threadIdx.x);
        if (threadIdx.x == 6) fpart[2 * c] = sum(60);
    } else {
        for (int t = threadIdx.x; t < 62; t += blockDim.x) {
            const double s = sum(t);
            if (t < 36) v.Sblk[blk_index(c, c, nc) * 36 + t] = s;
            else if (t < 42) v.Scf[6 * c + (t - 36)] = s;
            else if (t < 48) v.rhs[6 * c + (t - 42)] = s;
            else if (t < 54) v.gcf[6 * c + (t - 48)] = s;
            else if (t < 60) v.dcf[6 * c + (t - 54)] = s;
            else fpart[2 * c + (t - 60)] = s;
        }
        for (int i = 36 * (c + 1) + threadIdx.x; i < 36 * nc; i += blockDim.x) {
            const int cj = i / 36, ab = i - 36 * cj;
            v.Sblk[blk_index(c, cj, nc) * 36 + ab] = -sum(ROW_HDR + i);
        }
    }
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;    // wraps to 0: ready for the next launch
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        double ff = 0.0, gf = 0.0;
        for (int k = 0; k < nc; ++k) { ff += __ldcg(fpart + 2 * k); gf += __ldcg(fpart + 2 * k + 1); }
        const int fidx = 6 * nc;
        if (norm_only) colnorm[fidx] = ff;
        else {
            // the point kernel's last CTA stored the point part of S_ff and rhs_f there (stream order: it ran before)
            *v.Sff += ff; v.dcf[fidx] = ff; v.rhs[fidx] += gf; v.gcf[fidx] = gf;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Stable camera-major copy of the observation list (structure fixed across LM iterations, built once per problem): a counting
// sort by camera in which the rank of an observation among those of its camera is computed WITHOUT arrival-order atomics
// (__match_any_sync inside a warp, a prefix over the warps of a CTA, a scan over the CTAs) -- so each camera's list is in
// ascending point order, the same on every run.  SCATTER = false: per-CTA histograms only.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CMS_THREADS = 256;
constexpr int CMS_WARPS = CMS_THREADS / 32;
template <bool SCATTER>
__global__ void __launch_bounds__(CMS_THREADS) cm_sort_kernel(const int32_t* __restrict__ obs_cam, const float2* __restrict__ obs_xy,
                                                              const int32_t* __restrict__ obs_pt, const int32_t* __restrict__ pt_off, int nobs, int nc,
                                                              int* __restrict__ hist /* [ncta][nc]: counts (in), exclusive prefix over CTAs (SCATTER) */,
                                                              const int32_t* __restrict__ cm_off, float2* __restrict__ cm_xy, int32_t* __restrict__ cm_pt,
                                                              int32_t* __restrict__ cm_obs, uint8_t* __restrict__ cm_np) {
    extern __shared__ int wcnt[];                                           // [CMS_WARPS][nc]
    const int o = blockIdx.x * CMS_THREADS + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < CMS_WARPS * nc; i += CMS_THREADS) wcnt[i] = 0;
    __syncthreads();
    const bool valid = o < nobs;
    const int c = valid ? obs_cam[o] : -1 - lane;                           // invalid lanes get distinct keys
    const unsigned m = __match_any_sync(0xffffffffu, c);
    const int rank = __popc(m & ((1u << lane) - 1u));
    if (valid && rank == 0) wcnt[warp * nc + c] = __popc(m);
    __syncthreads();
    for (int k = threadIdx.x; k < nc; k += CMS_THREADS) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < CMS_WARPS; ++w) { const int t = wcnt[w * nc + k]; wcnt[w * nc + k] = run; run += t; }
        if (!SCATTER) hist[(size_t)blockIdx.x * nc + k] = run;
    }
    if (!SCATTER) return;
    __syncthreads();
    if (valid) {
        const int pos = cm_off[c] + hist[(size_t)blockIdx.x * nc + c] + wcnt[warp * nc + c] + rank;
        const int p = obs_pt[o];
        cm_xy[pos] = obs_xy[o]; cm_pt[pos] = p; cm_obs[pos] = o; cm_np[pos] = (uint8_t)(pt_off[p + 1] - o - 1);
    }
}
// per camera: exclusive scan of hist[.][c] over the CTAs (in place) and the camera's total.  One CTA per camera.
__global__ void __launch_bounds__(1024) cm_scan_ctas_kernel(int* __restrict__ hist, int ncta, int nc, int* __restrict__ total) {
    __shared__ int wsum[32], carry_s;
    const int c = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ncta; base += 1024) {
        const int i = base + threadIdx.x;
        const int val = i < ncta ? hist[(size_t)i * nc + c] : 0;
        int s = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += a; }
        if (lane == 31) wsum[w] = s;
        __syncthreads();
        if (w == 0) {
            int a = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, a, o); if (lane >= o) a += y; }
            wsum[lane] = a;
        }
        __syncthreads();
        if (i < ncta) hist[(size_t)i * nc + c] = carry_s + (w ? wsum[w - 1] : 0) + s - val;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) total[c] = carry_s;
}
