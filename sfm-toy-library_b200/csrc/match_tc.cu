// match_tc.cu -- K1 on the 5th-generation tensor cores: Hamming knn2 as an exact integer GEMM (tcgen05 + TMEM).
//
// For 256-bit descriptors  ham(q,t) = popc(q) + popc(t) - 2 * <q,t>  with q,t in {0,1}^256: the
// inner products of a 128-query x 256-train tile are ONE accumulator tile of  tcgen05.mma.kind::i8  (u8 x u8 -> s32,
// exact: every sum <= 256), M=128 N=256 K=32 per instruction, 8 instructions per tile, accumulators in TMEM.
//   * operands: bits are mapped to SIGNED bytes +1/-1, so that  <a,b> = 256 - 2*hamming  and the epilogue needs one compare
//     per element and no popcount terms.  The expansion happens ONCE per descriptor set (expand_blocks_kernel) into 64 KB
//     blocks of 256 rows laid out in the UMMA K-major, no-swizzle ("interleaved") canonical form: core matrix = 8 rows x
//     16 bytes, LBO = distance between the two 16-byte K chunks of one instruction, SBO = distance between 8-row groups
//     (cute/arch/mma_sm100_desc.hpp, make_umma_desc<Major::K>).  A tile then travels HBM -> shared memory as plain
//     cp.async.bulk copies completing on an mbarrier (no tensor map needed: a block is contiguous).
//   * warp-specialised pipeline: a loader thread (3 shared-memory stages of the train operand), an MMA thread (8
//     tcgen05.mma per tile, two 256-column accumulator stages = all 512 TMEM columns; tcgen05.commit releases the smem
//     stage and publishes the accumulator), and 4 epilogue warps (one per TMEM lane quarter, thread = query row) that pull
//     the accumulators with tcgen05.ld.32x32b.x32 and keep the running top-2 in registers with exactly the ordering of
//     the XOR/POPC kernel (strict '<', ascending train index) -> bit-identical results, same merge / ratio-test /
//     compaction epilogue (match.cu).
// Only descriptor width 32 bytes (ORB, the reference's case); other widths use the XOR/POPC kernel.
#include "common.cuh"
#include "match_common.cuh"

namespace {

constexpr int TC_M = 128;            // query rows per CTA  (UMMA M, TMEM lanes)
constexpr int TC_N = 256;            // train rows per tile (UMMA N, TMEM columns per accumulator stage) = rows per expanded block
constexpr int TC_K = 256;            // descriptor bits = K elements (one signed byte each after expansion)
constexpr int TC_BLOCK_BYTES = TC_N * TC_K;              // 64 KB: one 256-row block of expanded operands
constexpr int TC_A_BYTES = TC_M * TC_K;                  // 32 KB
constexpr int TC_BSTAGES = 3;                            // shared-memory stages of the train operand
constexpr int TC_EPI_WARPS = 8;                          // two per TMEM lane quarter: warps 0-3 take columns [0,128), warps 4-7 [128,256)
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;      // + warp 8: loader, warp 9: MMA issuer
constexpr int TC_SMEM = TC_A_BYTES + TC_BSTAGES * TC_BLOCK_BYTES + 256;
// instruction descriptor (UMMA::InstrDescriptor): c_format S32 (2) @bit4, a/b format signed INT8 (1) @bits 7/10, K-major both,
// n_dim = N>>3 @bit17, m_dim = M>>4 @bit24
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}

// 16 descriptor bits -> 16 signed bytes: bit 1 -> +1, bit 0 -> -1   (sum over k of a_k b_k = 256 - 2 * hamming)
__device__ __forceinline__ uint32_t pm1(uint32_t nib) {
    const uint32_t x = (nib * 0x00204081u) & 0x01010101u;
    return x | ((x ^ 0x01010101u) * 0xFFu);
}
__device__ __forceinline__ uint4 expand16(uint32_t bits16) {
    return make_uint4(pm1(bits16 & 0xF), pm1((bits16 >> 4) & 0xF), pm1((bits16 >> 8) & 0xF), pm1((bits16 >> 12) & 0xF));
}

// Expanded operand store (built once per descriptor set): per 256-row block 64 KB in the UMMA K-major no-swizzle canonical
// layout  [16-byte K chunk kc 0..15][row group r/8][r%8][16 B]  ->  LBO = 4096 B between K chunks, SBO = 128 B between
// 8-row groups.  A 128-row half of a block is the same layout at start offset +2048 B per chunk, so one store serves
// both the query (M=128) and the train (N=256) operand.  Rows beyond the image are zero (contribute nothing).
__global__ void __launch_bounds__(256) expand_blocks_kernel(const uint32_t* __restrict__ desc, const int2* __restrict__ blocks /* (first row, valid rows) */,
                                                            uint8_t* __restrict__ E) {
    const int2 b = blocks[blockIdx.x];
    uint8_t* out = E + (size_t)blockIdx.x * TC_BLOCK_BYTES;
    for (int it = threadIdx.x; it < TC_N * 8; it += 256) {
        const int w = it / TC_N, r = it - w * TC_N;
        const bool valid = r < b.y;
        const uint32_t v = valid ? __ldg(desc + (size_t)(b.x + r) * 8 + w) : 0u;
        const uint4 lo = valid ? expand16(v & 0xFFFFu) : make_uint4(0, 0, 0, 0), hi = valid ? expand16(v >> 16) : make_uint4(0, 0, 0, 0);
        const uint32_t off = (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
        *reinterpret_cast<uint4*>(out + (size_t)(2 * w) * (TC_N * 16) + off) = lo;
        *reinterpret_cast<uint4*>(out + (size_t)(2 * w + 1) * (TC_N * 16) + off) = hi;
    }
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory"); }
// bounded wait: returns false if the barrier never flips (a descriptor mistake must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}

// Warp-specialised: loader thread (cp.async.bulk of pre-expanded 64 KB blocks, 3 stages) -> MMA thread (8 x tcgen05.mma per
// tile into one of 2 TMEM accumulator stages, commits release the smem stage and publish the accumulator) -> 4 epilogue
// warps (tcgen05.ld, running top-2 per query row in registers).
__global__ void __launch_bounds__(TC_THREADS) knn2_hamming_tc_kernel(const uint8_t* __restrict__ E, const PairDesc* __restrict__ pairs, int qblocks,
                                                                     int splits, int4* __restrict__ partial, int* __restrict__ error_flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_A_BYTES + TC_BSTAGES * TC_BLOCK_BYTES);
    uint64_t* a_full = bars;                 // [1]
    uint64_t* full = bars + 1;               // [TC_BSTAGES] bytes of a B stage have landed
    uint64_t* smem_free = full + TC_BSTAGES; // [TC_BSTAGES] the MMAs that read the stage have completed
    uint64_t* acc_full = smem_free + TC_BSTAGES;   // [2] accumulator stage holds a finished tile
    uint64_t* acc_empty = acc_full + 2;      // [2] the 4 epilogue warps have drained the stage
    __shared__ uint32_t tmem_base_s;

    const PairDesc pd = pairs[blockIdx.y];
    const int qb = blockIdx.x / splits, sp = blockIdx.x % splits;
    if (qb * TC_M >= pd.nq) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_total = (pd.nt + TC_N - 1) / TC_N;
    const int tiles_per_split = (tiles_total + splits - 1) / splits;
    const int tile0 = sp * tiles_per_split, ntiles = max(0, min(tiles_total, tile0 + tiles_per_split) - tile0);
    const int q_row0 = qb * TC_M;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < TC_BSTAGES; ++i) { mbar_init(full + i, 1); mbar_init(smem_free + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_EPI_WARPS) {        // TMEM: two accumulator stages of 128 lanes x 256 columns (s32) = all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    bool ok = true;

    Top2 best = {INT_MAX, -1, INT_MAX, -1};
    // top-2 of the upper column half, merged by the lower-half warp at the end.  Lives in B stage 0: every bulk copy into a B
    // stage is consumed (waited on) before the last accumulator is published, whereas the query-tile copy into sA may still
    // be in flight when a split owns no train tile at all.
    int4* half_best = reinterpret_cast<int4*>(sB);
    if (warp == TC_EPI_WARPS) {
        if (lane == 0) {
            // ---- loader: query tile (a 128-row half of a block: 16 chunks of 2 KB), then the train blocks
            const uint8_t* qsrc = E + (size_t)(pd.q_blk + q_row0 / TC_N) * TC_BLOCK_BYTES + (size_t)((q_row0 / TC_M) & 1) * (TC_M * 16);
            mbar_expect_tx(a_full, TC_A_BYTES);
            for (int kc = 0; kc < 16; ++kc) bulk_g2s(sA + kc * (TC_M * 16), qsrc + (size_t)kc * (TC_N * 16), TC_M * 16, a_full);
            for (int t = 0; t < ntiles && ok; ++t) {
                const int s = t % TC_BSTAGES, u = t / TC_BSTAGES;
                if (u > 0 && !mbar_wait(smem_free + s, (u - 1) & 1)) { ok = false; break; }
                const uint8_t* src = E + (size_t)(pd.t_blk + tile0 + t) * TC_BLOCK_BYTES;
                mbar_expect_tx(full + s, TC_BLOCK_BYTES);
                for (int c = 0; c < 4; ++c) bulk_g2s(sB + s * TC_BLOCK_BYTES + c * (TC_BLOCK_BYTES / 4), src + c * (TC_BLOCK_BYTES / 4), TC_BLOCK_BYTES / 4, full + s);
            }
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        if (lane == 0) {
            // ---- MMA issuer
            if (!mbar_wait(a_full, 0)) ok = false;
            const uint32_t a0 = smem_u32(sA);
            for (int t = 0; t < ntiles && ok; ++t) {
                const int s = t % TC_BSTAGES, u = t / TC_BSTAGES, a = t & 1, ua = t >> 1;
                if (!mbar_wait(full + s, u & 1)) { ok = false; break; }
                if (ua > 0 && !mbar_wait(acc_empty + a, (ua - 1) & 1)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b0 = smem_u32(sB + s * TC_BLOCK_BYTES), d = tmem_base + (uint32_t)a * TC_N;
#pragma unroll
                for (int kk = 0; kk < TC_K / 32; ++kk) {                 // K = 32 bytes per instruction = two 16-byte chunks
                    const uint64_t ad = make_smem_desc(a0 + 2 * kk * (TC_M * 16), TC_M * 16, 128);
                    const uint64_t bd = make_smem_desc(b0 + 2 * kk * (TC_N * 16), TC_N * 16, 128);
                    tc_mma_i8(d, ad, bd, kk > 0 ? 1u : 0u);
                }
                tc_commit(smem_free + s);          // B stage reusable once these MMAs have read it
                tc_commit(acc_full + a);           // accumulator stage complete
            }
        }
    } else {
        // ---- epilogue warps.  Lane quarter q = warp & 3 (hardware rule: a warp reads TMEM lanes 32*(warp%4)..+31), column
        // half h = warp >> 2.  v = 256 - 2*hamming, so "hamming < best.d1" is "v > thr".  Per 32-column chunk: one
        // tcgen05.ld, then per 8-column group a max and a WARP-UNIFORM branch into a predicated (branch-free) insert
        // sequence.  (Per-element branches made the first version epilogue-bound: 8.7 k warp instructions per tile.)
        const int q = warp & 3, h = warp >> 2;
        int thr = INT_MIN;
        for (int t = 0; t < ntiles && ok; ++t) {
            const int a = t & 1, ua = t >> 1;
            if (!mbar_wait(acc_full + a, ua & 1)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int t_row0 = (tile0 + t) * TC_N, t_rows = min(TC_N, pd.nt - t_row0);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * TC_N + (uint32_t)h * (TC_N / 2);
#pragma unroll 1
            for (int cc = 0; cc < TC_N / 2; cc += 32) {
                const int c0 = h * (TC_N / 2) + cc;
                if (c0 >= t_rows) break;                                  // warp-uniform
                uint32_t v[32];
                tc_ld32(taddr + cc, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // two-level filter: 8-wide group maxima, one vote per group; only groups in which SOME lane has a candidate
                // better than its current second-best run the predicated insert sequence (expected: a fraction of a group
                // per chunk once a few hundred candidates have been seen)
                const bool partial = c0 + 32 > t_rows;                    // zero-padded rows would look like hamming 128
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    int m = (int)v[8 * g];
#pragma unroll
                    for (int j = 1; j < 8; ++j) m = max(m, (int)v[8 * g + j]);
                    if (__any_sync(0xffffffffu, m > thr) || partial) {
#pragma unroll
                        for (int j = 8 * g; j < 8 * g + 8; ++j) {
                            const int d = (TC_K - (int)v[j]) >> 1, idx = t_row0 + c0 + j;
                            const bool valid = !partial || (c0 + j < t_rows);
                            const bool lt0 = valid && d < best.d0, lt1 = valid && d < best.d1;
                            const int nd1 = lt0 ? best.d0 : (lt1 ? d : best.d1), ni1 = lt0 ? best.i0 : (lt1 ? idx : best.i1);
                            best.d0 = lt0 ? d : best.d0; best.i0 = lt0 ? idx : best.i0; best.d1 = nd1; best.i1 = ni1;
                        }
                        thr = best.d1 == INT_MAX ? INT_MIN : TC_K - 2 * best.d1;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + a);
        }
        if (h == 1) half_best[q * 32 + lane] = make_int4(best.d0, best.i0, best.d1, best.i1);
    }
    if (!ok) atomicExch(error_flag, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp < 4) {
        // merge the two column halves: lexicographic (distance, index), the order candidates would have arrived in
        const int4 o = half_best[warp * 32 + lane];
        auto lex_lt = [](int d, int i, int d2, int i2) { return d < d2 || (d == d2 && i < i2); };
        auto ins = [&](int d, int i) {
            if (i < 0) return;
            if (best.i0 < 0 || lex_lt(d, i, best.d0, best.i0)) { best.d1 = best.d0; best.i1 = best.i0; best.d0 = d; best.i0 = i; }
            else if (best.i1 < 0 || lex_lt(d, i, best.d1, best.i1)) { best.d1 = d; best.i1 = i; }
        };
        ins(o.x, o.y); ins(o.z, o.w);
        const int row = q_row0 + warp * 32 + lane;
        if (row < pd.nq) partial[(size_t)(pd.out_row + row) * splits + sp] = make_int4(best.d0, best.i0, best.d1, best.i1);
    }
    if (warp == TC_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

int match_tc_splits(int sm_count, int n_pairs, int nq_max, int nt_max) {
    const int qblocks = ceil_div(nq_max, TC_M), tiles = ceil_div(nt_max, TC_N);
    int splits = 1;
    while ((int64_t)qblocks * n_pairs * splits < 2LL * sm_count && splits * 2 <= tiles && splits < 16) splits *= 2;
    return splits;
}

size_t match_tc_block_bytes() { return TC_BLOCK_BYTES; }
int match_tc_block_rows() { return TC_N; }

int match_tc_expand(sfmb200_ctx* ctx, const uint32_t* d_desc, const int2* d_blocks, int n_blocks, uint8_t* d_E) {
    if (n_blocks == 0) return SFMB200_OK;
    expand_blocks_kernel<<<n_blocks, 256, 0, ctx->stream>>>(d_desc, d_blocks, d_E);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}

int match_tc_launch(sfmb200_ctx* ctx, const uint8_t* d_E, const PairDesc* d_pairs, int n_pairs, int nq_max, int splits,
                    int4* d_partial, int* d_error_flag) {
    // the attribute is per device and the context is per device; ctx->mu is held by every caller
    if (!ctx->tc_attr_set) { SFM_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)); ctx->tc_attr_set = true; }
    const int qblocks = ceil_div(nq_max, TC_M);
    dim3 grid(qblocks * splits, n_pairs);
    knn2_hamming_tc_kernel<<<grid, TC_THREADS, TC_SMEM, ctx->stream>>>(d_E, d_pairs, qblocks, splits, d_partial, d_error_flag);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}
