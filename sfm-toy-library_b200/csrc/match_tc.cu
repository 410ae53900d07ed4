// match_tc.cu -- K1 on the 5th-generation tensor cores: Hamming knn2 as an exact integer GEMM (tcgen05 + TMEM).
//
// For 256-bit descriptors  ham(q,t) = popc(q) + popc(t) - 2 * <q,t>  with q,t in {0,1}^256: the
// inner products of a 128-query x 256-train tile are ONE accumulator tile of  tcgen05.mma.kind::i8  (u8 x u8 -> s32,
// exact: every sum <= 256), M=128 N=256 K=32 per instruction, 8 instructions per tile, accumulators in TMEM.
//   * operands: the packed 32-byte descriptors are read from HBM as they are (no expanded copy) and expanded to one
//     byte per bit inside the kernel while they are written to shared memory in the UMMA K-major, no-swizzle
//     ("interleaved") canonical layout: core matrix = 8 rows x 16 bytes, LBO = distance between the two 16-byte K chunks
//     of one instruction, SBO = distance between 8-row groups (cute/arch/mma_sm100_desc.hpp, make_umma_desc<Major::K>).
//   * one elected thread issues the 8 MMAs of a tile and commits them to an mbarrier; the 128 threads (one per query
//     row = one per TMEM lane) then pull the accumulators with tcgen05.ld.32x32b.x32 and keep the running top-2 in
//     registers with exactly the ordering of the XOR/POPC kernel (strict '<', ascending train index), so the result is
//     bit-identical and feeds the same merge / ratio-test / compaction epilogue (match.cu).
//   * two accumulator stages (2 x 256 of the 512 TMEM columns) and two B stages in shared memory: the MMAs of tile t+1
//     run while tile t is being reduced.
// Only descriptor width 32 bytes (ORB, the reference's case); other widths use the XOR/POPC kernel.
#include "common.cuh"
#include "match_common.cuh"

namespace {

constexpr int TC_M = 128;            // query rows per CTA  (UMMA M, TMEM lanes)
constexpr int TC_N = 256;            // train rows per tile (UMMA N, TMEM columns per stage)
constexpr int TC_K = 256;            // descriptor bits = K elements (one byte each after expansion)
constexpr int TC_THREADS = 128;
constexpr int TC_A_BYTES = TC_M * TC_K;      // 32 KB
constexpr int TC_B_BYTES = TC_N * TC_K;      // 64 KB per stage
constexpr int TC_SMEM = TC_A_BYTES + 2 * TC_B_BYTES + 2 * TC_N * 4 + 64;
// instruction descriptor (UMMA::InstrDescriptor): c_format S32 (2) @bit4, a/b format UINT8 (0), K-major both,
// n_dim = N>>3 @bit17, m_dim = M>>4 @bit24
constexpr uint32_t TC_IDESC = (2u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}

// byte offset of (row r, 16-byte K chunk kc) inside a tile of R rows in the canonical layout:
//   chunk-major: [kc][r/8][r%8][16 B]  ->  LBO = R*16 bytes between K chunks, SBO = 128 bytes between 8-row groups
__device__ __forceinline__ uint32_t tile_offset(int r, int kc, int R) { return (uint32_t)kc * (R * 16) + (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16; }

// 16 descriptor bits -> 16 bytes of 0/1 (little endian bit order = element order k)
__device__ __forceinline__ uint4 expand16(uint32_t bits16) {
    uint4 o;
    o.x = ((bits16 & 0xF) * 0x00204081u) & 0x01010101u;
    o.y = (((bits16 >> 4) & 0xF) * 0x00204081u) & 0x01010101u;
    o.z = (((bits16 >> 8) & 0xF) * 0x00204081u) & 0x01010101u;
    o.w = (((bits16 >> 12) & 0xF) * 0x00204081u) & 0x01010101u;
    return o;
}

// expand `rows` packed descriptors (8 words each, row-major in global memory) into a canonical tile of R rows;
// rows beyond `rows` are zero-filled.  Also writes popcounts (may be null).  All TC_THREADS threads participate.
__device__ __forceinline__ void stage_tile(const uint32_t* __restrict__ src, int rows, int R, uint8_t* tile, int* popc) {
    // work item = (row, word): 8 words per row, each word = two 16-byte chunks
    for (int it = threadIdx.x; it < R * 8; it += TC_THREADS) {
        const int w = it / R, r = it - w * R;          // consecutive threads = consecutive rows: contiguous 16-byte stores
        const uint32_t v = r < rows ? __ldg(src + (size_t)r * 8 + w) : 0u;
        *reinterpret_cast<uint4*>(tile + tile_offset(r, 2 * w, R)) = expand16(v & 0xFFFFu);
        *reinterpret_cast<uint4*>(tile + tile_offset(r, 2 * w + 1, R)) = expand16(v >> 16);
    }
    if (popc) {
        for (int r = threadIdx.x; r < R; r += TC_THREADS) {
            int c = 0;
            if (r < rows) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * 8)), b = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * 8 + 4));
                c = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
            }
            popc[r] = c;
        }
    }
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
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

__global__ void __launch_bounds__(TC_THREADS) knn2_hamming_tc_kernel(const uint32_t* __restrict__ desc, const PairDesc* __restrict__ pairs, int qblocks,
                                                                     int splits, int4* __restrict__ partial, int* __restrict__ error_flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_A_BYTES;                                      // two stages
    int* sPopc = reinterpret_cast<int*>(smem + TC_A_BYTES + 2 * TC_B_BYTES);     // [2][TC_N]
    __shared__ __align__(8) uint64_t mma_done[2];
    __shared__ uint32_t tmem_base_s;

    const PairDesc pd = pairs[blockIdx.y];
    const int qb = blockIdx.x / splits, sp = blockIdx.x % splits;
    if (qb * TC_M >= pd.nq) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // train rows of this split, tile-aligned split boundaries keep the ascending-index order inside a split
    const int tiles_total = (pd.nt + TC_N - 1) / TC_N;
    const int tiles_per_split = (tiles_total + splits - 1) / splits;
    const int tile0 = sp * tiles_per_split, tile1 = min(tiles_total, tile0 + tiles_per_split);

    if (threadIdx.x == 0) { mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {        // TMEM: all 512 columns = two accumulator stages of 128 lanes x 256 columns (s32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A tile: this CTA's queries (rows beyond nq are zero)
    const int q_row0 = qb * TC_M, q_rows = min(TC_M, pd.nq - q_row0);
    stage_tile(desc + (size_t)(pd.q_row + q_row0) * 8, q_rows, TC_M, sA, nullptr);
    int pq = 0;
    {
        const int r = threadIdx.x;
        if (r < q_rows) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(pd.q_row + q_row0 + r) * 8)), b = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(pd.q_row + q_row0 + r) * 8 + 4));
            pq = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    Top2 best = {INT_MAX, -1, INT_MAX, -1};
    bool ok = true;
    uint32_t phase[2] = {0, 0};
    const int ntiles = tile1 - tile0;

    auto produce = [&](int t) {      // stage train tile t into B stage (t&1) and launch its MMAs into accumulator stage (t&1)
        const int st = t & 1;
        const int t_row0 = (tile0 + t) * TC_N, t_rows = min(TC_N, pd.nt - t_row0);
        stage_tile(desc + (size_t)(pd.t_row + t_row0) * 8, t_rows, TC_N, sB + st * TC_B_BYTES, sPopc + st * TC_N);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + st * TC_B_BYTES), d = tmem_base + (uint32_t)st * TC_N;
#pragma unroll
            for (int kk = 0; kk < TC_K / 32; ++kk) {                     // K = 32 bytes per instruction = two 16-byte chunks
                const uint64_t ad = make_smem_desc(a0 + 2 * kk * (TC_M * 16), TC_M * 16, 128);
                const uint64_t bd = make_smem_desc(b0 + 2 * kk * (TC_N * 16), TC_N * 16, 128);
                tc_mma_i8(d, ad, bd, kk > 0 ? 1u : 0u);
            }
            tc_commit(&mma_done[st]);
        }
    };

    if (ntiles > 0) produce(0);
    for (int t = 0; t < ntiles; ++t) {
        const int st = t & 1;
        if (t + 1 < ntiles) produce(t + 1);           // overlaps with the tensor core working on tile t
        if (ok && !mbar_wait(&mma_done[st], phase[st])) ok = false;
        phase[st] ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            const int t_row0 = (tile0 + t) * TC_N, t_rows = min(TC_N, pd.nt - t_row0);
            const int* pc = sPopc + st * TC_N;
            // this warp owns TMEM lanes [32*warp, 32*warp+32): lane field in bits [16,32) of the address
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)st * TC_N;
#pragma unroll 1
            for (int c0 = 0; c0 < TC_N; c0 += 32) {
                if (c0 >= t_rows) break;                                  // warp-uniform
                uint32_t v[32];
                tc_ld32(taddr + c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int d = pq + pc[c0 + j] - 2 * (int)v[j];
                    if (d < best.d1 && c0 + j < t_rows) top2_insert(best, d, t_row0 + c0 + j);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                 // accumulator stage `st` and its popcounts are free again
    }
    if (!ok && threadIdx.x == 0) atomicExch(error_flag, 1);
    {
        const int row = q_row0 + threadIdx.x;
        if (row < pd.nq) partial[(size_t)(pd.out_row + row) * splits + sp] = make_int4(best.d0, best.i0, best.d1, best.i1);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

int match_tc_splits(int sm_count, int n_pairs, int nq_max, int nt_max) {
    const int qblocks = ceil_div(nq_max, TC_M), tiles = ceil_div(nt_max, TC_N);
    int splits = 1;
    while ((int64_t)qblocks * n_pairs * splits < 2LL * sm_count && splits * 2 <= tiles && splits < 16) splits *= 2;
    return splits;
}

int match_tc_launch(sfmb200_ctx* ctx, const uint32_t* d_desc, const PairDesc* d_pairs, int n_pairs, int nq_max, int splits,
                    int4* d_partial, int* d_error_flag) {
    static bool attr_set = false;
    if (!attr_set) { SFM_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)); attr_set = true; }
    const int qblocks = ceil_div(nq_max, TC_M);
    dim3 grid(qblocks * splits, n_pairs);
    knn2_hamming_tc_kernel<<<grid, TC_THREADS, TC_SMEM, ctx->stream>>>(d_desc, d_pairs, qblocks, splits, d_partial, d_error_flag);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}
