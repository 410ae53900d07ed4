// match_tc.cu -- K1 on the 5th-generation tensor cores: knn2 as an exact integer GEMM (tcgen05 + TMEM), two instantiations.
//
// HAMMING (the reference's ORB case, SfM2DFeatureUtilities.cpp:39-40, 59-60).  For 256-bit descriptors
//   ham(q,t) = popc(q) + popc(t) - 2 <q,t>  with q,t in {0,1}^256; bits are mapped to SIGNED bytes +1/-1 so that
//   <a,b> = 256 - 2*hamming: one accumulator tile of  tcgen05.mma.kind::i8  (s8 x s8 -> s32, exact), M=128 N=256 K=32 per
//   instruction, 8 instructions per tile.
// L2 (BASELINE.json configs[3] wording: SIFT-128; cv::BFMatcher(NORM_L2)).  SIFT descriptors are integer-valued 0..255, so
//   |a-b|^2 = |a|^2 + |b|^2 - 2 <a,b>  with <a,b> <= 128*255^2 < 2^31 is EXACT in u8 x u8 -> s32: 4 instructions per tile; the
//   train norms ride along in shared memory, the epilogue ranks  e = |b|^2 - 2<a,b>  (as packed keys e * 128 + column, one
//   multiply-add per element, see l2_chunk) and adds |a|^2 at the end.
//
//   * operands: expanded ONCE per descriptor set (expand kernels) into blocks of 256 rows in the UMMA K-major, no-swizzle
//     ("interleaved") canonical form: core matrix = 8 rows x 16 bytes, LBO = distance between the 16-byte K chunks, SBO =
//     distance between 8-row groups (cute/arch/mma_sm100_desc.hpp, make_umma_desc<Major::K>).  A tile travels HBM -> shared
//     memory as plain cp.async.bulk copies completing on an mbarrier (no tensor map needed: a block is contiguous).
//   * warp-specialised pipeline: a loader thread, an MMA thread (two 256-column accumulator stages = all 512 TMEM columns;
//     tcgen05.commit releases the smem stage and publishes the accumulator) and 8 epilogue warps (two per TMEM lane quarter,
//     thread = query row) that pull the accumulators with tcgen05.ld.32x32b.x32.
//   * epilogue = the bound (round 1: ALU pipe 70 % busy, tensor pipe idle).  Hamming now keeps the running top-2 as two PACKED
//     KEYS  key = (256 - 2 ham) << 16 | (0xFFFF - index)  -- larger is better, ties go to the lower index exactly like
//     cv::batchDistance -- so an insert is a 3-instruction max/min network (no compare/select chains, no separate index
//     registers), groups of 8 columns are skipped when no lane's group maximum (3-input max, DPX) beats its second best.
// Bit-exact against the XOR/POPC kernel, the oracle and cv2 (tests/test_gpu_match.py); same merge / ratio / compaction epilogue.
#include "common.cuh"
#include "match_common.cuh"

namespace {

constexpr int TC_M = 128;            // query rows per CTA  (UMMA M, TMEM lanes)
constexpr int TC_N = 256;            // train rows per tile (UMMA N, TMEM columns per accumulator stage) = rows per expanded block
constexpr int TC_EPI_PARTS = 4;      // column parts of a 256-column accumulator tile, one epilogue warp per (TMEM lane quarter, part)
constexpr int TC_EPI_WARPS = 4 * TC_EPI_PARTS;      // 16 warps: the epilogue is latency-bound (tcgen05.ld, votes), more warps hide it
constexpr int TC_PART_COLS = TC_N / TC_EPI_PARTS;   // 64 columns = 2 chunks of 32 per warp and tile
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;      // + warp 8: loader, warp 9: MMA issuer
constexpr int TC_NORM_SLOTS = 6;     // L2: ring of train-norm tiles (1 KB each); a slot is reused 6 tiles later, when its epilogue is long done

template <bool L2> struct TcCfg {
    static constexpr int KB = L2 ? 128 : 256;                  // operand bytes per row = K elements
    static constexpr int BLOCK_BYTES = TC_N * KB;              // one 256-row block of expanded operands
    static constexpr int A_BYTES = TC_M * KB;
    static constexpr int BSTAGES = L2 ? 4 : 3;                 // shared-memory stages of the train operand
    static constexpr int NORM_BYTES = L2 ? TC_NORM_SLOTS * TC_N * 4 : 0;
    static constexpr int SMEM = A_BYTES + BSTAGES * BLOCK_BYTES + NORM_BYTES + 256;
    // instruction descriptor (UMMA::InstrDescriptor): c_format S32 (2) @bit4, a/b format @bits 7/10 (0 = unsigned, 1 = signed int8),
    // K-major both, n_dim = N>>3 @bit17, m_dim = M>>4 @bit24
    static constexpr uint32_t IDESC = (2u << 4) | ((L2 ? 0u : 1u) << 7) | ((L2 ? 0u : 1u) << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
};

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

// Expanded operand store (built once per descriptor set): per 256-row block the UMMA K-major no-swizzle canonical layout
//   [16-byte K chunk kc][row group r/8][r%8][16 B]   ->  LBO = 4096 B between K chunks, SBO = 128 B between 8-row groups.
// A 128-row half of a block is the same layout at start offset +2048 B per chunk, so one store serves both the query (M=128)
// and the train (N=256) operand.  Rows beyond the image are zero (contribute nothing).
__global__ void __launch_bounds__(256) expand_blocks_kernel(const uint32_t* __restrict__ desc, const int2* __restrict__ blocks /* (first row, valid rows) */,
                                                            uint8_t* __restrict__ E) {
    const int2 b = blocks[blockIdx.x];
    uint8_t* out = E + (size_t)blockIdx.x * TcCfg<false>::BLOCK_BYTES;
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
// L2: float descriptors [rows][dim] (dim <= 128, zero-padded to 128) -> u8 operand blocks + squared norms.  *bad is raised when
// a value is not an integer in [0, 255] (then the exact-GEMM formulation does not apply and the caller falls back to fp32).
__global__ void __launch_bounds__(256) expand_l2_blocks_kernel(const float* __restrict__ desc, int dim, const int2* __restrict__ blocks,
                                                               uint8_t* __restrict__ E, int32_t* __restrict__ norms, int* __restrict__ bad) {
    const int2 b = blocks[blockIdx.x];
    uint8_t* out = E + (size_t)blockIdx.x * TcCfg<true>::BLOCK_BYTES;
    const int r = threadIdx.x;                                              // one row per thread
    const bool valid = r < b.y;
    const float* src = desc + (size_t)(b.x + (valid ? r : 0)) * dim;
    int nrm = 0; bool ok = true;
    const uint32_t off = (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
    for (int kc = 0; kc < 8; ++kc) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = kc * 16 + j;
            float f = (valid && k < dim) ? __ldg(src + k) : 0.f;
            const int iv = (int)f;
            ok = ok && (f == (float)iv) && iv >= 0 && iv <= 255;
            const int c = min(max(iv, 0), 255);
            nrm += c * c;
            w[j >> 2] |= (uint32_t)c << (8 * (j & 3));
        }
        *reinterpret_cast<uint4*>(out + (size_t)kc * (TC_N * 16) + off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // packed for the epilogue's key arithmetic: |b|^2 * 128 + (row & 127: the column inside its 128-column half of the tile); |b|^2 <= 128 * 255^2 < 2^23
    norms[(size_t)blockIdx.x * TC_N + r] = ((valid ? nrm : 0) << 7) | (r & 127);
    if (!ok) atomicExch(bad, 1);
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
template <bool L2> __device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TcCfg<L2>::IDESC), "r"(accumulate), "r"(0u) : "memory");
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

// Hamming key: (256 - 2 ham) in the high half, 0xFFFF - (train index inside this CTA's split) in the low half.  The split of a
// CTA covers at most 256 tiles = 65536 rows (match_tc_splits), so the index fits.  EMPTY = INT_MIN sorts below every key.
constexpr int KEY_EMPTY = INT_MIN;

// One 32-column chunk of the Hamming epilogue.  v = 256 - 2 ham (s32 accumulators of this lane's query row).  Per group of 8
// columns: a 3-input-max tree, one vote, and -- only when SOME lane of the warp has a candidate above its second best -- the
// insert network on packed keys: {k0, k1, key} -> the two largest (2 min/max per element + a 3-input max per two elements).
// PART: columns >= ncols are zero padding of the image's last tile and must not become candidates.
template <bool PART>
__device__ __forceinline__ void hamming_chunk(const uint32_t (&v)[32], int cb, int ncols, int& k0, int& k1, int& thr) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        int m = __vimax3_s32((int)v[8 * g], (int)v[8 * g + 1], (int)v[8 * g + 2]);
        m = __vimax3_s32(m, (int)v[8 * g + 3], (int)v[8 * g + 4]);
        m = __vimax3_s32(m, (int)v[8 * g + 5], (int)v[8 * g + 6]);
        m = max(m, (int)v[8 * g + 7]);
        if (PART || __any_sync(0xffffffffu, m > thr)) {
#pragma unroll
            for (int j = 8 * g; j < 8 * g + 8; ++j) {
                int key = (int)v[j] * 65536 + (cb - j);
                if (PART && j >= ncols) key = KEY_EMPTY;
                const int hi = max(k0, key), lo = min(k0, key);
                k1 = max(k1, lo); k0 = hi;
            }
            thr = k1 >> 16;             // a later candidate needs a strictly larger v: with equal v its index loses
        }
    }
}

// L2 key: e * 128 + (column inside the 128-column half of the tile), e = |b|^2 - 2<a,b> in [-2 * 128 * 255^2, 128 * 255^2] -- 25 bits, so the key
// fits a signed 32-bit word exactly and ONE multiply-add makes it from the accumulator: key = nk - (acc << 8) with the packed norm
// nk = |b|^2 * 128 + column that rides along in shared memory.  Smaller key = smaller distance, ties to the smaller column: the order
// of cv::batchDistance inside a tile.  The keys only live for one tile (7 index bits): per 8-column group a 3-input-min tree and a
// vote against thrk (the running second best of the whole row as a key bound, tightened by the tile's own second best), then the
// {k0, k1, key} -> two smallest network; after the tile the two survivors are unpacked and merged into the (distance, index) pairs.
constexpr int L2KEY_EMPTY = INT_MAX;
template <bool PART>
__device__ __forceinline__ void l2_chunk(const uint32_t (&v)[32], const int32_t* __restrict__ nk, int ncols, int& k0, int& k1, int& thrk) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int4 n0 = *reinterpret_cast<const int4*>(nk + 8 * g), n1 = *reinterpret_cast<const int4*>(nk + 8 * g + 4);
        int key[8] = {n0.x - ((int)v[8 * g] << 8), n0.y - ((int)v[8 * g + 1] << 8), n0.z - ((int)v[8 * g + 2] << 8), n0.w - ((int)v[8 * g + 3] << 8),
                      n1.x - ((int)v[8 * g + 4] << 8), n1.y - ((int)v[8 * g + 5] << 8), n1.z - ((int)v[8 * g + 6] << 8), n1.w - ((int)v[8 * g + 7] << 8)};
        if (PART) {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (8 * g + j >= ncols) key[j] = L2KEY_EMPTY;
        }
        int m = __vimin3_s32(key[0], key[1], key[2]);
        m = __vimin3_s32(m, key[3], key[4]); m = __vimin3_s32(m, key[5], key[6]); m = min(m, key[7]);
        if (PART || __any_sync(0xffffffffu, m < thrk)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int lo = min(k0, key[j]), hi = max(k0, key[j]);
                k1 = min(k1, hi); k0 = lo;
            }
            thrk = min(thrk, k1);
        }
    }
}

// Warp-specialised: loader thread (cp.async.bulk of pre-expanded blocks) -> MMA thread (KB/32 x tcgen05.mma per tile into one
// of 2 TMEM accumulator stages, commits release the smem stage and publish the accumulator) -> 8 epilogue warps (tcgen05.ld,
// running top-2 per query row in registers).
template <bool L2>
__global__ void __launch_bounds__(TC_THREADS) knn2_tc_kernel(const uint8_t* __restrict__ E, const int32_t* __restrict__ norms, const PairDesc* __restrict__ pairs,
                                                             int qblocks, int splits, int4* __restrict__ partial, int* __restrict__ error_flag) {
    using C = TcCfg<L2>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::A_BYTES;
    int32_t* sNorm = reinterpret_cast<int32_t*>(smem + C::A_BYTES + C::BSTAGES * C::BLOCK_BYTES);      // [TC_NORM_SLOTS][TC_N]   (L2)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::A_BYTES + C::BSTAGES * C::BLOCK_BYTES + C::NORM_BYTES);
    uint64_t* a_full = bars;                 // [1]
    uint64_t* full = bars + 1;               // [BSTAGES] bytes of a B stage have landed
    uint64_t* smem_free = full + C::BSTAGES; // [BSTAGES] the MMAs that read the stage have completed
    uint64_t* acc_full = smem_free + C::BSTAGES;   // [2] accumulator stage holds a finished tile
    uint64_t* acc_empty = acc_full + 2;      // [2] the epilogue warps have drained the stage
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
        for (int i = 0; i < C::BSTAGES; ++i) { mbar_init(full + i, 1); mbar_init(smem_free + i, 1); }
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

    // Hamming: packed keys; L2: (e = |b|^2 - 2<a,b>, global index) pairs
    int k0 = KEY_EMPTY, k1 = KEY_EMPTY;
    Top2 best = {INT_MAX, -1, INT_MAX, -1};
    // top-2 of the upper column half, merged by the lower-half warp at the end.  Lives in B stage 0: every bulk copy into a B
    // stage is consumed (waited on) before the last accumulator is published, whereas the query-tile copy into sA may still
    // be in flight when a split owns no train tile at all.
    int4* half_best = reinterpret_cast<int4*>(sB);
    if (warp == TC_EPI_WARPS) {
        if (lane == 0) {
            // ---- loader: query tile (a 128-row half of a block: KB/16 chunks of 2 KB), then the train blocks (+ their norms)
            const uint8_t* qsrc = E + (size_t)(pd.q_blk + q_row0 / TC_N) * C::BLOCK_BYTES + (size_t)((q_row0 / TC_M) & 1) * (TC_M * 16);
            mbar_expect_tx(a_full, C::A_BYTES);
            for (int kc = 0; kc < C::KB / 16; ++kc) bulk_g2s(sA + kc * (TC_M * 16), qsrc + (size_t)kc * (TC_N * 16), TC_M * 16, a_full);
            for (int t = 0; t < ntiles && ok; ++t) {
                const int s = t % C::BSTAGES, u = t / C::BSTAGES;
                if (u > 0 && !mbar_wait(smem_free + s, (u - 1) & 1)) { ok = false; break; }
                const uint8_t* src = E + (size_t)(pd.t_blk + tile0 + t) * C::BLOCK_BYTES;
                mbar_expect_tx(full + s, C::BLOCK_BYTES + (L2 ? TC_N * 4 : 0));
                for (int c = 0; c < 4; ++c) bulk_g2s(sB + s * C::BLOCK_BYTES + c * (C::BLOCK_BYTES / 4), src + c * (C::BLOCK_BYTES / 4), C::BLOCK_BYTES / 4, full + s);
                if (L2) bulk_g2s(sNorm + (t % TC_NORM_SLOTS) * TC_N, norms + (size_t)(pd.t_blk + tile0 + t) * TC_N, TC_N * 4, full + s);
            }
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        if (lane == 0) {
            // ---- MMA issuer
            if (!mbar_wait(a_full, 0)) ok = false;
            const uint32_t a0 = smem_u32(sA);
            for (int t = 0; t < ntiles && ok; ++t) {
                const int s = t % C::BSTAGES, u = t / C::BSTAGES, a = t & 1, ua = t >> 1;
                if (!mbar_wait(full + s, u & 1)) { ok = false; break; }
                if (ua > 0 && !mbar_wait(acc_empty + a, (ua - 1) & 1)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b0 = smem_u32(sB + s * C::BLOCK_BYTES), d = tmem_base + (uint32_t)a * TC_N;
#pragma unroll
                for (int kk = 0; kk < C::KB / 32; ++kk) {                // K = 32 bytes per instruction = two 16-byte chunks
                    const uint64_t ad = make_smem_desc(a0 + 2 * kk * (TC_M * 16), TC_M * 16, 128);
                    const uint64_t bd = make_smem_desc(b0 + 2 * kk * (TC_N * 16), TC_N * 16, 128);
                    tc_mma_i8<L2>(d, ad, bd, kk > 0 ? 1u : 0u);
                }
                tc_commit(smem_free + s);          // B stage reusable once these MMAs have read it
                tc_commit(acc_full + a);           // accumulator stage complete
            }
        }
    } else {
        // ---- epilogue warps.  Lane quarter q = warp & 3 (hardware rule: a warp reads TMEM lanes 32*(warp%4)..+31), column
        // part h = warp >> 2.  Per 32-column chunk: one tcgen05.ld, then per 8-column group a 3-input-max tree and a
        // WARP-UNIFORM branch: only groups in which SOME lane has a candidate better than its current second best run the
        // insert sequence (expected: a third of the groups over a 5000-row image).
        const int q = warp & 3, h = warp >> 2;
        int thr = L2 ? INT_MAX : INT_MIN;                                 // Hamming: v must exceed it; L2: e must be below it
        for (int t = 0; t < ntiles && ok; ++t) {
            const int a = t & 1, ua = t >> 1;
            if (!mbar_wait(acc_full + a, ua & 1)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int t_row0 = (tile0 + t) * TC_N, t_rows = min(TC_N, pd.nt - t_row0);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * TC_N + (uint32_t)h * TC_PART_COLS;
            const int32_t* nrm = sNorm + (t % TC_NORM_SLOTS) * TC_N;
            int kt0 = L2KEY_EMPTY, kt1 = L2KEY_EMPTY;                         // L2: the tile's two smallest keys
            if (L2) thr = best.i1 < 0 ? INT_MAX : (best.d1 << 7);           // a later row must be STRICTLY closer than the running second best
            // one 32-column chunk of this warp's column half
            auto chunk = [&](const uint32_t (&v)[32], int cc) {
                const int c0 = h * TC_PART_COLS + cc;
                const bool part = c0 + 32 > t_rows;                       // zero-padded rows must not become candidates
                if (!L2) {
                    const int cb = 0xFFFF - (t * TC_N + c0);              // low half of the key of column 0 of this chunk
                    if (!part) hamming_chunk<false>(v, cb, 32, k0, k1, thr);
                    else hamming_chunk<true>(v, cb, t_rows - c0, k0, k1, thr);      // last tile of the image only
                } else {
                    if (!part) l2_chunk<false>(v, nrm + c0, 32, kt0, kt1, thr);
                    else l2_chunk<true>(v, nrm + c0, t_rows - c0, kt0, kt1, thr);
                }
            };
            // software pipeline over the 4 chunks: the tcgen05.ld of the next chunk is in flight while this one is ranked
            const int nch = min(TC_PART_COLS / 32, (t_rows - h * TC_PART_COLS + 31) / 32);  // chunks of this column part that hold real rows (warp-uniform)
            uint32_t va[32], vb[32];
            static_assert(TC_PART_COLS == 64, "the pipeline below is written for two chunks per warp and tile");
            if (nch > 0) { tc_ld32(taddr, va); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
            if (nch > 1) tc_ld32(taddr + 32, vb);
            if (nch > 0) chunk(va, 0);
            if (nch > 1) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); chunk(vb, 32); }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + a);
            if (L2) {                                                       // the tile's survivors -> (e, global index), ascending key order
                auto ins = [&](int key) {
                    if (key == L2KEY_EMPTY) return;
                    const int d = key >> 7, idx = t_row0 + (h >> 1) * 128 + (key & 127);       // this warp's 64 columns lie in half (h >> 1) of the tile
                    if (d < best.d0) { best.d1 = best.d0; best.i1 = best.i0; best.d0 = d; best.i0 = idx; }
                    else if (d < best.d1) { best.d1 = d; best.i1 = idx; }
                };
                ins(kt0); ins(kt1);
            }
        }
        if (h > 0) half_best[(h - 1) * TC_M + q * 32 + lane] = L2 ? make_int4(best.d0, best.i0, best.d1, best.i1) : make_int4(k0, k1, 0, 0);
    }
    if (!ok) atomicExch(error_flag, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp < 4) {
        const int row = q_row0 + warp * 32 + lane;
        if (!L2) {
            // merge the column parts with the same network, then unpack
#pragma unroll
            for (int part = 1; part < TC_EPI_PARTS; ++part) {
                const int4 o = half_best[(part - 1) * TC_M + warp * 32 + lane];
                int hi = max(k0, o.x), lo = min(k0, o.x); k1 = max(k1, lo); k0 = hi;
                hi = max(k0, o.y); lo = min(k0, o.y); k1 = max(k1, lo); k0 = hi;
            }
            auto unpack = [&](int key, int& d, int& i) {
                if (key == KEY_EMPTY) { d = INT_MAX; i = -1; return; }
                d = (256 - (key >> 16)) >> 1; i = tile0 * TC_N + (0xFFFF - (key & 0xFFFF));
            };
            int d0, i0, d1, i1; unpack(k0, d0, i0); unpack(k1, d1, i1);
            if (row < pd.nq) partial[(size_t)(pd.out_row + row) * splits + sp] = make_int4(d0, i0, d1, i1);
        } else {
            // lexicographic (e, index) merge, then |a|^2 + e = squared distance (exact integer)
            auto lex_lt = [](int d, int i, int d2, int i2) { return d < d2 || (d == d2 && i < i2); };
            auto ins = [&](int d, int i) {
                if (i < 0) return;
                if (best.i0 < 0 || lex_lt(d, i, best.d0, best.i0)) { best.d1 = best.d0; best.i1 = best.i0; best.d0 = d; best.i0 = i; }
                else if (best.i1 < 0 || lex_lt(d, i, best.d1, best.i1)) { best.d1 = d; best.i1 = i; }
            };
#pragma unroll
            for (int part = 1; part < TC_EPI_PARTS; ++part) {
                const int4 o = half_best[(part - 1) * TC_M + warp * 32 + lane];
                ins(o.x, o.y); ins(o.z, o.w);
            }
            if (row < pd.nq) {
                const int na = norms[(size_t)(pd.q_blk + q_row0 / TC_N) * TC_N + (q_row0 % TC_N) + warp * 32 + lane] >> 7;      // packed: |a|^2 * 128 + column
                partial[(size_t)(pd.out_row + row) * splits + sp] =
                    make_int4(best.i0 >= 0 ? na + best.d0 : INT_MAX, best.i0, best.i1 >= 0 ? na + best.d1 : INT_MAX, best.i1);
            }
        }
    }
    if (warp == TC_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

int match_tc_splits(int sm_count, int n_pairs, int nq_max, int nt_max) {
    const int qblocks = ceil_div(nq_max, TC_M), tiles = ceil_div(nt_max, TC_N);
    int splits = 1;
    while ((int64_t)qblocks * n_pairs * splits < 2LL * sm_count && splits * 2 <= tiles && splits < 16) splits *= 2;
    while (ceil_div(tiles, splits) > 256) splits *= 2;       // the 16-bit index field of the Hamming key: <= 256 tiles per split
    return splits;
}

size_t match_tc_block_bytes(bool l2) { return l2 ? TcCfg<true>::BLOCK_BYTES : TcCfg<false>::BLOCK_BYTES; }
int match_tc_block_rows() { return TC_N; }

int match_tc_expand(sfmb200_ctx* ctx, const uint32_t* d_desc, const int2* d_blocks, int n_blocks, uint8_t* d_E) {
    if (n_blocks == 0) return SFMB200_OK;
    expand_blocks_kernel<<<n_blocks, 256, 0, ctx->stream>>>(d_desc, d_blocks, d_E);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}
int match_tc_expand_l2(sfmb200_ctx* ctx, const float* d_desc, int dim, const int2* d_blocks, int n_blocks, uint8_t* d_E, int32_t* d_norms, int* d_bad) {
    if (n_blocks == 0) return SFMB200_OK;
    expand_l2_blocks_kernel<<<n_blocks, 256, 0, ctx->stream>>>(d_desc, dim, d_blocks, d_E, d_norms, d_bad);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}

int match_tc_launch(sfmb200_ctx* ctx, bool l2, const uint8_t* d_E, const int32_t* d_norms, const PairDesc* d_pairs, int n_pairs, int nq_max, int splits,
                    int4* d_partial, int* d_error_flag) {
    // the attribute is per device and the context is per device; ctx->mu is held by every caller
    if (!ctx->tc_attr_set) {
        SFM_CUDA(ctx, cudaFuncSetAttribute(knn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<false>::SMEM));
        SFM_CUDA(ctx, cudaFuncSetAttribute(knn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<true>::SMEM));
        ctx->tc_attr_set = true;
    }
    const int qblocks = ceil_div(nq_max, TC_M);
    dim3 grid(qblocks * splits, n_pairs);
    if (l2) knn2_tc_kernel<true><<<grid, TC_THREADS, TcCfg<true>::SMEM, ctx->stream>>>(d_E, d_norms, d_pairs, qblocks, splits, d_partial, d_error_flag);
    else knn2_tc_kernel<false><<<grid, TC_THREADS, TcCfg<false>::SMEM, ctx->stream>>>(d_E, d_norms, d_pairs, qblocks, splits, d_partial, d_error_flag);
    SFM_LAUNCH_CHECK(ctx);
    return SFMB200_OK;
}
