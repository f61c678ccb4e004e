// tcgen05 split-bf16 GEMM for the dense projections of the aggregator (SURVEY.md §8a rows a8, a11:
// Q/V projection snuffy.py:188, FFN snuffy.py:225).  These are the true contractions of the path.
//
//   C[m, n] = epilogue( sum_k A(m,k) * B(n,k) ),   A = A_hi + A_lo, B = B_hi + B_lo  (bf16 planes)
//   A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi       (3 tensor-core passes, fp32 accumulate in TMEM)
//
// which reproduces the fp32 product to ~2^-16 relative -- plain bf16/tf32 inputs break the 1e-4 parity bar and
// flip top-K indices.  Operands arrive as pre-tiled planes (common.cuh): every pipeline stage
// is a contiguous chunk in HBM moved by one cp.async.bulk per plane, landing directly in the SWIZZLE_NONE
// K-major core-matrix layout the UMMA shared-memory descriptor describes.  No tensor maps, no swizzle.
// The weight-gradient form (gemm_tc_kernel<BN, true>) contracts over the ROWS of two activation plane sets instead: its
// stages are strided tiles of those planes, fetched by one tensor-map TMA per operand and read through MN-major descriptors.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      bulk-copy producer (+ TMEM alloc/dealloc)          full/empty mbarrier ring
//   warp 1      single-thread tcgen05.mma issuer                    accumulators double-buffered in TMEM
//   warps 2..9  epilogue: tcgen05.ld -> bias/activation -> (a) split-bf16 planes for the next GEMM, written
//               straight from the row-owner register layout (512 B coalesced per warp store), and/or
//               (b) fp32 rows (+ residual read through row_map) transposed through shared memory.
#include <stdlib.h>
#include <mutex>
#include <cuda.h>                      // CUtensorMap (types only: the encoder is looked up through the runtime, no -lcuda)
#include "tc_ptx.cuh"

namespace snuffy {

void launch_fold_partials(const float* part, int splits, int64_t n4, float* out, cudaStream_t stream);

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 320;          // producer warp + MMA warp + 8 epilogue warps
constexpr uint32_t TC_A_PLANE_BYTES = TC_BM * PLANE_KB * 2;          // 8 KB per plane per stage

template <int BN> struct TcCfg {
    static constexpr uint32_t B_PLANE_BYTES = BN * PLANE_KB * 2;
    static constexpr uint32_t STAGE_BYTES = 2 * TC_A_PLANE_BYTES + 2 * B_PLANE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr uint32_t STG_BYTES = 8 * 32 * 33 * 4;             // epilogue transpose buffers (one per warp)
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 256;
    static constexpr uint32_t TMEM_COLS = 2 * BN;                      // two accumulator stages
};

#ifdef GEMM_DEBUG_TIMING
// clock64 stamps of CTA 0 (tools/gemm_timeline.py): [role 0 producer | 1 mma | 2 epilogue warp 2][item < 64][4 stamps]
__device__ long long g_gemm_dbg[3 * 64 * 4];
#define GDBG(role, t, k) do { if (blockIdx.x == 0 && (t) < 64) g_gemm_dbg[((role) * 64 + (t)) * 4 + (k)] = clock64(); } while (0)
#else
#define GDBG(role, t, k) do { } while (0)
#endif

struct TcGemmParams {
    const __nv_bfloat16* A; int64_t a_plane_stride;
    const __nv_bfloat16* B; int64_t b_plane_stride;
    int M, N, K;
    int m_tiles, n_tiles, num_kb;
    int npairs;                      // 3: hi.hi + hi.lo + lo.hi   1: hi.hi only
    const float* bias; int act;
    const float* resid; int64_t ldr; const int32_t* row_map; const float* resid_alt;
    float* out; int64_t ldc;
    float* preact;                   // optional fp32 [M, ldc] value before the activation
    __nv_bfloat16* out_planes; int64_t out_plane_stride;   // optional: result as A-operand planes, K_next = N
    float drop_p; uint64_t seed, offset;                   // dropout on the activated value (before the residual)
    // backward of an activation fused into the dX product: v *= act'(gate[m, n]) before the dropout mask (dh = (dY W) * act'(h_pre))
    const float* gate; int64_t ldg; int gate_act;
    // gate form only: column sums of the result per 32-row block, [m_tiles * 4][N] (folded by the host entry: the bias gradient
    // db = sum over rows of dh, so dh itself never has to be written as fp32)
    float* colsum_part;
    // ReLU backward read off the forward's activated operand planes instead of a saved fp32 pre-activation:
    // dh = (dY W) * (a > 0 ? gate_scale : 0) with a = dropout(relu(h)) as the forward wrote it (hi plane, layout of out_planes)
    // and gate_scale = 1 / (1 - p): a dropped or clamped element is 0 there, so neither h nor the dropout draw is needed again
    const __nv_bfloat16* gate_planes; float gate_scale;
    // block-diagonal B operand (attention backward against head-block matrices): B[n, k] is non-zero only where
    // n / group_n == k / group_k, so a column tile contracts only over the k-blocks of the groups it touches (group_n == 0: dense)
    int group_n, group_k;
    // block-diagonal OUTPUT (dKp = diagonal blocks of dS^T Q): only tiles that meet a block with row / diag_m == col / diag_n are
    // computed, the others are skipped by all three roles (diag_m == 0: every tile)
    int diag_m, diag_n;
    // split-K (dW = dY^T X contracts over all N patches but has few output tiles): tile = (mt, nt, ks), split ks covers
    // k-blocks [ks*kb_per, ...) and writes its raw fp32 partial to out + ks*split_stride (folded by the caller)
    int ksplit, kb_per; int64_t split_stride;
    // K window of the A planes: the operand is columns [a_kb_off*32, ...) of a wider plane set with a_nkb k-blocks per row
    // tile (e.g. Q or V inside the [rows, 2d] planes the Q|V projection wrote)
    int a_nkb, a_kb_off;
    // MN-major form (gemm_tc_kernel<BN, true>): both operands are ROW planes of activations ([rows, features], 128 rows per
    // chunk) contracted over their rows: A = planes of dY [K rows, M features], B = planes of X [K rows, N features], k-blocks
    // of 32 rows, fetched through the two tensor maps passed beside this struct
    // Tail split ("stream-K" for the last, partial wave): items [0, full_items) are whole tiles (times ksplit); the remaining
    // tail_tiles tiles, which would keep only tail_tiles of the SMs busy for a whole tile time, are cut into tail_s K slices of
    // tail_per k-blocks, one CTA each.  Slice 0 owns the tile's epilogue; the others hand their raw accumulators over through
    // tail_part [tail tile][slice - 1][128 x BN values in register order] and count themselves in tail_flags[2 * tile] (8 warps each); the owner adds
    // them in slice order (deterministic) and re-arms the flags.
    int full_items, tail_tiles, tail_s, tail_per;
    float* tail_part; unsigned int* tail_flags;
};

// One work item of a CTA's persistent loop.
struct TcItem { int mt, nt, ksp, kb0, kb1, mode, tt, slice; };     // mode -1 skip, 0 whole tile, 1 owning tail slice, 2 other tail slice

// out of line on purpose: 32 inlined copies of the activation switch per chunk bloat the epilogue (instruction-cache misses
// slow the whole CTA down); ReLU, the common case, is handled inline
__device__ __noinline__ float act_grad_call(int act, float t) { return act_grad(act, t); }

// k-blocks [kb0, kb1) a tile contracts over: its K split, narrowed to the non-zero window of a block-diagonal B operand (any
// superset of the window is exact: the extra blocks only multiply zeros)
template <int BN>
__device__ __forceinline__ void tile_kb_range(const TcGemmParams& p, int ksp, int nt, int& kb0, int& kb1) {
    kb0 = ksp * p.kb_per;
    kb1 = min(p.num_kb, kb0 + p.kb_per);
    if (p.group_n > 0) {
        const int c_lo = nt * BN, c_hi = min(p.N, c_lo + BN) - 1;
        const int k_lo = (c_lo / p.group_n) * p.group_k, k_hi = (c_hi / p.group_n + 1) * p.group_k;
        kb0 = max(kb0, k_lo / PLANE_KB);
        kb1 = min(kb1, (k_hi + PLANE_KB - 1) / PLANE_KB);
    }
}

__host__ __device__ __forceinline__ bool tile_live(int M, int N, int bn, int diag_m, int diag_n, int mt, int nt) {
    if (diag_m <= 0) return true;
    const int r_lo = mt * TC_BM, r_hi = min(M, r_lo + TC_BM) - 1, c_lo = nt * bn, c_hi = min(N, c_lo + bn) - 1;
    return r_lo / diag_m <= c_hi / diag_n && c_lo / diag_n <= r_hi / diag_m;
}

template <int BN>
__device__ __forceinline__ bool tc_item(const TcGemmParams& p, int it, TcItem& w) {
    const int idx = blockIdx.x + it * gridDim.x;
    if (idx < p.full_items) {
        w.ksp = idx % p.ksplit;
        const int t2 = idx / p.ksplit;
        w.mt = t2 / p.n_tiles; w.nt = t2 % p.n_tiles; w.tt = 0; w.slice = 0;
        w.mode = tile_live(p.M, p.N, BN, p.diag_m, p.diag_n, w.mt, w.nt) ? 0 : -1;
        tile_kb_range<BN>(p, w.ksp, w.nt, w.kb0, w.kb1);
        return true;
    }
    const int j = idx - p.full_items;
    if (j >= p.tail_tiles * p.tail_s) return false;
    w.tt = j / p.tail_s; w.slice = j % p.tail_s; w.ksp = 0;
    const int t2 = p.full_items + w.tt;                            // the tail split is only planned for ksplit == 1
    w.mt = t2 / p.n_tiles; w.nt = t2 % p.n_tiles;
    w.kb0 = w.slice * p.tail_per; w.kb1 = min(p.num_kb, w.kb0 + p.tail_per);
    w.mode = w.slice == 0 ? 1 : 2;
    return true;
}

// Cluster-pair form: the two CTAs of a cluster take the row tiles 2 mp and 2 mp + 1 of the SAME column tile and walk the same
// k-blocks, so the weight tile is fetched from L2 once per pair (each CTA loads one plane of it and multicasts it to both).
// mode 3: the odd row tile that does not exist (it still loads and multiplies, so that its peer's pipeline runs; no output).
template <int BN>
__device__ __forceinline__ bool tc_item_pair(const TcGemmParams& p, int it, int rank, TcItem& w) {
    const int s = (int)(blockIdx.x >> 1) + it * (int)(gridDim.x >> 1);
    const int m_pairs = (p.m_tiles + 1) >> 1;
    if (s >= m_pairs * p.n_tiles) return false;
    w.nt = s % p.n_tiles; w.mt = 2 * (s / p.n_tiles) + rank;
    w.ksp = 0; w.kb0 = 0; w.kb1 = p.num_kb; w.tt = 0; w.slice = 0;
    w.mode = w.mt < p.m_tiles ? 0 : 3;
    return true;
}

template <int BN, bool MN = false, bool CL = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const TcGemmParams p, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b) {
    using Cfg = TcCfg<BN>;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* stage_base = tc_smem;
    float* stg_base = reinterpret_cast<float*>(tc_smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base word
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * Cfg::STAGES;
    const uint32_t tfull0 = empty0 + 8 * Cfg::STAGES, tempty0 = tfull0 + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        // CL: a stage is refilled by both CTAs of the pair (own A planes, one multicast B plane each), so it is free only
        // when the MMAs of BOTH CTAs have retired from it
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, CL ? 2 : 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int crank = CL ? (int)cluster_ctarank() : 0;
    if (CL) cluster_sync_all();                              // the peer's barriers exist before anything is sent to them

    const bool use_lo = p.npairs > 1;
    const uint32_t stage_tx = use_lo ? Cfg::STAGE_BYTES : (TC_A_PLANE_BYTES + Cfg::B_PLANE_BYTES);

    if (warp == 0) {
        // ------------------------------------------------ producer
        int stage = 0; uint32_t phase = 0;
        const int64_t a_chunk = (int64_t)TC_BM * PLANE_KB, b_chunk = (int64_t)BN * PLANE_KB;   // elements
        TcItem w;
        for (int it = 0; CL ? tc_item_pair<BN>(p, it, crank, w) : tc_item<BN>(p, it, w); ++it) {
            if (w.mode < 0) continue;
            const int mt = w.mt, nt = w.nt, kb0 = w.kb0, kb1 = w.kb1;
            const __nv_bfloat16* a_src = p.A + ((int64_t)(w.mode == 3 ? p.m_tiles - 1 : mt) * p.a_nkb + p.a_kb_off) * a_chunk;
            const __nv_bfloat16* b_src = p.B + (int64_t)nt * p.num_kb * b_chunk;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (MN) {
                    // 32 rows of (128 | BN) features of both planes: one tensor-map tile per operand.  The row planes are
                    // [row tile][8-feature group][row 0..127][8]: the box takes, for each of the tile's groups, the 512-byte
                    // run of these 32 rows and lays them down group after group, hi plane then lo plane = the canonical
                    // MN-major operand (8 k = 128 B apart, 8 features = 512 B apart).
                    if (lane == 0) {
                        const uint32_t bar = full0 + 8 * stage;
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES);
                        const uint32_t sb = sa + 2 * TC_A_PLANE_BYTES;
                        mbar_expect_tx(bar, Cfg::STAGE_BYTES);
                        tma_load_4d(sa, &tm_a, (kb & 3) * 256, 16 * mt, kb >> 2, 0, bar);
                        tma_load_4d(sb, &tm_b, (kb & 3) * 256, (BN / 8) * nt, kb >> 2, 0, bar);
                    }
                } else if (CL && lane == 0) {
                    const uint32_t bar = full0 + 8 * stage;
                    const uint32_t sa = smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + 2 * TC_A_PLANE_BYTES;
                    mbar_expect_tx(bar, stage_tx);                 // own A planes + both B planes, whoever sends them
                    bulk_g2s(sa, a_src + kb * a_chunk, TC_A_PLANE_BYTES, bar);
                    if (use_lo) bulk_g2s(sa + TC_A_PLANE_BYTES, a_src + p.a_plane_stride + kb * a_chunk, TC_A_PLANE_BYTES, bar);
                    if (crank == 0)
                        bulk_g2s_multicast(sb, b_src + kb * b_chunk, Cfg::B_PLANE_BYTES, bar, (uint16_t)3);
                    else if (use_lo)
                        bulk_g2s_multicast(sb + Cfg::B_PLANE_BYTES, b_src + p.b_plane_stride + kb * b_chunk, Cfg::B_PLANE_BYTES, bar,
                                           (uint16_t)3);
                } else if (lane == 0) {
                    const uint32_t bar = full0 + 8 * stage;
                    const uint32_t sa = smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + 2 * TC_A_PLANE_BYTES;
                    mbar_expect_tx(bar, stage_tx);
                    bulk_g2s(sa, a_src + kb * a_chunk, TC_A_PLANE_BYTES, bar);
                    bulk_g2s(sb, b_src + kb * b_chunk, Cfg::B_PLANE_BYTES, bar);
                    if (use_lo) {
                        bulk_g2s(sa + TC_A_PLANE_BYTES, a_src + p.a_plane_stride + kb * a_chunk, TC_A_PLANE_BYTES, bar);
                        bulk_g2s(sb + Cfg::B_PLANE_BYTES, b_src + p.b_plane_stride + kb * b_chunk, Cfg::B_PLANE_BYTES, bar);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (CL) {
            // the peer's MMAs still arrive on this CTA's empty barriers: wait until every stage has been released once more
            for (int s = 0; s < Cfg::STAGES; ++s) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24) | (MN ? ((1u << 15) | (1u << 16)) : 0u);
        // K-major: LBO = between the two k-halves of a K = 16 step, SBO = between 8-row groups; MN-major: LBO = between 8-k
        // groups (128 B), SBO = between 8-feature groups (512 B), a K = 16 step = 256 B
        constexpr uint32_t LBO_A = MN ? 128 : TC_BM * 16, LBO_B = MN ? 128 : BN * 16, SBO = MN ? 512 : 128;
        constexpr uint32_t KS_A = MN ? 256 : 2 * LBO_A, KS_B = MN ? 256 : 2 * LBO_B;
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        TcItem w;
        for (int it = 0; CL ? tc_item_pair<BN>(p, it, crank, w) : tc_item<BN>(p, it, w); ++it) {
            if (w.mode < 0) continue;
            const int kb0 = w.kb0, kb1 = w.kb1;
            if (lane == 0) GDBG(1, it, 0);
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            if (lane == 0) GDBG(1, it, 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + 2 * TC_A_PLANE_BYTES;
                    // MN: the contraction runs over rows, whose planes are not padded with zeros: the last block stops at K
                    const int nks = MN ? min(PLANE_KB / 16, (p.K - kb * PLANE_KB + 15) / 16) : PLANE_KB / 16;
#pragma unroll
                    for (int ks = 0; ks < PLANE_KB / 16; ++ks) {
                        if (MN && ks >= nks) break;
                        const uint64_t a_hi = make_smem_desc(sa + ks * KS_A, LBO_A, SBO);
                        const uint64_t b_hi = make_smem_desc(sb + ks * KS_B, LBO_B, SBO);
                        if (use_lo) {
                            const uint64_t a_lo = make_smem_desc(sa + TC_A_PLANE_BYTES + ks * KS_A, LBO_A, SBO);
                            const uint64_t b_lo = make_smem_desc(sb + Cfg::B_PLANE_BYTES + ks * KS_B, LBO_B, SBO);
                            // small cross terms first, the dominant hi.hi term last
                            tc_mma_bf16(d_tmem, a_lo, b_hi, IDESC, (kb > kb0 || ks) ? 1u : 0u);
                            tc_mma_bf16(d_tmem, a_hi, b_lo, IDESC, 1u);
                            tc_mma_bf16(d_tmem, a_hi, b_hi, IDESC, 1u);
                        } else {
                            tc_mma_bf16(d_tmem, a_hi, b_hi, IDESC, (kb > kb0 || ks) ? 1u : 0u);
                        }
                    }
                    if (CL) tc_commit_multicast(empty0 + 8 * stage, (uint16_t)3);   // ... in both CTAs of the pair
                    else tc_commit(empty0 + 8 * stage);            // frees the smem slot when these MMAs retire
                    if (kb == kb1 - 1) tc_commit(tfull0 + 8 * acc);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) GDBG(1, it, 2);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------------------------ epilogue: 8 warps, warp % 4 = TMEM lane quadrant,
        // (warp - 2) / 4 = which half of the tile's columns.  All global loads of a chunk are issued before its stores.
        const int quad = warp & 3, half = (warp - 2) >> 2;
        float* stg = stg_base + (size_t)(warp - 2) * 32 * 33;
        int acc = 0; uint32_t acc_phase = 0;
        const int nkb_next = (int)plane_kblocks(p.N);
        const int kpad_next = nkb_next * PLANE_KB;
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;           // coalesced phase: 4 rows x 8 float4 per pass
        constexpr int CH = BN / 64;                                // 32-column chunks per half
        TcItem w;
        for (int it = 0; CL ? tc_item_pair<BN>(p, it, crank, w) : tc_item<BN>(p, it, w); ++it) {
            if (w.mode < 0) continue;
            const int mt = w.mt, nt = w.nt, ksp = w.ksp;
            float* const outp = p.out ? p.out + (int64_t)ksp * p.split_stride : nullptr;
            const int rr_own = quad * 32 + lane;                   // row of the tile this thread owns in TMEM
            const int64_t m_own = (int64_t)mt * TC_BM + rr_own;
            const int64_t m_base = (int64_t)mt * TC_BM + quad * 32;
            if (warp == 2 && lane == 0) GDBG(2, it, 0);
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            if (warp == 2 && lane == 0) GDBG(2, it, 1);
            tc_fence_after();
            // tail slices: this thread's row of the raw accumulators other slices hand over / this slice hands over
            // (stored in the register layout [warp][chunk][float4 j][lane]: every warp access is 512 contiguous bytes)
            float* const part_row = p.tail_part + ((int64_t)w.tt * (p.tail_s - 1) + (w.slice > 0 ? w.slice - 1 : 0)) * (TC_BM * BN) +
                                    (int64_t)(quad * 2 + half) * (CH * 1024) + lane * 4;
            if (w.mode == 1) {
                const unsigned int want = (unsigned int)(p.tail_s - 1) * 8u;
                unsigned int have;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(have) : "l"(p.tail_flags + 2 * w.tt) : "memory");
                    if (have < want) __nanosleep(64);
                } while (have < want);
            }
#pragma unroll 1
            for (int cc = 0; cc < CH; ++cc) {
                const int c = half * CH + cc;
                const int col0 = nt * BN + c * 32;
                if (w.mode == 3 || (col0 >= p.N && col0 >= kpad_next)) break;
                float v[32];
                tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                if (w.mode == 2) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        __stcg(reinterpret_cast<float4*>(part_row + cc * 1024 + j * 32), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    continue;
                }
                if (w.mode == 1) {
#pragma unroll 1
                    for (int sl = 1; sl < p.tail_s; ++sl) {
                        const float* src = part_row + (int64_t)(sl - 1) * TC_BM * BN + cc * 1024;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 q = __ldcg(reinterpret_cast<const float4*>(src + j * 32));
                            v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                        }
                    }
                }
                if (p.bias) {
                    if (col0 + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
                    }
                }
                if (p.preact) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
                    __syncwarp();
#pragma unroll
                    for (int rr = 0; rr < 32; rr += 4) {
                        const int row = rr + rsub;
                        const int64_t m = m_base + row;
                        const int col = col0 + c4;
                        if (m < p.M && col < p.N)
                            *reinterpret_cast<float4*>(p.preact + m * p.ldc + col) =
                                make_float4(stg[row * 33 + c4], stg[row * 33 + c4 + 1], stg[row * 33 + c4 + 2],
                                            stg[row * 33 + c4 + 3]);
                    }
                    __syncwarp();
                }
                if (p.gate_planes) {
                    // thread owns row m_own: its 32 gate values are four 16-byte units of the hi plane, 128 rows * 16 B apart
                    // (the 32 lanes read 512 contiguous bytes per unit); rows >= M and columns >= N are zeros there
                    if (col0 < kpad_next) {
                        const __nv_bfloat16* gsrc = p.gate_planes + (((int64_t)mt * nkb_next + (col0 >> 5)) * 4) * (TC_BM * 8) + rr_own * 8;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint4 q = __ldg(reinterpret_cast<const uint4*>(gsrc + u * (TC_BM * 8)));
                            const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int16_t bits = (int16_t)(wd[j >> 1] >> ((j & 1) * 16));
                                v[u * 8 + j] = bits > 0 ? v[u * 8 + j] * p.gate_scale : 0.f;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
                }
                if (p.gate) {
                    // activation backward in the epilogue: the gate rows are read in the coalesced layout of the output stores
                    // (4 rows x 128 B per pass), the gated values go back through the transpose buffer for the plane writer
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
                    __syncwarp();
                    const int col = col0 + c4;
                    const DrawKey gk = rng_resolve(p.seed, p.offset);
                    const DropFast gfast = drop_fast_setup(gk.seed, gk.offset, p.drop_p);
                    const bool gsmall = (int64_t)p.M * p.N < (1ll << 32);
                    float4 g4[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int64_t m = m_base + i * 4 + rsub;
                        g4[i] = (m < p.M && col < p.N) ? __ldg(reinterpret_cast<const float4*>(p.gate + m * p.ldg + col))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    // act'(gate) in place, one uniform branch per chunk (a per-element switch bloats the code: see the forward path)
                    if (p.gate_act == ACT_RELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            g4[i].x = g4[i].x > 0.f ? 1.f : 0.f; g4[i].y = g4[i].y > 0.f ? 1.f : 0.f;
                            g4[i].z = g4[i].z > 0.f ? 1.f : 0.f; g4[i].w = g4[i].w > 0.f ? 1.f : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            g4[i].x = act_grad_call(p.gate_act, g4[i].x); g4[i].y = act_grad_call(p.gate_act, g4[i].y);
                            g4[i].z = act_grad_call(p.gate_act, g4[i].z); g4[i].w = act_grad_call(p.gate_act, g4[i].w);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + rsub;
                        const int64_t m = m_base + row;
                        float o[4] = {stg[row * 33 + c4], stg[row * 33 + c4 + 1], stg[row * 33 + c4 + 2], stg[row * 33 + c4 + 3]};
                        const float gg[4] = {g4[i].x, g4[i].y, g4[i].z, g4[i].w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            o[jj] *= gg[jj];
                            if (p.drop_p > 0.f)
                                o[jj] *= gsmall ? (drop_fast_keep(gfast, (uint32_t)(m * p.N + col + jj)) ? gfast.scale : 0.f)
                                                : drop_keep_scale(gk.seed, gk.offset, (uint64_t)(m * p.N + col + jj), p.drop_p);
                        }
                        const bool live = m < p.M && col < p.N;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) stg[row * 33 + c4 + jj] = live ? o[jj] : 0.f;
                        if (outp && live) *reinterpret_cast<float4*>(outp + m * p.ldc + col) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    __syncwarp();
                    if (p.out_planes && col0 < kpad_next) {
                        __nv_bfloat16* dst = p.out_planes + (((int64_t)mt * nkb_next + (col0 >> 5)) * 4) * (TC_BM * 8) + rr_own * 8;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            bf16x8 h, l;
#pragma unroll
                            for (int j = 0; j < 8; ++j) split_bf16(stg[lane * 33 + u * 8 + j], h.v[j], l.v[j]);
                            *reinterpret_cast<bf16x8*>(dst + u * (TC_BM * 8)) = h;
                            *reinterpret_cast<bf16x8*>(dst + p.out_plane_stride + u * (TC_BM * 8)) = l;
                        }
                    }
                    if (p.colsum_part) {                  // lane = column: rows beyond M hold zeros in the buffer
                        float sum = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum += stg[i * 33 + lane];
                        if (col0 + lane < p.N) p.colsum_part[((int64_t)mt * 4 + quad) * p.N + col0 + lane] = sum;
                    }
                    __syncwarp();
                    continue;
                }
                switch (p.act) {          // hoisted: one uniform branch per chunk, straight-line code per activation
                    case ACT_RELU:
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                        break;
                    case ACT_NONE: break;
                    default:
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = act_apply(p.act, v[j]);
                }
                if (p.drop_p > 0.f) {
                    const DrawKey dk_ = rng_resolve(p.seed, p.offset);
                    if ((int64_t)p.M * p.N < (1ll << 32)) {          // 32-bit element index: the hoisted form of the same draw
                        const DropFast df = drop_fast_setup(dk_.seed, dk_.offset, p.drop_p);
                        const uint32_t e0 = (uint32_t)(m_own * p.N + col0);            // even: N % 4 == 0, col0 % 32 == 0
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {                              // one hash per pair of elements
                            const uint32_t x = drop_fast_hash(df, (e0 + j) >> 1);
                            v[j] = (x & 0xFFFFu) >= df.thr ? v[j] * df.scale : 0.f;
                            v[j + 1] = (x >> 16) >= df.thr ? v[j + 1] * df.scale : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] *= drop_keep_scale(dk_.seed, dk_.offset, (uint64_t)(m_own * p.N + col0 + j), p.drop_p);
                    }
                }
                if (p.out_planes) {
                    // thread owns row m_own and 32 consecutive k of the next GEMM = one chunk column block:
                    // kb = col0 / 32, four 16-byte units (kg = 0..3) that are 128 rows * 16 B apart.  Rows >= M are
                    // written as zeros so that consumers contracting over rows never meet non-finite padding.
                    if (col0 < kpad_next) {
                        const bool live = m_own < p.M;
                        __nv_bfloat16* dst = p.out_planes + (((int64_t)mt * nkb_next + (col0 >> 5)) * 4) * (TC_BM * 8) + rr_own * 8;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            bf16x8 h, l;
#pragma unroll
                            for (int j = 0; j < 8; ++j) split_bf16(live ? v[u * 8 + j] : 0.f, h.v[j], l.v[j]);
                            *reinterpret_cast<bf16x8*>(dst + u * (TC_BM * 8)) = h;
                            *reinterpret_cast<bf16x8*>(dst + p.out_plane_stride + u * (TC_BM * 8)) = l;
                        }
                    }
                }
                if (p.colsum_part) {                      // column sums of this warp's 32 x 32 block (rows >= M count as zeros)
                    const bool live = m_own < p.M;
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = live ? v[j] : 0.f;
                    __syncwarp();
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum += stg[i * 33 + lane];
                    if (col0 + lane < p.N) p.colsum_part[((int64_t)mt * 4 + quad) * p.N + col0 + lane] = sum;
                    __syncwarp();
                }
                if (p.out) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
                    __syncwarp();
                    const int col = col0 + c4;
                    float4 r4[8];
                    if (p.resid) {
                        const float* rrow[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int64_t m = m_base + i * 4 + rsub;
                            rrow[i] = p.resid + (m < p.M ? m : 0) * p.ldr;
                            if (p.row_map && m < p.M) {
                                const int32_t slot = __ldg(p.row_map + m);
                                if (slot >= 0) rrow[i] = p.resid_alt + (int64_t)slot * p.ldr;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            r4[i] = (col < p.N) ? __ldg(reinterpret_cast<const float4*>(rrow[i] + col))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + rsub;
                        const int64_t m = m_base + row;
                        if (m < p.M && col < p.N) {
                            float4 o = make_float4(stg[row * 33 + c4], stg[row * 33 + c4 + 1], stg[row * 33 + c4 + 2],
                                                   stg[row * 33 + c4 + 3]);
                            if (p.resid) { o.x += r4[i].x; o.y += r4[i].y; o.z += r4[i].z; o.w += r4[i].w; }
                            *reinterpret_cast<float4*>(outp + m * p.ldc + col) = o;
                        }
                    }
                    __syncwarp();
                }
            }
            if (w.mode == 2) {                         // hand-over complete: count this warp in (release)
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(p.tail_flags + 2 * w.tt, 1u);
            } else if (w.mode == 1) {                  // the last of the owner's warps re-arms the flags for the next launch
                __syncwarp();
                if (lane == 0 && atomicAdd(p.tail_flags + 2 * w.tt + 1, 1u) == 7u) {
                    p.tail_flags[2 * w.tt] = 0u;
                    p.tail_flags[2 * w.tt + 1] = 0u;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (warp == 2 && lane == 0) GDBG(2, it, 2);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();                              // neither CTA leaves while the other may still send to it
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// Rows-per-chunk (= BLOCK_N of the kernel that will consume them) a [N, K] weight should be tiled with.
int snuffy_gemm_tc_block_n(int64_t N) {
    const int64_t n128 = (N + 127) / 128;
    return (n128 % 2 == 0) ? 256 : 128;
}

// bf16 elements in ONE plane of a [rows, K] operand tiled with rc rows per chunk
int64_t snuffy_plane_elems(int64_t rows, int64_t K, int rc) { return plane_elems(rows, K, rc); }

// C = epilogue(A . B^T) with A, B given as split-bf16 planes (A tiled with 128 rows per chunk, B with
// snuffy_gemm_tc_block_n(N)).  Outputs (each optional, at least one): fp32 `out` [M, ldc] (+ residual, optionally
// redirected through row_map), fp32 `preact`, and `out_planes` = the activated result as A planes with K_next = N.
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<TensorMapEncodeTiledFn>(sym);
    }();
    return fn;
}

// Row planes of an [R, F] activation (128 rows per chunk) as a 4-D tensor: (element of a group's 128-row run, 8-feature group,
// row tile, hi / lo plane); one box = 32 rows x `groups` groups x both planes.
static int row_planes_tensor_map(CUtensorMap* tm, const void* planes, int64_t plane_stride, int64_t R, int64_t F, int groups) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    SNUFFY_REQUIRE(enc, "snuffy_gemm_tc_splitk_rows: cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t ngroups = (cuuint64_t)plane_kblocks(F) * 4;
    const cuuint64_t dims[4] = {(cuuint64_t)TC_BM * 8, ngroups, (cuuint64_t)plane_rtiles(R, TC_BM), 2};
    const cuuint64_t strides[3] = {(cuuint64_t)TC_BM * 16, ngroups * TC_BM * 16, (cuuint64_t)plane_stride * 2};   // bytes
    const cuuint32_t box[4] = {256, (cuuint32_t)groups, 1, 2};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(planes), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("snuffy_gemm_tc_splitk_rows: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return 1;
    }
    return 0;
}

// Hand-over buffers of the tail split: a few per device, each claimed by the first stream that needs one (launches on one
// stream are ordered, so a buffer is never shared by two running kernels).  Allocated on the first launch outside a stream
// capture; a launch that finds none (all claimed, or first use inside a capture) simply runs without the tail split.
struct TailWorkspace { int device; cudaStream_t stream; bool claimed; float* part; unsigned int* flags; };
constexpr int TAIL_WS_PER_DEVICE = 4, TAIL_WS_MAX = 64;
constexpr size_t TAIL_PART_BYTES = (size_t)160 * TC_BM * 256 * sizeof(float);      // >= SM count partial tiles
static TailWorkspace g_tail_ws[TAIL_WS_MAX];
static int g_tail_ws_n = 0;
static std::mutex g_tail_ws_mutex;

static TailWorkspace* tail_workspace(cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_tail_ws_mutex);
    TailWorkspace* free_slot = nullptr;
    bool have_device = false;
    for (int i = 0; i < g_tail_ws_n; ++i) {
        TailWorkspace& w = g_tail_ws[i];
        if (w.device != dev) continue;
        have_device = true;
        if (w.claimed && w.stream == stream) return &w;
        if (!w.claimed && !free_slot) free_slot = &w;
    }
    if (!have_device) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
        if (g_tail_ws_n + TAIL_WS_PER_DEVICE > TAIL_WS_MAX) return nullptr;
        for (int i = 0; i < TAIL_WS_PER_DEVICE; ++i) {
            TailWorkspace w{dev, nullptr, false, nullptr, nullptr};
            if (cudaMalloc(&w.part, TAIL_PART_BYTES) != cudaSuccess || cudaMalloc(&w.flags, 4096) != cudaSuccess ||
                cudaMemset(w.flags, 0, 4096) != cudaSuccess) {
                cudaGetLastError();
                return free_slot;
            }
            g_tail_ws[g_tail_ws_n++] = w;
            if (!free_slot) free_slot = &g_tail_ws[g_tail_ws_n - 1];
        }
        cudaDeviceSynchronize();                           // the flags are zero before any stream uses them
    }
    if (free_slot) { free_slot->claimed = true; free_slot->stream = stream; }
    return free_slot;
}

static int launch_gemm_tc(TcGemmParams& p, int64_t N, cudaStream_t stream, const char* who, int force_bn = 0,
                          const CUtensorMap* tm_a = nullptr, const CUtensorMap* tm_b = nullptr) {
    const int bn = force_bn ? force_bn : snuffy_gemm_tc_block_n(N);
    const int total = p.m_tiles * p.n_tiles * p.ksplit;
    int grid = total < sm_count() ? total : sm_count();
    static const CUtensorMap none{};
    const bool mn = tm_a != nullptr;
    p.full_items = total; p.tail_tiles = 0; p.tail_s = 1; p.tail_per = p.num_kb; p.tail_part = nullptr; p.tail_flags = nullptr;
    static const bool tail_split = [] { const char* e = getenv("SNUFFY_B200_TAIL_SPLIT"); return !e || atoi(e) != 0; }();
    const int sm = sm_count(), rem = total % sm;
    if (tail_split && !mn && p.ksplit == 1 && p.diag_m == 0 && p.group_n == 0 && rem > 0 && sm <= 160) {
        // Measured (tools/time_gemm_shapes.py, 10 000 rows): a hand-over costs ~6 us end to end (128 KB out through L2, fence,
        // flag, 128 KB back in, before the owner's epilogue can start), so the split only pays when the tile it shortens is
        // long: K >= 1536 (FFN-down and its dX twin: 73 -> 63 us); at K = 512 / 1024 it loses 1-5 us and is not planned.
        int sl = p.num_kb >= 48 ? 4 : 1;
        if (sl > sm / rem) sl = sm / rem;
        if (sl > p.num_kb / 2) sl = p.num_kb / 2;
        if (sl >= 2) {
            if (TailWorkspace* ws = tail_workspace(stream)) {
                const int per = (p.num_kb + sl - 1) / sl;
                sl = (p.num_kb + per - 1) / per;                           // every slice owns at least one k-block
                p.full_items = total - rem; p.tail_tiles = rem; p.tail_s = sl; p.tail_per = per;
                p.tail_part = ws->part; p.tail_flags = ws->flags;
                grid = p.full_items > 0 ? sm : rem * sl;
            }
        }
    }
    // Large plain products with a short contraction: CTA pairs (2-CTA clusters) that share each weight tile through TMA
    // multicast: a pair fetches the B planes from L2 once (32 instead of 48 KB per CTA and k-block).  Measured (160 000 rows,
    // tools/time_gemm_shapes.py): FFN-up 913 -> 874 us, Q|V 459 -> 445 us at K = 512; nothing at K = 2048 and -3.6 % at K = 1024
    // (the lockstep of the pair costs more than the feed saves once the MMAs are bound by their shared-memory operand reads),
    // so it is used for K <= 512 only.
    static const bool pairs = [] { const char* e = getenv("SNUFFY_B200_GEMM_PAIRS"); return !e || atoi(e) != 0; }();
    if (pairs && !mn && p.tail_tiles == 0 && p.ksplit == 1 && p.diag_m == 0 && p.group_n == 0 && p.m_tiles >= 2 && total >= sm &&
        sm % 2 == 0 && p.num_kb <= 16) {
        const int supers = ((p.m_tiles + 1) / 2) * p.n_tiles;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * (supers < sm / 2 ? supers : sm / 2)));
        cfg.blockDim = dim3(TC_THREADS);
        cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        // the persistent loop strides by the number of clusters: size the grid to what is co-resident (a GPC with an odd number
        // of free SMs leaves one out), or a late cluster would run its whole share after the others have finished
        static int resident[2] = {0, 0};                   // [bn == 256]
        int& res = resident[bn == 256 ? 1 : 0];
        if (bn == 256) {
            SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<256, false, true>), (int)TcCfg<256>::SMEM_BYTES));
            cfg.dynamicSmemBytes = TcCfg<256>::SMEM_BYTES;
            if (res == 0) SNUFFY_CUDA(cudaOccupancyMaxActiveClusters(&res, gemm_tc_kernel<256, false, true>, &cfg));
        } else {
            SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<128, false, true>), (int)TcCfg<128>::SMEM_BYTES));
            cfg.dynamicSmemBytes = TcCfg<128>::SMEM_BYTES;
            if (res == 0) SNUFFY_CUDA(cudaOccupancyMaxActiveClusters(&res, gemm_tc_kernel<128, false, true>, &cfg));
        }
        if (res >= 8) {
            const int nclusters = supers < res ? supers : res;
            cfg.gridDim = dim3((unsigned)(2 * nclusters));
            if (bn == 256) SNUFFY_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<256, false, true>, p, none, none));
            else SNUFFY_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<128, false, true>, p, none, none));
            return check_launch(who);
        }
    }
    if (mn && bn == 256) {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<256, true>), (int)TcCfg<256>::SMEM_BYTES));
        gemm_tc_kernel<256, true><<<grid, TC_THREADS, TcCfg<256>::SMEM_BYTES, stream>>>(p, *tm_a, *tm_b);
    } else if (mn) {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<128, true>), (int)TcCfg<128>::SMEM_BYTES));
        gemm_tc_kernel<128, true><<<grid, TC_THREADS, TcCfg<128>::SMEM_BYTES, stream>>>(p, *tm_a, *tm_b);
    } else if (bn == 256) {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<256>), (int)TcCfg<256>::SMEM_BYTES));
        gemm_tc_kernel<256><<<grid, TC_THREADS, TcCfg<256>::SMEM_BYTES, stream>>>(p, none, none);
    } else {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<128>), (int)TcCfg<128>::SMEM_BYTES));
        gemm_tc_kernel<128><<<grid, TC_THREADS, TcCfg<128>::SMEM_BYTES, stream>>>(p, none, none);
    }
    return check_launch(who);
}

static int gemm_tc_full(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                        int64_t M, int64_t N, int64_t K, int passes, const float* bias, int act, const float* resid,
                        int64_t ldr, const int32_t* row_map, const float* resid_alt, float* out, int64_t ldc,
                        float* preact, void* out_planes, int64_t out_plane_stride, float dropout_p, uint64_t seed,
                        uint64_t offset, const float* gate, int64_t ldg, int gate_act, cudaStream_t stream,
                        float* colsum_part = nullptr, const void* gate_planes = nullptr, float gate_scale = 1.f) {
    SNUFFY_REQUIRE(!gate || (ldg % 4 == 0 && ldg >= N && (uintptr_t)gate % 16 == 0),
                   "snuffy_gemm_tc_actgrad: the gate matrix must be 16-byte aligned with ldg %% 4 == 0, ldg >= N");
    SNUFFY_REQUIRE(A_planes && B_planes, "snuffy_gemm_tc: null operand");
    SNUFFY_REQUIRE(out || out_planes || preact || colsum_part, "snuffy_gemm_tc: no output requested");
    SNUFFY_REQUIRE(M >= 1 && N >= 1 && K >= 1, "snuffy_gemm_tc: empty problem");
    SNUFFY_REQUIRE(passes == 1 || passes == 3, "snuffy_gemm_tc: passes must be 1 or 3");
    SNUFFY_REQUIRE(N % 4 == 0 && (!out || (ldc % 4 == 0 && (uintptr_t)out % 16 == 0)) &&
                       (!preact || (ldc % 4 == 0 && (uintptr_t)preact % 16 == 0)) &&
                       (!resid || (ldr % 4 == 0 && (uintptr_t)resid % 16 == 0)),
                   "snuffy_gemm_tc: N, ldc, ldr must be multiples of 4 and fp32 pointers 16-byte aligned");
    SNUFFY_REQUIRE(!row_map || (resid && resid_alt), "snuffy_gemm_tc: row_map needs resid and resid_alt");
    SNUFFY_REQUIRE((uintptr_t)A_planes % 16 == 0 && (uintptr_t)B_planes % 16 == 0 && a_plane_stride % 8 == 0 &&
                       b_plane_stride % 8 == 0, "snuffy_gemm_tc: planes must be 16-byte aligned");
    TcGemmParams p{};
    p.A = reinterpret_cast<const __nv_bfloat16*>(A_planes); p.a_plane_stride = a_plane_stride;
    p.B = reinterpret_cast<const __nv_bfloat16*>(B_planes); p.b_plane_stride = b_plane_stride;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    const int bn = snuffy_gemm_tc_block_n(N);
    p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    p.n_tiles = (int)((N + bn - 1) / bn);
    p.num_kb = (int)plane_kblocks(K);
    p.npairs = passes;
    p.bias = bias; p.act = act; p.resid = resid; p.ldr = ldr; p.row_map = row_map; p.resid_alt = resid_alt;
    p.out = out; p.ldc = ldc; p.preact = preact;
    p.out_planes = reinterpret_cast<__nv_bfloat16*>(out_planes); p.out_plane_stride = out_plane_stride;
    p.drop_p = dropout_p; p.seed = seed; p.offset = offset;
    p.ksplit = 1; p.kb_per = p.num_kb; p.split_stride = 0;
    p.a_nkb = p.num_kb; p.a_kb_off = 0;
    p.gate = gate; p.ldg = ldg; p.gate_act = gate_act; p.colsum_part = colsum_part;
    p.gate_planes = reinterpret_cast<const __nv_bfloat16*>(gate_planes); p.gate_scale = gate_scale;
    return launch_gemm_tc(p, N, stream, gate ? "snuffy_gemm_tc_actgrad" : "snuffy_gemm_tc");
}

int snuffy_gemm_tc(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                   int64_t M, int64_t N, int64_t K, int passes, const float* bias, int act, const float* resid,
                   int64_t ldr, const int32_t* row_map, const float* resid_alt, float* out, int64_t ldc,
                   float* preact, void* out_planes, int64_t out_plane_stride, float dropout_p, uint64_t seed,
                   uint64_t offset, cudaStream_t stream) {
    return gemm_tc_full(A_planes, a_plane_stride, B_planes, b_plane_stride, M, N, K, passes, bias, act, resid, ldr, row_map,
                        resid_alt, out, ldc, preact, out_planes, out_plane_stride, dropout_p, seed, offset, nullptr, 0, 0, stream);
}

// dX product with the activation backward in its epilogue:  result = (A . B^T) * act'(gate[m, n]) * dropout_mask(m*N + n),
// as fp32 (out) and / or as A-operand planes for the next product.  For dh = (dY W2) * act'(h_pre) (autograd of snuffy.py:225).
int snuffy_gemm_tc_actgrad(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                           int64_t M, int64_t N, int64_t K, int passes, const float* gate, int64_t ldg, int gate_act,
                           float dropout_p, uint64_t seed, uint64_t offset, float* out, int64_t ldc, void* out_planes,
                           int64_t out_plane_stride, float* colsum, float* colsum_partials, cudaStream_t stream) {
    SNUFFY_REQUIRE(gate, "snuffy_gemm_tc_actgrad: null gate");
    SNUFFY_REQUIRE(!colsum || colsum_partials, "snuffy_gemm_tc_actgrad: column sums need their partials buffer");
    if (int rc = gemm_tc_full(A_planes, a_plane_stride, B_planes, b_plane_stride, M, N, K, passes, nullptr, ACT_NONE, nullptr, 0,
                              nullptr, nullptr, out, ldc, nullptr, out_planes, out_plane_stride, dropout_p, seed, offset, gate,
                              ldg, gate_act, stream, colsum ? colsum_partials : nullptr))
        return rc;
    if (colsum) {
        const int64_t parts = ((M + TC_BM - 1) / TC_BM) * 4;
        fold_wide_kernel<float><<<(unsigned)((N + 15) / 16), 256, 0, stream>>>(colsum_partials, (int)parts, N, colsum);
        return check_launch("snuffy_gemm_tc_actgrad");
    }
    return 0;
}

// out[M, N] (fp32, ldc) = A_window . B^T where A is the K window [a_col0, a_col0 + K) of a wider A-plane set over
// [rows, a_cols_total] columns (a_col0, K multiples of 32).  Used by the attention backward to contract the Q / V halves of the
// Q|V planes the forward already wrote, without re-splitting them.
int snuffy_gemm_tc_awindow(const void* A_planes, int64_t a_plane_stride, int64_t a_cols_total, int64_t a_col0,
                           const void* B_planes, int64_t b_plane_stride, int64_t M, int64_t N, int64_t K, int passes,
                           float* out, int64_t ldc, cudaStream_t stream) {
    SNUFFY_REQUIRE(A_planes && B_planes && out, "snuffy_gemm_tc_awindow: null pointer");
    SNUFFY_REQUIRE(M >= 1 && N >= 1 && K >= 1 && N % 4 == 0 && ldc % 4 == 0 && (uintptr_t)out % 16 == 0,
                   "snuffy_gemm_tc_awindow: bad problem");
    SNUFFY_REQUIRE(passes == 1 || passes == 3, "snuffy_gemm_tc_awindow: passes must be 1 or 3");
    SNUFFY_REQUIRE(a_col0 % 32 == 0 && K % 32 == 0 && a_col0 + K <= plane_kblocks(a_cols_total) * 32,
                   "snuffy_gemm_tc_awindow: the K window must be 32-aligned inside the planes");
    TcGemmParams p{};
    p.A = reinterpret_cast<const __nv_bfloat16*>(A_planes); p.a_plane_stride = a_plane_stride;
    p.B = reinterpret_cast<const __nv_bfloat16*>(B_planes); p.b_plane_stride = b_plane_stride;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    const int bn = snuffy_gemm_tc_block_n(N);
    p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    p.n_tiles = (int)((N + bn - 1) / bn);
    p.num_kb = (int)plane_kblocks(K);
    p.npairs = passes;
    p.act = ACT_NONE;
    p.out = out; p.ldc = ldc;
    p.ksplit = 1; p.kb_per = p.num_kb; p.split_stride = 0;
    p.a_nkb = (int)plane_kblocks(a_cols_total); p.a_kb_off = (int)(a_col0 / 32);
    return launch_gemm_tc(p, N, stream, "snuffy_gemm_tc_awindow");
}

// out[M, N] = A_window . B^T against a BLOCK-DIAGONAL B: B[n, k] != 0 only where n / group_n == k / group_k (the head-block
// operands Kbd / dObd of the attention backward and their transposes).  Each column tile contracts only over the k-blocks of the
// groups it touches.  b_rc = rows per chunk of the B planes (128 or 256) = the column tile width used.
int snuffy_gemm_tc_blockdiag(const void* A_planes, int64_t a_plane_stride, int64_t a_cols_total, int64_t a_col0,
                             const void* B_planes, int64_t b_plane_stride, int b_rc, int64_t M, int64_t N, int64_t K,
                             int passes, int64_t group_n, int64_t group_k, float* out, int64_t ldc, cudaStream_t stream) {
    SNUFFY_REQUIRE(A_planes && B_planes && out, "snuffy_gemm_tc_blockdiag: null pointer");
    SNUFFY_REQUIRE(M >= 1 && N >= 1 && K >= 1 && N % 4 == 0 && ldc % 4 == 0 && (uintptr_t)out % 16 == 0,
                   "snuffy_gemm_tc_blockdiag: bad problem");
    SNUFFY_REQUIRE(passes == 1 || passes == 3, "snuffy_gemm_tc_blockdiag: passes must be 1 or 3");
    SNUFFY_REQUIRE(b_rc == 128 || b_rc == 256, "snuffy_gemm_tc_blockdiag: b_rc must be 128 or 256");
    SNUFFY_REQUIRE(group_n >= 1 && group_k >= 1 && (N + group_n - 1) / group_n == (K + group_k - 1) / group_k,
                   "snuffy_gemm_tc_blockdiag: N / group_n and K / group_k must give the same number of groups");
    SNUFFY_REQUIRE(a_col0 % 32 == 0 && (a_col0 == 0 || K % 32 == 0) && a_col0 + K <= plane_kblocks(a_cols_total) * 32,
                   "snuffy_gemm_tc_blockdiag: the K window must be 32-aligned inside the planes");
    TcGemmParams p{};
    p.A = reinterpret_cast<const __nv_bfloat16*>(A_planes); p.a_plane_stride = a_plane_stride;
    p.B = reinterpret_cast<const __nv_bfloat16*>(B_planes); p.b_plane_stride = b_plane_stride;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    p.n_tiles = (int)((N + b_rc - 1) / b_rc);
    p.num_kb = (int)plane_kblocks(K);
    p.npairs = passes;
    p.act = ACT_NONE;
    p.out = out; p.ldc = ldc;
    p.ksplit = 1; p.kb_per = p.num_kb; p.split_stride = 0;
    p.a_nkb = (int)plane_kblocks(a_cols_total); p.a_kb_off = (int)(a_col0 / 32);
    p.group_n = (int)group_n; p.group_k = (int)group_k;
    return launch_gemm_tc(p, N, stream, "snuffy_gemm_tc_blockdiag", b_rc);
}

// The same product for ReLU, gated by the forward's own activated planes: result = (A . B^T) * (a[m, n] > 0 ? 1 / (1 - p) : 0)
// with a = dropout(relu(h)) as written by snuffy_gemm_tc's out_planes (only the hi plane is read).  Neither the fp32
// pre-activation nor the dropout draw is needed: a clamped or dropped element is exactly 0 in the planes.
int snuffy_gemm_tc_relugrad(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                            int64_t M, int64_t N, int64_t K, int passes, const void* act_planes, float dropout_p, float* out,
                            int64_t ldc, void* out_planes, int64_t out_plane_stride, float* colsum, float* colsum_partials,
                            cudaStream_t stream) {
    SNUFFY_REQUIRE(act_planes && (uintptr_t)act_planes % 16 == 0, "snuffy_gemm_tc_relugrad: null or misaligned activation planes");
    SNUFFY_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "snuffy_gemm_tc_relugrad: dropout_p out of range");
    SNUFFY_REQUIRE(!colsum || colsum_partials, "snuffy_gemm_tc_relugrad: column sums need their partials buffer");
    if (int rc = gemm_tc_full(A_planes, a_plane_stride, B_planes, b_plane_stride, M, N, K, passes, nullptr, ACT_NONE, nullptr, 0,
                              nullptr, nullptr, out, ldc, nullptr, out_planes, out_plane_stride, 0.f, 0, 0, nullptr, 0, 0, stream,
                              colsum ? colsum_partials : nullptr, act_planes, 1.f / (1.f - dropout_p)))
        return rc;
    if (colsum) {
        const int64_t parts = ((M + TC_BM - 1) / TC_BM) * 4;
        fold_wide_kernel<float><<<(unsigned)((N + 15) / 16), 256, 0, stream>>>(colsum_partials, (int)parts, N, colsum);
        return check_launch("snuffy_gemm_tc_relugrad");
    }
    return 0;
}

// Split-K form for the weight gradients dW[M, N] = A . B^T with a long contraction (K = all patches of the step) and few
// output tiles: `ksplit` <= 0 picks one that fills the machine.  partials: snuffy_gemm_tc_splitk_workspace floats * 4 bytes.
int64_t snuffy_gemm_tc_auto_ksplit(int64_t M, int64_t N, int64_t K) {
    const int bn = snuffy_gemm_tc_block_n(N);
    const int64_t tiles = ((M + TC_BM - 1) / TC_BM) * ((N + bn - 1) / bn);
    const int64_t num_kb = plane_kblocks(K);
    int64_t ks = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
    if (ks > num_kb / 8) ks = num_kb / 8;                  // at least 8 k-blocks (256 k) per split
    if (ks < 1) ks = 1;
    const int64_t per = (num_kb + ks - 1) / ks;
    return (num_kb + per - 1) / per;                       // every split owns at least one k-block
}
int64_t snuffy_gemm_tc_splitk_workspace(int64_t M, int64_t N, int64_t ksplit) { return ksplit > 1 ? M * N * ksplit * 4 : 0; }

// K split for the block-diagonal-output form: fills the machine with the LIVE tiles only.
int64_t snuffy_gemm_tc_diag_ksplit(int64_t M, int64_t N, int64_t K, int b_rc, int64_t diag_m, int64_t diag_n) {
    if (b_rc != 128 && b_rc != 256) return 1;
    const int m_tiles = (int)((M + TC_BM - 1) / TC_BM), n_tiles = (int)((N + b_rc - 1) / b_rc);
    int64_t live = 0;
    for (int mt = 0; mt < m_tiles; ++mt)
        for (int nt = 0; nt < n_tiles; ++nt) live += tile_live((int)M, (int)N, b_rc, (int)diag_m, (int)diag_n, mt, nt) ? 1 : 0;
    if (live < 1) live = 1;
    const int64_t num_kb = plane_kblocks(K);
    int64_t ks = (2 * (int64_t)sm_count() + live - 1) / live;
    if (ks > num_kb / 8) ks = num_kb / 8;
    if (ks < 1) ks = 1;
    const int64_t per = (num_kb + ks - 1) / ks;
    return (num_kb + per - 1) / per;
}

static int gemm_tc_splitk_full(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                               int b_rc, int64_t M, int64_t N, int64_t K, int passes, int64_t ksplit, int64_t diag_m,
                               int64_t diag_n, float* out, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    SNUFFY_REQUIRE(A_planes && B_planes && out, "snuffy_gemm_tc_splitk: null pointer");
    SNUFFY_REQUIRE(M >= 1 && N >= 1 && K >= 1 && N % 4 == 0 && (uintptr_t)out % 16 == 0, "snuffy_gemm_tc_splitk: bad problem");
    SNUFFY_REQUIRE(passes == 1 || passes == 3, "snuffy_gemm_tc_splitk: passes must be 1 or 3");
    SNUFFY_REQUIRE((diag_m == 0 && diag_n == 0) || (diag_m >= 1 && diag_n >= 1 && (b_rc == 128 || b_rc == 256)),
                   "snuffy_gemm_tc_splitk_blockdiag: bad block sizes");
    if (ksplit <= 0) ksplit = diag_m ? snuffy_gemm_tc_diag_ksplit(M, N, K, b_rc, diag_m, diag_n) : snuffy_gemm_tc_auto_ksplit(M, N, K);
    const int64_t num_kb = plane_kblocks(K);
    const int64_t per = (num_kb + ksplit - 1) / ksplit;
    ksplit = (num_kb + per - 1) / per;
    SNUFFY_REQUIRE(ksplit == 1 || (workspace && workspace_bytes >= snuffy_gemm_tc_splitk_workspace(M, N, ksplit) &&
                                   (uintptr_t)workspace % 16 == 0), "snuffy_gemm_tc_splitk: workspace too small");
    TcGemmParams p{};
    p.A = reinterpret_cast<const __nv_bfloat16*>(A_planes); p.a_plane_stride = a_plane_stride;
    p.B = reinterpret_cast<const __nv_bfloat16*>(B_planes); p.b_plane_stride = b_plane_stride;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    const int bn = diag_m ? b_rc : snuffy_gemm_tc_block_n(N);
    p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    p.n_tiles = (int)((N + bn - 1) / bn);
    p.num_kb = (int)num_kb;
    p.npairs = passes;
    p.act = ACT_NONE;
    p.out = ksplit > 1 ? reinterpret_cast<float*>(workspace) : out; p.ldc = N;
    p.ksplit = (int)ksplit; p.kb_per = (int)per; p.split_stride = M * N;
    p.a_nkb = p.num_kb; p.a_kb_off = 0;
    p.diag_m = (int)diag_m; p.diag_n = (int)diag_n;
    if (int rc = launch_gemm_tc(p, N, stream, "snuffy_gemm_tc_splitk", diag_m ? bn : 0)) return rc;
    if (ksplit > 1) {
        launch_fold_partials(reinterpret_cast<const float*>(workspace), (int)ksplit, M * N / 4, out, stream);
        return check_launch("snuffy_gemm_tc_splitk");
    }
    return 0;
}

int snuffy_gemm_tc_splitk(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                          int64_t M, int64_t N, int64_t K, int passes, int64_t ksplit, float* out, void* workspace,
                          int64_t workspace_bytes, cudaStream_t stream) {
    return gemm_tc_splitk_full(A_planes, a_plane_stride, B_planes, b_plane_stride, 0, M, N, K, passes, ksplit, 0, 0, out, workspace,
                               workspace_bytes, stream);
}

// dW[M, N] = dY^T X straight from the ROW planes of dY [R, M] and X [R, N] (128 rows per chunk, as the GEMM epilogues and
// snuffy_ln_rows_fwd write them): the contraction runs over the rows, read through MN-major descriptors, so no transposed
// copy of either operand is made.  M % 128 == 0, N % snuffy_gemm_tc_block_n(N) == 0; rows R .. ceil16(R) of both plane sets
// must be finite (zero when R % 16 != 0: snuffy_planes_zero_rows).  Workspace as snuffy_gemm_tc_splitk(M, N, R).
int snuffy_gemm_tc_splitk_rows(const void* dY_planes, int64_t a_plane_stride, const void* X_planes, int64_t b_plane_stride,
                               int64_t M, int64_t N, int64_t R, int passes, int64_t ksplit, float* out, void* workspace,
                               int64_t workspace_bytes, cudaStream_t stream) {
    SNUFFY_REQUIRE(dY_planes && X_planes && out, "snuffy_gemm_tc_splitk_rows: null pointer");
    const int bn = snuffy_gemm_tc_block_n(N);
    SNUFFY_REQUIRE(M >= 128 && N >= 128 && R >= 1 && M % TC_BM == 0 && N % bn == 0 && (uintptr_t)out % 16 == 0,
                   "snuffy_gemm_tc_splitk_rows: M must be a multiple of 128 and N of its column tile");
    SNUFFY_REQUIRE(passes == 1 || passes == 3, "snuffy_gemm_tc_splitk_rows: passes must be 1 or 3");
    if (ksplit <= 0) ksplit = snuffy_gemm_tc_auto_ksplit(M, N, R);
    const int64_t num_kb = plane_kblocks(R);
    const int64_t per = (num_kb + ksplit - 1) / ksplit;
    ksplit = (num_kb + per - 1) / per;
    SNUFFY_REQUIRE(ksplit == 1 || (workspace && workspace_bytes >= snuffy_gemm_tc_splitk_workspace(M, N, ksplit) &&
                                   (uintptr_t)workspace % 16 == 0), "snuffy_gemm_tc_splitk_rows: workspace too small");
    TcGemmParams p{};
    p.A = reinterpret_cast<const __nv_bfloat16*>(dY_planes); p.a_plane_stride = a_plane_stride;
    p.B = reinterpret_cast<const __nv_bfloat16*>(X_planes); p.b_plane_stride = b_plane_stride;
    p.M = (int)M; p.N = (int)N; p.K = (int)R;
    p.m_tiles = (int)(M / TC_BM);
    p.n_tiles = (int)(N / bn);
    p.num_kb = (int)num_kb;
    p.npairs = passes;
    p.act = ACT_NONE;
    p.out = ksplit > 1 ? reinterpret_cast<float*>(workspace) : out; p.ldc = N;
    p.ksplit = (int)ksplit; p.kb_per = (int)per; p.split_stride = M * N;
    p.a_nkb = (int)plane_kblocks(M);
    CUtensorMap tm_a, tm_b;
    if (int rc = row_planes_tensor_map(&tm_a, dY_planes, a_plane_stride, R, M, 16)) return rc;
    if (int rc = row_planes_tensor_map(&tm_b, X_planes, b_plane_stride, R, N, bn / 8)) return rc;
    if (int rc = launch_gemm_tc(p, N, stream, "snuffy_gemm_tc_splitk_rows", 0, &tm_a, &tm_b)) return rc;
    if (ksplit > 1) {
        launch_fold_partials(reinterpret_cast<const float*>(workspace), (int)ksplit, M * N / 4, out, stream);
        return check_launch("snuffy_gemm_tc_splitk_rows");
    }
    return 0;
}

// Split-K product of which only the diagonal blocks (row / diag_m == col / diag_n) are wanted: tiles that meet no such block are
// skipped and their part of `out` is UNSPECIFIED.  b_rc = rows per chunk of the B planes = column tile width (128 or 256).
// dKp of the attention backward = the diagonal blocks of dS^T Q (M = h*Ksel, N = d, diag_m = Ksel, diag_n = d / h).
int snuffy_gemm_tc_splitk_blockdiag(const void* A_planes, int64_t a_plane_stride, const void* B_planes, int64_t b_plane_stride,
                                    int b_rc, int64_t M, int64_t N, int64_t K, int passes, int64_t ksplit, int64_t diag_m,
                                    int64_t diag_n, float* out, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    SNUFFY_REQUIRE(diag_m >= 1 && diag_n >= 1, "snuffy_gemm_tc_splitk_blockdiag: block sizes must be positive");
    return gemm_tc_splitk_full(A_planes, a_plane_stride, B_planes, b_plane_stride, b_rc, M, N, K, passes, ksplit, diag_m, diag_n, out,
                               workspace, workspace_bytes, stream);
}

#ifdef GEMM_DEBUG_TIMING
int snuffy_gemm_debug_read(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_gemm_dbg, sizeof(g_gemm_dbg)) == cudaSuccess ? 0 : 1;
}
#endif

}  // extern "C"
#pragma GCC visibility pop
