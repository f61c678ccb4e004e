// Snuffy sparse attention on the 5th-gen tensor cores (snuffy.py:160-168, 187-201).
//
//   S_j = Q_j Kp_j^T / sqrt(dk)   [N, Ksel]    softmax over the Ksel keys of each query row
//   O_j = P_j^T V_j               [Ksel, dk]   transposed aggregation: the contraction runs over the N queries
//
// Both contractions are split-bf16 tcgen05.mma (fp32 accumulate in TMEM) so the result matches fp32 to ~2^-16; with fp32
// SIMT math this kernel is ~8x compute-bound over its HBM time.
//
// Operands:
//   Q, V  : the Q|V projection GEMM writes its result directly as split-bf16 A-operand planes over [B*N, 2d]
//           (csrc/common.cuh).  A (row-tile, head) block of Q is 128 x dk, K-major: the A operand of S = Q Kp^T.  The same
//           block of V, read through an MN-major descriptor (M = dv contiguous in each 16-byte unit, K = query row), is the
//           A operand of O^T = V^T P.  No transposes, one bulk copy each.
//   Kp    : fp32 [B*Ksel, d]; each CTA splits its head's keys into shared memory once per work item.
//   P     : never leaves the SM: the softmax warps write it as MN-major split-bf16 planes (N = key, K = query).
//
// O is accumulated TRANSPOSED: O^T [dv, keys] = sum_queries V^T P has M = dv and N = all keys of the chunk, so one 128-query
// tile costs 8 k-steps x 2 MMAs of M = 128 (the hi and lo planes of V stacked along M: rows [0, dk) = V_hi^T (P_hi + P_lo),
// rows [dk, 2 dk) = V_lo^T (P_hi + P_lo)) and N = KP — 1664 tensor clocks at 208 keys, where O = P^T V (M = keys padded to
// 256, N = dk) needed 64 narrow MMAs that were bound by their shared-memory operand reads (3560 clocks measured).
//
// Work item = (bag, head, key chunk, row-tile range).  Per 128-query tile:
//   warp 0      bulk-copy producer: Q tile, V tile (single buffers, re-armed by tcgen05.commit)
//   warp 1      MMA issuer:  S(t) = Q Kp^T -> TMEM[0, KP);  O^T += V(t-1)^T P(t-1) -> TMEM[256, 256 + KP)
//   warps 2-17  one query row per thread, four warps per TMEM lane quadrant each owning a contiguous range of 8-key groups:
//               the row's scores are read from TMEM ONCE into registers and S is released at once (the next tile's scores are
//               computed while this one is normalised); row max and row sum are exchanged between the four parts through
//               shared memory; exp once per score; normalise + integer-pipe bf16 split -> P planes in smem.
//               At the end of the item O^T -> registers -> per-split partial (folded by fold_partials_kernel).
// Ksel > 256: two launches over key chunks, MODE 1 = per-chunk row statistics, MODE 2 = merged statistics, P and V^T P per
// chunk.  Measured timeline and what bounds each phase: profiles/r02_summary.md (round 1 design: profiles/r01_summary.md).
#include "tc_ptx.cuh"

namespace snuffy {

void launch_fold_partials(const float* part, int splits, int64_t n4, float* out, cudaStream_t stream);

#ifndef AT_PARTS_N
#define AT_PARTS_N 3
#endif
constexpr int AT_PARTS = AT_PARTS_N;    // softmax warps per TMEM lane quadrant (each owns a contiguous range of 8-key groups)
constexpr int AT_VG = (32 + AT_PARTS - 1) / AT_PARTS;       // 8-key groups one thread can own (KP <= 256: 32 groups in all)
constexpr int AT_NLD = AT_VG / 4;                           // full 32-column TMEM loads per thread
static_assert(AT_VG % 4 == 0 || AT_VG % 4 == 3, "the tail of a thread's key groups is loaded as 24 columns");
constexpr int AT_SOFT = 128 * AT_PARTS;             // softmax threads (warps 2 .. 2 + 4 * AT_PARTS - 1)
constexpr int AT_VWARP = 2 + 4 * AT_PARTS;          // the V producer warp
constexpr int AT_THREADS = 64 + AT_SOFT + 32;       // Q producer + MMA issuer + softmax warps + V producer
constexpr int AT_TILE = 128;            // queries per tile (UMMA M of S, K of O^T)
constexpr uint32_t AT_O_COL = 256;      // TMEM column of the O^T accumulator

// n / d for n < 2^31 as a multiply-high and a shift (the item decode runs once per work item in code that is cold in the
// instruction cache: four integer divisions cost over a thousand clocks there)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, int d, uint32_t mul, uint32_t shr) {
    return d == 1 ? n : __umulhi(n, mul) >> shr;
}
static void fast_div_setup(int d, uint32_t& mul, uint32_t& shr) {
    if (d <= 1) { mul = 0; shr = 0; return; }
    uint32_t lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;                    // ceil(log2 d)
    const uint64_t pw = 31 + lg;
    mul = (uint32_t)((((uint64_t)1 << pw) + (uint64_t)d - 1) / (uint64_t)d);
    shr = (uint32_t)(pw - 32);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifdef ATTN_DEBUG_TIMING
// Per-phase clock64 stamps of CTA 0 (tools/attn_timeline.py): [role 0 softmax | 1 mma | 2 producer][tile < 64][8 stamps]
__device__ long long g_attn_dbg[4 * 64 * 8];
#define DBG_STAMP(role, t, k) do { if (blockIdx.x == 0 && (t) < 64) g_attn_dbg[((role) * 64 + (t)) * 8 + (k)] = clock64(); } while (0)
#else
#define DBG_STAMP(role, t, k) do { } while (0)
#endif

struct AttnTcParams {
    const __nv_bfloat16* planes; int64_t plane_stride;   // Q|V planes over [rows, ldk], A layout, RC = 128
    int nkb;                          // k-blocks (of 32 columns) per row tile of the planes = ldk / 32
    int q_kb0, v_kb0;                 // first k-block of Q / V columns (head j adds j * dk / 32)
    const float* Kp;                  // [B*Ksel, d]
    __nv_bfloat16* k_planes;          // workspace: per (bag, key chunk, head) the K-major split-bf16 B operand of S = Q Kp^T,
                                      // [hi | lo][dk/8][KP][8], written by attn_key_planes_kernel (zero rows beyond Ksel)
    int B, N, Ksel, KP, h, dk, d;
    int splits, tiles_per_split;
    uint32_t div_mul[3], div_shr[3];  // item -> (split, key chunk, head, bag): division by splits, nkc, h as multiply-high + shift
    int nkc, KC;                      // key chunks per (bag, head) and keys per chunk (KC % 16 == 0; KP = padded chunk size)
    float* stats_part;                // [nkc][B, h, N, 2] per-chunk (row max of raw scores, sum of exp2): MODE 1 -> MODE 2
    float c_log2;                     // log2(e) / sqrt(dk)
    float* O_part;                    // [splits, B*Ksel, d]
    float* P_out;                     // [B, h, N, Ksel] or null (pre-dropout)
    float* stats_out;                 // [B, h, N, 2] (row max of the scaled scores, 1 / row sum) or null
    uint8_t* drop_mask;               // [B, h, N, ceil(Ksel / 8)] keep bits of the attention dropout (bit e of byte g = key 8 g + e) or null
    float drop_p; uint64_t seed, offset;
    const int64_t* cu_seqlens;        // packed variable-length bags: rows [cu[b], cu[b+1]) (null: b*N .. (b+1)*N)
};

// Splits the projected keys into the shared-memory image of the attention kernel's K operand, one block per (bag, key chunk,
// head): the attention CTAs then fetch a head's keys with two bulk copies (issued by the producer warp while the previous
// work item is still being finished) instead of splitting them themselves at every item start.
__global__ void __launch_bounds__(256)
attn_key_planes_kernel(const AttnTcParams p) {
    const int blk = blockIdx.x;                              // (b * nkc + kc) * h + j
    const int j = blk % p.h, kc = (blk / p.h) % p.nkc, b = blk / (p.h * p.nkc);
    const int dk = p.dk, KP = p.KP, key0 = kc * p.KC, kvalid = min(p.KC, p.Ksel - key0);
    const size_t plane = (size_t)KP * dk;                    // elements of one plane
    __nv_bfloat16* out = p.k_planes + (size_t)blk * 2 * plane;
    const int units = KP * (dk / 8);
    for (int idx = threadIdx.x; idx < units; idx += blockDim.x) {
        const int key = idx % KP, kg = idx / KP;
        bf16x8 hi, lo;
        if (key < kvalid) {
            const float* src = p.Kp + ((int64_t)b * p.Ksel + key0 + key) * p.d + j * dk + kg * 8;
            const float4 a = __ldg(reinterpret_cast<const float4*>(src));
            const float4 c = __ldg(reinterpret_cast<const float4*>(src + 4));
            const float f[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16(f[e], hi.v[e], lo.v[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { hi.v[e] = __float2bfloat16_rn(0.f); lo.v[e] = hi.v[e]; }
        }
        *reinterpret_cast<bf16x8*>(out + (size_t)idx * 8) = hi;
        *reinterpret_cast<bf16x8*>(out + plane + (size_t)idx * 8) = lo;
    }
}

// MODE 0: all keys of a head fit one chunk (Ksel <= 256): scores, softmax and V^T P fused in one pass.
// MODE 1: several key chunks, pass 1: per-chunk row statistics only (no P, no V, no O).
// MODE 2: several key chunks, pass 2: merge the chunk statistics, P for this chunk, O[chunk keys] += V^T P.
// EXTRA: the attention tensor is written out and / or attention dropout is drawn (kept out of the inference instantiation).
template <int MODE, bool EXTRA>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const AttnTcParams p) {
    extern __shared__ __align__(1024) unsigned char at_smem[];
    const int dk = p.dk, KP = p.KP;
    const uint32_t P_PLANE = (uint32_t)KP * 256u;            // [KP/8 key groups][128 queries][8 keys] bf16
    const uint32_t QV_PLANE = (uint32_t)AT_TILE * dk * 2u;   // [dk/8][128][8] bf16
    const uint32_t KP_PLANE = (uint32_t)KP * dk * 2u;        // [dk/8][KP][8] bf16
    // order matters: the M = 128 A operand of O^T = V^T P starts at a V plane and reads 16 groups of 2 KB whatever dk is (rows
    // beyond the valid ones only feed accumulator lanes nobody reads); what follows sV must be mapped shared memory
    unsigned char* sK = at_smem;
    unsigned char* sQ = sK + 2 * KP_PLANE;
    unsigned char* sV = sQ + 2 * QV_PLANE;
    unsigned char* sP = sV + 2 * QV_PLANE;
    // the P region doubles as the fp32 staging area of the O read-out: [up to KP + 31 keys][dk] when the operand is stacked
    const uint32_t stage_bytes = 2 * dk <= 128 ? ((uint32_t)(KP + 32) * dk * 4u + 127u) & ~127u : 0u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (2 * P_PLANE > stage_bytes ? 2 * P_PLANE : stage_bytes));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    // [tile parity][max | sum][AT_PARTS][128 rows]: with ONE exchange barrier per tile a fast warp could otherwise write tile
    // t + 1's values while a slow one still reads tile t's; writes to parity p of tile t + 2 come after the barrier of tile t + 1
    float* sRed0 = reinterpret_cast<float*>(bars + 13);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t q_full = b0, q_empty = b0 + 8, v_full = b0 + 16, v_empty = b0 + 24, s_full = b0 + 32,
                   s_empty = b0 + 40, p_full = b0 + 48, p_empty = b0 + 56, o_full = b0 + 64, k_ready = b0 + 72,
                   o_empty = b0 + 80, k_empty = b0 + 88;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(s_empty, 4 * AT_PARTS); mbar_init(p_full, 4 * AT_PARTS); mbar_init(p_empty, 1); mbar_init(o_full, 1);
        mbar_init(k_ready, 1); mbar_init(o_empty, 4 * AT_PARTS); mbar_init(k_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ksteps = dk / 16;
    const bool stacked = 2 * dk <= 128;     // V_hi | V_lo stacked along M of one MMA (the planes are contiguous in the MN-major layout)
    const int chunks_per_head = dk / PLANE_KB;
    const int64_t chunk_elems = (int64_t)AT_TILE * PLANE_KB;
    const int items = p.B * p.h * p.nkc * p.splits;
    uint32_t it = 0;          // tiles processed by this CTA so far (every role counts the same sequence)
    uint32_t item_no = 0;

    for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_no) {
        const uint32_t r1 = fast_div((uint32_t)item, p.splits, p.div_mul[0], p.div_shr[0]);
        const uint32_t r2 = fast_div(r1, p.nkc, p.div_mul[1], p.div_shr[1]);
        const int b = (int)fast_div(r2, p.h, p.div_mul[2], p.div_shr[2]);
        const int split = item - (int)r1 * p.splits, kc = (int)(r1 - r2 * p.nkc), j = (int)r2 - b * p.h;
        const int key0 = kc * p.KC;                                         // first key of this chunk
        const int kvalid = min(p.KC, p.Ksel - key0);                        // keys of this chunk (the rest of KP is padding)
        const int64_t g_lo = p.cu_seqlens ? p.cu_seqlens[b] : (int64_t)b * p.N;            // global rows of this bag
        const int64_t g_hi = p.cu_seqlens ? p.cu_seqlens[b + 1] : g_lo + p.N;
        const int64_t t_first = g_lo / AT_TILE, t_last = (g_hi + AT_TILE - 1) / AT_TILE;
        const int64_t t0 = t_first + (int64_t)split * p.tiles_per_split;
        const int64_t t1 = min(t_last, t0 + p.tiles_per_split);
        const int ntiles = (int)max((int64_t)0, t1 - t0);

        // No CTA-wide barrier between items: the producers fetch the next item's keys and first Q / V tiles while the softmax
        // warps still finish this one.  k_empty: the last S MMA of an item has retired (the K planes may be overwritten);
        // k_ready: the K planes of this item have landed; o_empty: the previous item's O has been read out (the MMA warp
        // waits before its first O MMA).
        if (warp == 0) {
            // ------------------------------------------------ Q producer (its own warp: a Q tile is requested the moment the
            // S MMAs of the previous tile retire, never behind a wait for the V buffer)
            if (item_no > 0) mbar_wait(k_empty, (item_no - 1) & 1);
            if (lane == 0) {
                const __nv_bfloat16* ksrc = p.k_planes + ((size_t)(b * p.nkc + kc) * p.h + j) * 2 * ((size_t)KP * dk);
                mbar_expect_tx(k_ready, 2 * KP_PLANE);
                bulk_g2s(smem_u32(sK), ksrc, KP_PLANE, k_ready);
                bulk_g2s(smem_u32(sK) + KP_PLANE, ksrc + (size_t)KP * dk, KP_PLANE, k_ready);
            }
            __syncwarp();
            for (int t = 0; t < ntiles; ++t, ++it) {
                const int64_t rt = t0 + t;
                const __nv_bfloat16* qsrc = p.planes + (rt * p.nkb + p.q_kb0 + j * chunks_per_head) * chunk_elems;
                if (lane == 0) DBG_STAMP(2, it, 0);
                mbar_wait(q_empty, (it & 1) ^ 1);
                if (lane == 0) DBG_STAMP(2, it, 1);
                if (lane == 0) {
                    mbar_expect_tx(q_full, 2 * QV_PLANE);
                    bulk_g2s(smem_u32(sQ), qsrc, QV_PLANE, q_full);
                    bulk_g2s(smem_u32(sQ) + QV_PLANE, qsrc + p.plane_stride, QV_PLANE, q_full);
                }
                __syncwarp();
            }
        } else if (warp == AT_VWARP) {
            // ------------------------------------------------ V producer
            if (MODE != 1) {
                for (int t = 0; t < ntiles; ++t, ++it) {
                    const int64_t rt = t0 + t;
                    const __nv_bfloat16* vsrc = p.planes + (rt * p.nkb + p.v_kb0 + j * chunks_per_head) * chunk_elems;
                    mbar_wait(v_empty, (it & 1) ^ 1);
                    if (lane == 0) DBG_STAMP(2, it, 2);
                    if (lane == 0) {
                        mbar_expect_tx(v_full, 2 * QV_PLANE);
                        bulk_g2s(smem_u32(sV), vsrc, QV_PLANE, v_full);
                        bulk_g2s(smem_u32(sV) + QV_PLANE, vsrc + p.plane_stride, QV_PLANE, v_full);
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 1) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);
            // O^T = V^T P: A = V planes (MN-major: 8 consecutive dv per 16-byte unit), B = P planes (MN-major: 8 consecutive keys)
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(KP >> 3) << 17) | (8u << 24);
            // Descriptors of k-step 0; a k-step advances the start-address field (bits 0-13, in 16-byte units).  The loops stay
            // rolled: this warp is paced by the tensor pipe, and a small body keeps the instruction cache for the softmax warps.
            const uint64_t q_hi0 = make_smem_desc(smem_u32(sQ), 2048, 128), q_lo0 = make_smem_desc(smem_u32(sQ) + QV_PLANE, 2048, 128);
            const uint64_t k_hi0 = make_smem_desc(smem_u32(sK), KP * 16, 128), k_lo0 = make_smem_desc(smem_u32(sK) + KP_PLANE, KP * 16, 128);
            const uint64_t p_hi0 = make_smem_desc(smem_u32(sP), 128, 2048), p_lo0 = make_smem_desc(smem_u32(sP) + P_PLANE, 128, 2048);
            const uint64_t v_hi0 = make_smem_desc(smem_u32(sV), 128, 2048), v_lo0 = make_smem_desc(smem_u32(sV) + QV_PLANE, 128, 2048);
            const uint64_t q_step = (2 * 2048) >> 4, k_step = (uint64_t)(2 * KP * 16) >> 4, pv_step = 256 >> 4;
            const uint32_t d_o = tmem_base + AT_O_COL;
            mbar_wait(k_ready, item_no & 1);
            // iteration t: S(t) = Q(t) Kp^T, then O^T += V(t-1)^T P(t-1)  (one code site each)
#pragma unroll 1
            for (int t = 0; t <= ntiles; ++t) {
                if (t < ntiles) {
                    if (lane == 0) DBG_STAMP(1, it, 0);
                    mbar_wait(q_full, it & 1);
                    if (lane == 0) DBG_STAMP(1, it, 1);
                    mbar_wait(s_empty, (it & 1) ^ 1);
                    if (lane == 0) DBG_STAMP(1, it, 2);
                    tc_fence_after();
                    if (lane == 0) {
#pragma unroll 1
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t qo = q_step * ks, ko = k_step * ks;
                            tc_mma_bf16(tmem_base, q_lo0 + qo, k_hi0 + ko, idesc1, ks ? 1u : 0u);
                            tc_mma_bf16(tmem_base, q_hi0 + qo, k_lo0 + ko, idesc1, 1u);
                            tc_mma_bf16(tmem_base, q_hi0 + qo, k_hi0 + ko, idesc1, 1u);
                        }
                        tc_commit(q_empty);
                        tc_commit(s_full);
                        if (t == ntiles - 1) tc_commit(k_empty);
                        DBG_STAMP(1, it, 3);
                    }
                    __syncwarp();
                    ++it;
                }
                if (MODE != 1 && t > 0) {
                    const uint32_t itp = it - (t < ntiles ? 2 : 1);          // the tile whose P is consumed
                    if (t == 1 && item_no > 0) mbar_wait(o_empty, (item_no - 1) & 1);     // the previous item's O has been read out
                    if (lane == 0) DBG_STAMP(1, itp, 4);
                    mbar_wait(p_full, itp & 1);
                    if (lane == 0) DBG_STAMP(1, itp, 5);
                    mbar_wait(v_full, itp & 1);
                    if (lane == 0) DBG_STAMP(1, itp, 6);
                    tc_fence_after();
                    if (lane == 0) {
#pragma unroll 1
                        for (int ks = 0; ks < AT_TILE / 16; ++ks) {
                            const uint64_t o = pv_step * ks;
                            const uint32_t acc0 = (t == 1 && ks == 0) ? 0u : 1u;
                            if (stacked) {           // lanes [0, dk): V_hi^T (P_lo + P_hi), lanes [dk, 2 dk): V_lo^T (P_lo + P_hi)
                                tc_mma_bf16(d_o, v_hi0 + o, p_lo0 + o, idesc2, acc0);
                                tc_mma_bf16(d_o, v_hi0 + o, p_hi0 + o, idesc2, 1u);
                            } else {                 // lanes [0, dk): V_lo^T P_hi + V_hi^T P_lo + V_hi^T P_hi
                                tc_mma_bf16(d_o, v_lo0 + o, p_hi0 + o, idesc2, acc0);
                                tc_mma_bf16(d_o, v_hi0 + o, p_lo0 + o, idesc2, 1u);
                                tc_mma_bf16(d_o, v_hi0 + o, p_hi0 + o, idesc2, 1u);
                            }
                        }
                        tc_commit(p_empty);
                        tc_commit(v_empty);
                        DBG_STAMP(1, itp, 7);
                    }
                    __syncwarp();
                }
            }
            if (ntiles == 0) {
                // an empty split issues nothing, but its hand-shakes still have to stay one phase apart from the consumers':
                // the softmax warps must have passed the previous item's o_full before this item's o_full is signalled
                if (MODE != 1 && item_no > 0) mbar_wait(o_empty, (item_no - 1) & 1);
                if (lane == 0) mbar_arrive(k_empty);             // the planes are free at once
            }
            if (lane == 0 && MODE != 1) tc_commit(o_full);
            __syncwarp();
        } else {
            // ------------------------------------------------ softmax warps.  Thread = one query row of the tile; the four
            // warps of a TMEM lane quadrant own contiguous ranges of 8-key groups and exchange (max, sum) through smem.
            const int quad = warp & 3, part = (warp - 2) >> 2;      // part 0 .. AT_PARTS-1
            const int rr = quad * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
            const int ngroups = KP / 8, gbase = ngroups / AT_PARTS, grem = ngroups % AT_PARTS;
            const int g0 = part * gbase + min(part, grem);          // first 8-key group of this warp
            const int ng = gbase + (part < grem ? 1 : 0);           // its group count (<= AT_VG: KP <= 256)
            constexpr int RED = AT_PARTS * 128;
            const uint32_t bar_id = 1 + quad;                       // named barrier of this quadrant's part warps
            for (int t = 0; t < ntiles; ++t, ++it) {
                const int64_t g = (t0 + t) * AT_TILE + rr;
                const bool valid = g >= g_lo && g < g_hi;
                const int n = (int)(g - g_lo);
                const int64_t srow = ((int64_t)b * p.h + j) * p.N + n;      // row of the [B, h, N, *] statistics
                float* sRed = sRed0 + (it & 1) * 2 * RED;
                const bool stamp = warp == 2 && lane == 0;
                if (stamp) DBG_STAMP(0, it, 0);
                mbar_wait(s_full, it & 1);
                if (stamp) DBG_STAMP(0, it, 1);
                tc_fence_after();
                // ---- this thread's scores: TMEM -> registers ONCE, then S is free for the next tile's MMAs
                float v[AT_VG * 8];
#pragma unroll
                for (int i = 0; i < AT_NLD; ++i)
                    if (i == 0 || ng > 4 * i) tc_ld32(lane_addr + (uint32_t)(g0 * 8 + 32 * i), *reinterpret_cast<float(*)[32]>(v + 32 * i));
                if (AT_VG % 4 == 3 && ng > 4 * AT_NLD) tc_ld24(lane_addr + (uint32_t)(g0 * 8 + 32 * AT_NLD), v + 32 * AT_NLD);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty);
                if (stamp) DBG_STAMP(0, it, 2);
                // padding keys (the chunk is padded to a multiple of 16) carry Q . 0 = 0: give them -inf once, so that the loops
                // below need no masks (max ignores them, exp2 turns them into exact zeros)
                if ((g0 + ng) * 8 > kvalid) {
#pragma unroll
                    for (int gi = 0; gi < AT_VG; ++gi) {
                        if (gi < ng && (g0 + gi) * 8 + 8 > kvalid) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) if ((g0 + gi) * 8 + e >= kvalid) v[gi * 8 + e] = -INFINITY;
                        }
                    }
                }
                float mx = -INFINITY, inv = 0.f, mc = 0.f, rescale = 1.f;
                if (MODE != 2) {
                    // ---- row max over this warp's keys, then exp ONCE per score relative to that LOCAL max: the parts of a
                    // row exchange (max, sum) once and rescale (the factor is folded into the normalisation below)
                    float ml = -INFINITY;
#pragma unroll
                    for (int gi = 0; gi < AT_VG; ++gi) {
                        if (gi < ng) {
                            const float* sv = v + gi * 8;
                            const float m01 = fmaxf(fmaxf(sv[0], sv[1]), sv[2]), m23 = fmaxf(fmaxf(sv[3], sv[4]), sv[5]);
                            ml = fmaxf(fmaxf(ml, m01), fmaxf(fmaxf(m23, sv[6]), sv[7]));
                        }
                    }
                    // rows of a neighbouring bag / padding: exp2(s*c - inf) = 0 -> P = 0 exactly (never inf * 0 = NaN); a part
                    // that owns only padding keys has ml = -inf: all its exp2(-inf - 0) are exact zeros
                    const float mcl = !valid ? INFINITY : (ml == -INFINITY ? 0.f : ml * p.c_log2);
                    float sum = 0.f, sum2 = 0.f;
#pragma unroll
                    for (int gi = 0; gi < AT_VG; ++gi) {
                        if (gi < ng) {
                            float* sv = v + gi * 8;
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                sv[e] = ex2_approx(fmaf(sv[e], p.c_log2, -mcl));
                                sv[e + 1] = ex2_approx(fmaf(sv[e + 1], p.c_log2, -mcl));
                                sum += sv[e]; sum2 += sv[e + 1];
                            }
                        }
                    }
                    sRed[part * 128 + rr] = ml;
                    sRed[RED + part * 128 + rr] = sum + sum2;
                    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(AT_PARTS * 32) : "memory");
                    float mq[AT_PARTS];
#pragma unroll
                    for (int q = 0; q < AT_PARTS; ++q) { mq[q] = sRed[q * 128 + rr]; mx = fmaxf(mx, mq[q]); }
                    if (!valid) mx = 0.f;
                    sum = 0.f;
#pragma unroll
                    for (int q = 0; q < AT_PARTS; ++q) sum += sRed[RED + q * 128 + rr] * ex2_approx((mq[q] - mx) * p.c_log2);
                    rescale = ex2_approx((ml - mx) * p.c_log2);          // this part's exponentials relative to the row max
                    if (stamp) DBG_STAMP(0, it, 3);
                    if (MODE == 1) {
                        // per-chunk partial statistics for pass 2 (raw-score max, sum of exp2 relative to it)
                        if (valid && part == 0) {
                            float* sp = p.stats_part + ((int64_t)kc * p.B * p.h * p.N + srow) * 2;
                            sp[0] = mx; sp[1] = sum;
                        }
                        continue;
                    }
                    inv = valid ? 1.f / sum : 0.f;
                    (void)mc;
                } else {
                    // ---- merge the chunk statistics of this row (written by the MODE 1 launch)
                    float lsum = 0.f;
                    if (valid) {
                        const float* sp = p.stats_part + srow * 2;
                        const int64_t cs = (int64_t)p.B * p.h * p.N * 2;
                        for (int c2 = 0; c2 < p.nkc; ++c2) mx = fmaxf(mx, __ldcg(sp + c2 * cs));
                        for (int c2 = 0; c2 < p.nkc; ++c2)
                            lsum += __ldcg(sp + c2 * cs + 1) * ex2_approx((__ldcg(sp + c2 * cs) - mx) * p.c_log2);
                        inv = 1.f / lsum;
                    } else {
                        mx = 0.f;
                    }
                    mc = valid ? mx * p.c_log2 : INFINITY;
#pragma unroll
                    for (int gi = 0; gi < AT_VG; ++gi) {
                        if (gi < ng) {
                            float* sv = v + gi * 8;
#pragma unroll
                            for (int e = 0; e < 8; ++e) sv[e] = ex2_approx(fmaf(sv[e], p.c_log2, -mc));
                        }
                    }
                }
                if (p.stats_out && valid && part == 0 && kc == 0) {
                    float* so = p.stats_out + srow * 2;
                    so[0] = mx * (p.c_log2 * 0.69314718055994530942f);     // max of the scaled scores (natural units)
                    so[1] = inv;
                }
                const DrawKey dkey = rng_resolve(p.seed, p.offset);
                const DropFast dfast = drop_fast_setup(dkey.seed, dkey.offset, EXTRA ? p.drop_p : 0.f);
                const bool small_index = (int64_t)p.B * p.h * p.N * p.Ksel < (1ll << 32);
                // The attention dropout is drawn in a ROLLED loop (one 8-key group per trip) into a bit mask: unrolled over this
                // thread's ~70 scores the hash alone was 1400 instructions in the tile loop, and the softmax warps then ran out of
                // the instruction cache (0.146 -> 0.436 ms per 8 slides); the unrolled P phase below only tests a bit.
                uint64_t keep_lo = 0ull, keep_hi = 0ull;
                if (EXTRA && p.drop_p > 0.f && valid) {
#pragma unroll 1
                    for (int gi = 0; gi < ng; ++gi) {
                        const int kb = (g0 + gi) * 8;
                        uint32_t bits = 0;
                        if (small_index) {                           // element indices below 2^32: hoisted form of the same draw
                            const uint32_t e0 = (uint32_t)srow * (uint32_t)p.Ksel + (uint32_t)(key0 + kb);
                            if ((e0 & 1u) == 0u) {               // the usual case (even Ksel): one hash per pair of scores
#pragma unroll
                                for (int e = 0; e < 8; e += 2) {
                                    const uint32_t x = drop_fast_hash(dfast, (e0 + e) >> 1);
                                    bits |= ((x & 0xFFFFu) >= dfast.thr ? 1u : 0u) << e;
                                    bits |= ((x >> 16) >= dfast.thr ? 1u : 0u) << (e + 1);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) bits |= (drop_fast_keep(dfast, e0 + e) ? 1u : 0u) << e;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const uint64_t idx = ((uint64_t)srow * p.Ksel + key0 + kb + e);
                                bits |= (drop_keep_scale(dkey.seed, dkey.offset, idx, p.drop_p) != 0.f ? 1u : 0u) << e;
                            }
                        }
                        // the backward kernel reads the draw back instead of hashing every score again (twice)
                        if (p.drop_mask && kb < kvalid) p.drop_mask[srow * ((p.Ksel + 7) / 8) + ((key0 + kb) >> 3)] = (uint8_t)bits;
                        if (gi < 8) keep_lo |= (uint64_t)bits << (8 * gi); else keep_hi |= (uint64_t)bits << (8 * (gi - 8));
                    }
                }
                if (stamp) DBG_STAMP(0, it, 4);
                mbar_wait(p_empty, (it & 1) ^ 1);            // the MMAs of the previous tile have consumed P
                if (stamp) DBG_STAMP(0, it, 5);
                const float pscale = valid ? inv * rescale : 0.f;
#pragma unroll
                for (int gi = 0; gi < AT_VG; ++gi) {
                    if (gi < ng) {
                        const int kgp = g0 + gi, kb = kgp * 8;
                        float w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] = v[gi * 8 + e] * pscale;
                        if (EXTRA && p.P_out && valid) {
                            float* po = p.P_out + srow * p.Ksel + key0 + kb;
#pragma unroll
                            for (int e = 0; e < 8; ++e) if (kb + e < kvalid) po[e] = w[e];
                        }
                        if (EXTRA && p.drop_p > 0.f && valid) {
                            const uint32_t b8 = (uint32_t)(gi < 8 ? keep_lo >> (8 * gi) : keep_hi >> (8 * (gi - 8))) & 0xFFu;
#pragma unroll
                            for (int e = 0; e < 8; ++e) w[e] = ((b8 >> e) & 1u) ? w[e] * dfast.scale : 0.f;
                        }
                        // hi = truncated bf16 (exact), lo = bf16 of the exact remainder: integer/FMA pipes only
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t ua = __float_as_uint(w[2 * q]), ub = __float_as_uint(w[2 * q + 1]);
                            const float la = w[2 * q] - __uint_as_float(ua & 0xFFFF0000u);
                            const float lb = w[2 * q + 1] - __uint_as_float(ub & 0xFFFF0000u);
                            hw[q] = __byte_perm(ua, ub, 0x7632);
                            lw[q] = __byte_perm(__float_as_uint(la), __float_as_uint(lb), 0x7632);
                        }
                        *reinterpret_cast<uint4*>(sP + (size_t)kgp * 2048 + rr * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(sP + P_PLANE + (size_t)kgp * 2048 + rr * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (stamp) DBG_STAMP(0, it, 6);
                if (lane == 0) mbar_arrive(p_full);
            }
            if (MODE != 1) {
                // ---- item epilogue: O^T (TMEM lane = dv, column = key) -> this split's partial.  With the stacked operand the
                // lanes [dk, 2 dk) hold the V_lo contribution of column dv = lane - dk: staged through shared memory (the P planes
                // are idle now) and added by the lanes [0, dk).
                mbar_wait(o_full, item_no & 1);
                tc_fence_after();
                // o_full already orders every warp's P stores before the staging stores below (P stores -> p_full -> the MMAs ->
                // their commit); this barrier states the same order among the softmax warps in a form compute-sanitizer can see
                asm volatile("bar.sync %0, %1;" ::"r"(5), "r"(AT_SOFT) : "memory");
                if (warp == 2 && lane == 0) DBG_STAMP(2, it, 3);
                const int L = rr;                                    // TMEM lane of this thread
                float* stage = reinterpret_cast<float*>(sP);         // [KP keys][dk] fp32 (<= the P planes: KP * 512 B)
                float* dst = p.O_part + (((int64_t)split * p.B + b) * p.Ksel + key0) * p.d + j * dk;
                const bool keep = ntiles > 0;                        // an empty split contributes zeros (its TMEM is stale)
                const int nch = (kvalid + 31) / 32;                  // 32-key chunks, dealt round-robin to the parts of a quadrant
                if (stacked) {
                    // lanes [0, dk) (the V_hi rows) store their accumulator transposed, stage[key][dv]; lanes [dk, 2 dk) add
                    // the V_lo rows of the same dv and write the sum out (a warp writes 128 contiguous bytes per key)
                    if (L < dk) {
                        for (int c = part; c < nch; c += AT_PARTS) {
                            float o[32];
                            tc_ld32(lane_addr + AT_O_COL + (uint32_t)(c * 32), o);
#pragma unroll
                            for (int e = 0; e < 32; ++e) stage[(c * 32 + e) * dk + L] = o[e];     // rows up to KP + 31 < 288: inside sP
                        }
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(5), "r"(AT_SOFT) : "memory");
                    if (warp == 2 && lane == 0) DBG_STAMP(2, it, 4);
                    if (L >= dk && L < 2 * dk) {
                        for (int c = part; c < nch; c += AT_PARTS) {
                            float o[32];
                            tc_ld32(lane_addr + AT_O_COL + (uint32_t)(c * 32), o);
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = keep ? o[e] + stage[(c * 32 + e) * dk + (L - dk)] : 0.f;
                            float* dg = dst + (int64_t)(c * 32) * p.d + (L - dk);
                            if (c * 32 + 32 <= kvalid) {                 // whole chunk valid (warp-uniform): no per-key branches
#pragma unroll
                                for (int e = 0; e < 32; ++e) dg[(int64_t)e * p.d] = o[e];
                            } else {
#pragma unroll
                                for (int e = 0; e < 32; ++e) if (c * 32 + e < kvalid) dg[(int64_t)e * p.d] = o[e];
                            }
                        }
                    }
                } else if (L < dk) {
                    for (int c = part; c < nch; c += AT_PARTS) {
                        float o[32];
                        tc_ld32(lane_addr + AT_O_COL + (uint32_t)(c * 32), o);
                        float* dg = dst + (int64_t)(c * 32) * p.d + L;
                        if (c * 32 + 32 <= kvalid) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) dg[(int64_t)e * p.d] = keep ? o[e] : 0.f;
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e) if (c * 32 + e < kvalid) dg[(int64_t)e * p.d] = keep ? o[e] : 0.f;
                        }
                    }
                }
                if (warp == 2 && lane == 0) DBG_STAMP(3, it, 4);
                tc_fence_before();
                // the staging area is the P planes: nobody may start the next item's P before every reader is done
                asm volatile("bar.sync %0, %1;" ::"r"(5), "r"(AT_PARTS * 128) : "memory");
                if (lane == 0) mbar_arrive(o_empty);
                if (warp == 2 && lane == 0) DBG_STAMP(2, it, 5);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

struct AttnTcPlan { int KP, KC, nkc, splits, tiles_per_split, grid; size_t smem; bool ok; };

static size_t attn_tc_smem(int KP, int dk) {
    const size_t kq = (size_t)2 * KP * dk * 2 + (size_t)2 * AT_TILE * dk * 2;       // K planes, Q planes
    const size_t v = (size_t)2 * AT_TILE * dk * 2;                                  // V planes
    size_t pp = (size_t)2 * KP * 256;                                               // P planes / O staging area
    if (2 * dk <= 128) { const size_t st = (((size_t)(KP + 32) * dk * 4) + 127) & ~(size_t)127; if (st > pp) pp = st; }
    size_t total = kq + v + pp + 128 + 4 * AT_PARTS * 128 * 4;
    // the M = 128 A operand of O^T = V^T P reads 16 groups of 2 KB from the start of a V plane (csrc comment at the top)
    const size_t a_reach = kq + v / 2 + (size_t)16 * 2048 + 16;
    return total > a_reach ? total : a_reach;
}

static AttnTcPlan plan_attn_tc(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d) {
    AttnTcPlan pl{};
    if (h <= 0 || d % h) return pl;
    const int dk = (int)(d / h);
    if (dk % 32 || dk > 128 || Ksel < 1 || Ksel > 65536) return pl;
    // fewest key chunks such that a chunk fits TMEM (S: KP <= 256 columns next to the O accumulators) and the operand
    // tiles fit 227 KB of shared memory
    for (int nkc = 1; nkc <= 512; ++nkc) {
        const int KC = (int)(((Ksel + nkc - 1) / nkc + 15) / 16 * 16);
        if (KC > 256) continue;                               // S: KC TMEM columns from 0, O^T: KC columns from 256
        const size_t smem = attn_tc_smem(KC, dk);
        if (smem > 227 * 1024) continue;
        pl.nkc = (int)((Ksel + KC - 1) / KC); pl.KC = KC; pl.KP = KC; pl.smem = smem;
        break;
    }
    if (pl.nkc == 0) return pl;
    const int64_t tiles = (N + AT_TILE - 1) / AT_TILE + 1;         // a bag may straddle one extra row tile
    const int64_t per = B * h * pl.nkc;
    // Row tiles per work item: the persistent CTAs take items round-robin, so the launch lasts ceil(items / SMs) rounds of one
    // item each.  Pick the split that minimises rounds x (tiles per item + per-item overhead: last O MMAs, O read-out, item
    // set-up — measured ~10 k clocks against ~5.5 k per tile at 208 keys x 64, profiles/r02_summary.md) plus the cost of
    // folding `splits` partial outputs, all in units of one tile (~3.6 us).
    const double fold_per_split = (double)B * (double)Ksel * (double)d * 4.0 / 4.0e6 / 3.6;
    double best = 1e30;
    int best_t = (int)tiles;
    for (int64_t t = 1; t <= tiles; ++t) {
        const int64_t s = (tiles + t - 1) / t;
        if (s > 64) continue;
        const int64_t rounds = (per * s + sm_count() - 1) / sm_count();
        const double cost = (double)rounds * ((double)t + 1.8) + fold_per_split * (double)s;
        if (cost < best - 1e-9) { best = cost; best_t = (int)t; }
    }
    pl.tiles_per_split = best_t;
    pl.splits = (int)((tiles + pl.tiles_per_split - 1) / pl.tiles_per_split);
    const int64_t items = per * pl.splits;
    pl.grid = (int)(items < sm_count() ? items : sm_count());
    pl.ok = true;
    return pl;
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// bytes of workspace for snuffy_sparse_attn_tc_fwd, or -1 when the shape is not served by the tensor-core kernel
// (needs dk % 32 == 0 and dk <= 128) -> use snuffy_sparse_attn_fwd.  Ksel > 256 (or a head too wide for one chunk) runs as
// two launches over key chunks: per-chunk row statistics, then P and P^T V per chunk with the merged statistics.
int64_t snuffy_sparse_attn_tc_workspace(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d) {
    const AttnTcPlan pl = plan_attn_tc(B, N, Ksel, h, d);
    if (!pl.ok) return -1;
    int64_t bytes = (int64_t)pl.splits * B * Ksel * d * 4 + 256;
    bytes += (int64_t)B * pl.nkc * pl.KP * d * 4;                        // split-bf16 key planes (hi + lo) per (bag, chunk, head)
    if (pl.nkc > 1) bytes += (int64_t)pl.nkc * B * h * N * 2 * 4;
    return bytes;
}

// Same contract as snuffy_sparse_attn_fwd, but Q and V arrive as the split-bf16 planes the Q|V projection wrote
// (planes over [B*N, ldk] columns, Q at column q_col0, V at column v_col0; ldk, q_col0, v_col0 multiples of 32).
static int launch_attn_tc(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0, int64_t v_col0,
                          const float* Kp, int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d,
                          float dropout_p, uint64_t seed, uint64_t offset, float* O, float* P_out, float* stats_out,
                          uint8_t* drop_mask, void* workspace, int64_t workspace_bytes, const int64_t* cu_seqlens,
                          cudaStream_t stream) {
    SNUFFY_REQUIRE(qv_planes && Kp && O && workspace, "snuffy_sparse_attn_tc_fwd: null pointer");
    const AttnTcPlan pl = plan_attn_tc(B, N, Ksel, h, d);
    SNUFFY_REQUIRE(pl.ok, "snuffy_sparse_attn_tc_fwd: unsupported shape (h=%lld d=%lld Ksel=%lld)", (long long)h,
                   (long long)d, (long long)Ksel);
    SNUFFY_REQUIRE(ldk % 32 == 0 && q_col0 % 32 == 0 && v_col0 % 32 == 0 && q_col0 + d <= ldk && v_col0 + d <= ldk,
                   "snuffy_sparse_attn_tc_fwd: Q/V column ranges must be 32-aligned inside the planes");
    SNUFFY_REQUIRE(workspace_bytes >= snuffy_sparse_attn_tc_workspace(B, N, Ksel, h, d),
                   "snuffy_sparse_attn_tc_fwd: workspace too small");
    SNUFFY_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "snuffy_sparse_attn_tc_fwd: dropout_p out of range");
    SNUFFY_REQUIRE((uintptr_t)Kp % 16 == 0 && (uintptr_t)O % 16 == 0 && (uintptr_t)qv_planes % 16 == 0 && d % 4 == 0,
                   "snuffy_sparse_attn_tc_fwd: pointers must be 16-byte aligned");
    AttnTcParams p{};
    p.planes = reinterpret_cast<const __nv_bfloat16*>(qv_planes); p.plane_stride = plane_stride;
    p.nkb = (int)(ldk / 32); p.q_kb0 = (int)(q_col0 / 32); p.v_kb0 = (int)(v_col0 / 32);
    p.Kp = Kp; p.B = (int)B; p.N = (int)N; p.Ksel = (int)Ksel; p.KP = pl.KP; p.h = (int)h; p.dk = (int)(d / h); p.d = (int)d;
    p.splits = pl.splits; p.tiles_per_split = pl.tiles_per_split;
    p.nkc = pl.nkc; p.KC = pl.KC;
    fast_div_setup(p.splits, p.div_mul[0], p.div_shr[0]);
    fast_div_setup(p.nkc, p.div_mul[1], p.div_shr[1]);
    fast_div_setup(p.h, p.div_mul[2], p.div_shr[2]);
    p.c_log2 = (float)(1.4426950408889634 / sqrt((double)p.dk));
    p.O_part = reinterpret_cast<float*>(workspace);
    p.k_planes = reinterpret_cast<__nv_bfloat16*>(p.O_part + (int64_t)pl.splits * B * Ksel * d);
    p.stats_part = reinterpret_cast<float*>(p.k_planes + (int64_t)B * pl.nkc * pl.KP * d * 2);
    p.P_out = P_out; p.stats_out = stats_out; p.drop_mask = drop_mask;
    p.drop_p = dropout_p; p.seed = seed; p.offset = offset;
    p.cu_seqlens = cu_seqlens;
    SNUFFY_REQUIRE((uintptr_t)workspace % 16 == 0, "snuffy_sparse_attn_tc_fwd: workspace must be 16-byte aligned");
    attn_key_planes_kernel<<<(unsigned)(B * pl.nkc * h), 256, 0, stream>>>(p);
    const bool extra = P_out != nullptr || dropout_p > 0.f;
#define SNUFFY_ATTN_LAUNCH(MODE_, EXTRA_)                                                                                       \
    do {                                                                                                                        \
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_tc_kernel<MODE_, EXTRA_>), (int)pl.smem));          \
        attn_tc_kernel<MODE_, EXTRA_><<<pl.grid, AT_THREADS, pl.smem, stream>>>(p);                                             \
    } while (0)
    if (pl.nkc == 1) {
        if (extra) SNUFFY_ATTN_LAUNCH(0, true); else SNUFFY_ATTN_LAUNCH(0, false);
    } else {
        SNUFFY_ATTN_LAUNCH(1, false);
        if (extra) SNUFFY_ATTN_LAUNCH(2, true); else SNUFFY_ATTN_LAUNCH(2, false);
    }
#undef SNUFFY_ATTN_LAUNCH
    launch_fold_partials(p.O_part, pl.splits, B * Ksel * d / 4, O, stream);
    return check_launch("snuffy_sparse_attn_tc_fwd", pl.nkc == 1 ? 3 : 4);
}

int snuffy_sparse_attn_tc_fwd(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0, int64_t v_col0,
                              const float* Kp, int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d,
                              float dropout_p, uint64_t seed, uint64_t offset, float* O, float* P_out, float* stats_out,
                              uint8_t* drop_mask, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    return launch_attn_tc(qv_planes, plane_stride, ldk, q_col0, v_col0, Kp, B, N, Ksel, h, d, dropout_p, seed, offset, O, P_out,
                          stats_out, drop_mask, workspace, workspace_bytes, nullptr, stream);
}

// Packed variable-length bags (inference): the planes cover the packed [T, ldk] rows, bag b = rows
// [cu_seqlens[b], cu_seqlens[b+1]), every bag has Ksel keys (Kp [B*Ksel, d]); max_n = longest bag sizes the work split and
// the workspace (snuffy_sparse_attn_tc_workspace(B, max_n, ...)).
int snuffy_sparse_attn_tc_varlen_fwd(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0,
                                     int64_t v_col0, const float* Kp, const int64_t* cu_seqlens, int64_t B, int64_t max_n,
                                     int64_t Ksel, int64_t h, int64_t d, float* O, void* workspace, int64_t workspace_bytes,
                                     cudaStream_t stream) {
    SNUFFY_REQUIRE(cu_seqlens, "snuffy_sparse_attn_tc_varlen_fwd: null cu_seqlens");
    return launch_attn_tc(qv_planes, plane_stride, ldk, q_col0, v_col0, Kp, B, max_n, Ksel, h, d, 0.f, 0, 0, O, nullptr, nullptr,
                          nullptr, workspace, workspace_bytes, cu_seqlens, stream);
}

#ifdef ATTN_DEBUG_TIMING
int snuffy_attn_debug_read(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_attn_dbg, sizeof(g_attn_dbg)) == cudaSuccess ? 0 : 1;
}
#endif

}  // extern "C"
#pragma GCC visibility pop
