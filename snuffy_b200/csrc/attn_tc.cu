// Snuffy sparse attention on the 5th-gen tensor cores (snuffy.py:160-168, 187-201).
//
//   S_j = Q_j Kp_j^T / sqrt(dk)   [N, Ksel]    softmax over the Ksel keys of each query row
//   O_j = P_j^T V_j               [Ksel, dk]   transposed aggregation: the contraction runs over the N queries
//
// Both contractions are split-bf16 3-pass tcgen05.mma (fp32 accumulate in TMEM) so the result matches fp32 to
// ~2^-16; with fp32 SIMT math this kernel is ~8x compute-bound over its HBM time.
//
// Operands:
//   Q, V  : the Q|V projection GEMM writes its result directly as split-bf16 A-operand planes over
//           [B*N, 2d] (csrc/common.cuh).  A (row-tile, head) block of Q is 128 x dk, K-major: the A operand of
//           S = Q Kp^T.  The same block of V, read through an MN-major descriptor, is the B operand of
//           O = P^T V (N = dv contiguous in each 16-byte unit, K = query row).  No transposes, one bulk copy each.
//   Kp    : fp32 [B*Ksel, d]; each CTA splits its head's keys into shared memory once per work item.
//   P     : never leaves the SM: the softmax warps write it as MN-major split-bf16 planes (M = key, K = query).
//
// Work item = (bag, head, key chunk, row-tile range).  Per 128-query tile:
//   warp 0     bulk-copy producer: Q tile, V tile (single buffers, re-armed by tcgen05.commit)
//   warp 1     MMA issuer:  S(t) = Q Kp^T -> TMEM[0, KP);  O += P(t-1)^T V(t-1) -> TMEM[256, ...): per 128-key block 2 dk
//              columns, [0, dk) = P_hi^T V_hi + P_lo^T V_hi, [dk, 2 dk) = P_hi^T V_lo (the two V planes are one N = 2 dk operand)
//   warps 2-9  one query row per thread (AT_PARTS warps per TMEM lane quadrant split the key chunks): row max, then exp ONCE
//              per score parked back into TMEM (tcgen05.st), then normalise + integer-pipe bf16 split -> P planes in smem;
//              at the end of the item O -> registers -> per-split partial (folded by fold_partials_kernel).
// Ksel > 256 (or a wide head): two launches over key chunks, MODE 1 = per-chunk row statistics, MODE 2 = merged statistics,
// P and P^T V per chunk.  Measured timeline and what bounds each phase: profiles/r01_summary.md (rounds 1c, 1d).
#include "tc_ptx.cuh"

namespace snuffy {

void launch_fold_partials(const float* part, int splits, int64_t n4, float* out, cudaStream_t stream);

#ifndef AT_PARTS_N
#define AT_PARTS_N 2
#endif
constexpr int AT_PARTS = AT_PARTS_N;     // softmax warps per TMEM lane quadrant (each owns a contiguous range of key chunks)
constexpr int AT_THREADS = 64 + 128 * AT_PARTS;   // producer + MMA issuer + 4 * AT_PARTS softmax warps
constexpr int AT_TILE = 128;            // queries per tile (UMMA M)
constexpr uint32_t AT_O_COL = 256;      // TMEM column of the O accumulators

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifdef ATTN_DEBUG_TIMING
// Per-phase clock64 stamps of CTA 0 (tools/attn_timeline.py): [role 0 softmax | 1 mma | 2 producer][tile < 64][8 stamps]
__device__ long long g_attn_dbg[3 * 64 * 8];
#define DBG_STAMP(role, t, k) do { if (blockIdx.x == 0 && (t) < 64) g_attn_dbg[((role) * 64 + (t)) * 8 + (k)] = clock64(); } while (0)
#else
#define DBG_STAMP(role, t, k) do { } while (0)
#endif

struct AttnTcParams {
    const __nv_bfloat16* planes; int64_t plane_stride;   // Q|V planes over [rows, ldk], A layout, RC = 128
    int nkb;                          // k-blocks (of 32 columns) per row tile of the planes = ldk / 32
    int q_kb0, v_kb0;                 // first k-block of Q / V columns (head j adds j * dk / 32)
    const float* Kp;                  // [B*Ksel, d]
    int B, N, Ksel, KP, h, dk, d;
    int splits, tiles_per_split;
    int nkc, KC;                      // key chunks per (bag, head) and keys per chunk (KC % 16 == 0; KP = padded chunk size)
    float* stats_part;                // [nkc][B, h, N, 2] per-chunk (row max of raw scores, sum of exp2): MODE 1 -> MODE 2
    float c_log2;                     // log2(e) / sqrt(dk)
    float* O_part;                    // [splits, B*Ksel, d]
    float* P_out;                     // [B, h, N, Ksel] or null (pre-dropout)
    float* stats_out;                 // [B, h, N, 2] (row max of the scaled scores, 1 / row sum) or null
    float drop_p; uint64_t seed, offset;
    const int64_t* cu_seqlens;        // packed variable-length bags: rows [cu[b], cu[b+1]) (null: b*N .. (b+1)*N)
};

// MODE 0: all keys of a head fit one chunk (Ksel <= 256): scores, softmax and P^T V fused in one pass.
// MODE 1: several key chunks, pass 1: per-chunk row statistics only (no P, no V, no O).
// MODE 2: several key chunks, pass 2: merge the chunk statistics, P for this chunk, O[chunk keys] += P^T V.
template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const AttnTcParams p) {
    extern __shared__ __align__(1024) unsigned char at_smem[];
    const int dk = p.dk, KP = p.KP;
    const uint32_t P_PLANE = (uint32_t)KP * 256u;            // [KP/8 key groups][128 queries][8 keys] bf16
    const uint32_t QV_PLANE = (uint32_t)AT_TILE * dk * 2u;   // [dk/8][128][8] bf16
    const uint32_t KP_PLANE = (uint32_t)KP * dk * 2u;        // [dk/8][KP][8] bf16
    // order matters: the second 128-key MMA block reads past KP inside sP; what follows must be mapped smem
    unsigned char* sP = at_smem;
    unsigned char* sK = sP + 2 * P_PLANE;
    unsigned char* sQ = sK + 2 * KP_PLANE;
    unsigned char* sV = sQ + 2 * QV_PLANE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * QV_PLANE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* sRed = reinterpret_cast<float*>(bars + 10);       // [max | sum][AT_PARTS][128 rows] (single-buffered: s_full(t+1) orders all reads of tile t first)
    const uint32_t b0 = smem_u32(bars);
    const uint32_t q_full = b0, q_empty = b0 + 8, v_full = b0 + 16, v_empty = b0 + 24, s_full = b0 + 32,
                   s_empty = b0 + 40, p_full = b0 + 48, p_empty = b0 + 56, o_full = b0 + 64;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(s_empty, 4 * AT_PARTS); mbar_init(p_full, 4 * AT_PARTS); mbar_init(p_empty, 1); mbar_init(o_full, 1);
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

    const int mblocks = (KP + 127) / 128;
    const int ksteps = dk / 16;
    const int chunks_per_head = dk / PLANE_KB;
    const int64_t chunk_elems = (int64_t)AT_TILE * PLANE_KB;
    const int items = p.B * p.h * p.nkc * p.splits;
    uint32_t it = 0;          // tiles processed by this CTA so far (every role counts the same sequence)
    uint32_t item_no = 0;

    for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_no) {
        const int split = item % p.splits;
        const int kc = (item / p.splits) % p.nkc;
        const int j = (item / (p.splits * p.nkc)) % p.h;
        const int b = item / (p.splits * p.nkc * p.h);
        const int key0 = kc * p.KC;                                         // first key of this chunk
        const int kvalid = min(p.KC, p.Ksel - key0);                        // keys of this chunk (the rest of KP is padding)
        const int64_t g_lo = p.cu_seqlens ? p.cu_seqlens[b] : (int64_t)b * p.N;            // global rows of this bag
        const int64_t g_hi = p.cu_seqlens ? p.cu_seqlens[b + 1] : g_lo + p.N;
        const int64_t t_first = g_lo / AT_TILE, t_last = (g_hi + AT_TILE - 1) / AT_TILE;
        const int64_t t0 = t_first + (int64_t)split * p.tiles_per_split;
        const int64_t t1 = min(t_last, t0 + p.tiles_per_split);
        const int ntiles = (int)max((int64_t)0, t1 - t0);

        __syncthreads();                       // previous item: O read out, every MMA retired
        if (warp >= 2) {
            // ---- split this head's keys into the K-major B operand of S = Q Kp^T (zero rows beyond Ksel)
            const int units = KP * (dk / 8);
            for (int idx = threadIdx.x - 64; idx < units; idx += AT_THREADS - 64) {
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
                *reinterpret_cast<bf16x8*>(sK + (size_t)idx * 16) = hi;
                *reinterpret_cast<bf16x8*>(sK + KP_PLANE + (size_t)idx * 16) = lo;
            }
            fence_proxy_async_smem();
        }
        __syncthreads();

        if (warp == 0) {
            // ------------------------------------------------ producer
            for (int t = 0; t < ntiles; ++t, ++it) {
                const int64_t rt = t0 + t;
                const __nv_bfloat16* qsrc = p.planes + (rt * p.nkb + p.q_kb0 + j * chunks_per_head) * chunk_elems;
                const __nv_bfloat16* vsrc = p.planes + (rt * p.nkb + p.v_kb0 + j * chunks_per_head) * chunk_elems;
                if (lane == 0) DBG_STAMP(2, it, 0);
                mbar_wait(q_empty, (it & 1) ^ 1);
                if (lane == 0) DBG_STAMP(2, it, 1);
                if (lane == 0) {
                    mbar_expect_tx(q_full, 2 * QV_PLANE);
                    bulk_g2s(smem_u32(sQ), qsrc, QV_PLANE, q_full);
                    bulk_g2s(smem_u32(sQ) + QV_PLANE, qsrc + p.plane_stride, QV_PLANE, q_full);
                }
                if (MODE != 1) {
                    mbar_wait(v_empty, (it & 1) ^ 1);
                    if (lane == 0) DBG_STAMP(2, it, 2);
                    if (lane == 0) {
                        mbar_expect_tx(v_full, 2 * QV_PLANE);
                        bulk_g2s(smem_u32(sV), vsrc, QV_PLANE, v_full);
                        bulk_g2s(smem_u32(sV) + QV_PLANE, vsrc + p.plane_stride, QV_PLANE, v_full);
                    }
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)(dk >> 3) << 17) | (8u << 24);
            // N = 2 dk: the hi and lo planes of V are contiguous in the MN-major layout (QV_PLANE = dk/8 groups x SBO), so
            // P_hi^T [V_hi | V_lo] is ONE MMA whose A operand is read once (P^T V is bound by its shared-memory reads)
            const uint32_t idesc2w = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                     ((uint32_t)(dk >> 2) << 17) | (8u << 24);
            const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aP = smem_u32(sP), aV = smem_u32(sV);
            auto mma2 = [&](uint32_t itp, bool first) {
                // O[mblk] (+)= P(t)^T V(t): A = P planes (MN-major, M = key), B = V planes (MN-major, N = dv)
                if (lane == 0) DBG_STAMP(1, itp, 4);
                mbar_wait(p_full, itp & 1);
                if (lane == 0) DBG_STAMP(1, itp, 5);
                mbar_wait(v_full, itp & 1);
                if (lane == 0) DBG_STAMP(1, itp, 6);
                tc_fence_after();
                if (lane == 0) {
                    for (int mb = 0; mb < mblocks; ++mb) {
                        // accumulators of key block mb: columns [0, dk) = P_hi^T V_hi + P_lo^T V_hi, [dk, 2 dk) = P_hi^T V_lo
                        const uint32_t d_tmem = tmem_base + AT_O_COL + (uint32_t)(mb * 2 * dk);
                        for (int ks = 0; ks < AT_TILE / 16; ++ks) {
                            const uint32_t pa = aP + mb * (16 * 2048) + ks * 256, va = aV + ks * 256;
                            const uint64_t p_hi = make_smem_desc(pa, 128, 2048), p_lo = make_smem_desc(pa + P_PLANE, 128, 2048);
                            const uint64_t v_hi = make_smem_desc(va, 128, 2048), v_lo = make_smem_desc(va + QV_PLANE, 128, 2048);
                            (void)v_lo;
                            tc_mma_bf16(d_tmem, p_hi, v_hi, idesc2w, (first && ks == 0) ? 0u : 1u);
                            tc_mma_bf16(d_tmem, p_lo, v_hi, idesc2, 1u);
                        }
                    }
                    tc_commit(p_empty);
                    tc_commit(v_empty);
                    DBG_STAMP(1, itp, 7);
                }
                __syncwarp();
            };
            for (int t = 0; t < ntiles; ++t, ++it) {
                if (lane == 0) DBG_STAMP(1, it, 0);
                mbar_wait(q_full, it & 1);
                if (lane == 0) DBG_STAMP(1, it, 1);
                mbar_wait(s_empty, (it & 1) ^ 1);
                if (lane == 0) DBG_STAMP(1, it, 2);
                tc_fence_after();
                if (lane == 0) {
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t qa = aQ + ks * 2 * 2048, ka = aK + ks * 2 * (KP * 16);
                        const uint64_t q_hi = make_smem_desc(qa, 2048, 128), q_lo = make_smem_desc(qa + QV_PLANE, 2048, 128);
                        const uint64_t k_hi = make_smem_desc(ka, KP * 16, 128), k_lo = make_smem_desc(ka + KP_PLANE, KP * 16, 128);
                        tc_mma_bf16(tmem_base, q_lo, k_hi, idesc1, ks ? 1u : 0u);
                        tc_mma_bf16(tmem_base, q_hi, k_lo, idesc1, 1u);
                        tc_mma_bf16(tmem_base, q_hi, k_hi, idesc1, 1u);
                    }
                    tc_commit(q_empty);
                    tc_commit(s_full);
                    DBG_STAMP(1, it, 3);
                }
                __syncwarp();
                if (MODE != 1 && t > 0) mma2(it - 1, t == 1);
            }
            if (MODE != 1) {
                if (ntiles > 0) mma2(it - 1, ntiles == 1);
                if (lane == 0) tc_commit(o_full);
                __syncwarp();
            }
        } else {
            // ------------------------------------------------ softmax warps.  Thread = one query row of the tile;
            // the two warps of a TMEM lane quadrant split the key chunks and exchange (max, sum) through smem.
            const int quad = warp & 3, half = (warp - 2) >> 2;      // `half` = this warp's part (0 .. AT_PARTS-1) of the key chunks
            const int rr = quad * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
            const int nchunks = (KP + 31) / 32;
            const int cper = (nchunks + AT_PARTS - 1) / AT_PARTS;
            const int c_lo = min(nchunks, half * cper), c_hi = min(nchunks, c_lo + cper);
            constexpr int RED = AT_PARTS * 128;
            const uint32_t bar_id = 1 + quad;                       // named barrier of this quadrant's warp pair
            for (int t = 0; t < ntiles; ++t, ++it) {
                const int64_t g = (t0 + t) * AT_TILE + rr;
                const bool valid = g >= g_lo && g < g_hi;
                const int n = (int)(g - g_lo);
                const int64_t srow = ((int64_t)b * p.h + j) * p.N + n;      // row of the [B, h, N, *] statistics
                const bool stamp = warp == 2 && lane == 0;
                if (stamp) DBG_STAMP(0, it, 0);
                mbar_wait(s_full, it & 1);
                if (stamp) DBG_STAMP(0, it, 1);
                tc_fence_after();
                float mx = -INFINITY, inv = 0.f, mc = 0.f;
                if (MODE != 2) {
                    // ---- pass 1: row max over this warp's key chunks
                    for (int c = c_lo; c < c_hi; ++c) {
                        float v[32];
                        tc_ld32(lane_addr + (uint32_t)(c * 32), v);
                        if (c * 32 + 32 <= kvalid) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, v[e]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e) if (c * 32 + e < kvalid) mx = fmaxf(mx, v[e]);
                        }
                    }
                    sRed[half * 128 + rr] = mx;
                    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(AT_PARTS * 32) : "memory");
#pragma unroll
                    for (int q = 0; q < AT_PARTS; ++q) mx = fmaxf(mx, sRed[q * 128 + rr]);
                    if (stamp) DBG_STAMP(0, it, 2);
                    if (!valid) mx = 0.f;
                    // rows of a neighbouring bag / padding: exp2(s*c - inf) = 0 -> P = 0 exactly (never inf * 0 = NaN)
                    mc = valid ? mx * p.c_log2 : INFINITY;
                    // ---- pass 2: exp ONCE per score (the XU pipe is the limiter); MODE 0 parks it in TMEM over S
                    float sum = 0.f;
                    for (int c = c_lo; c < c_hi; ++c) {
                        float v[32];
                        tc_ld32(lane_addr + (uint32_t)(c * 32), v);
                        const bool full = c * 32 + 32 <= kvalid;
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float ee = ex2_approx(fmaf(v[e], p.c_log2, -mc));
                            v[e] = (full || c * 32 + e < kvalid) ? ee : 0.f;
                            sum += v[e];
                        }
                        if (MODE == 0) tc_st32(lane_addr + (uint32_t)(c * 32), v);
                    }
                    if (MODE == 0) tc_wait_st();
                    sRed[RED + half * 128 + rr] = sum;
                    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(AT_PARTS * 32) : "memory");
                    sum = 0.f;
#pragma unroll
                    for (int q = 0; q < AT_PARTS; ++q) sum += sRed[RED + q * 128 + rr];
                    if (stamp) DBG_STAMP(0, it, 3);
                    if (MODE == 1) {
                        // per-chunk partial statistics for pass 2 (raw-score max, sum of exp2 relative to it)
                        if (valid && half == 0) {
                            float* sp = p.stats_part + ((int64_t)kc * p.B * p.h * p.N + srow) * 2;
                            sp[0] = mx; sp[1] = sum;
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_empty);
                        continue;
                    }
                    inv = valid ? 1.f / sum : 0.f;
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
                }
                if (p.stats_out && valid && half == 0 && kc == 0) {
                    float* so = p.stats_out + srow * 2;
                    so[0] = mx * (p.c_log2 * 0.69314718055994530942f);     // max of the scaled scores (natural units)
                    so[1] = inv;
                }
                if (stamp) DBG_STAMP(0, it, 4);
                mbar_wait(p_empty, (it & 1) ^ 1);            // MMA of the previous tile has consumed P
                if (stamp) DBG_STAMP(0, it, 5);
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[32];
                    tc_ld32(lane_addr + (uint32_t)(c * 32), v);
                    if (MODE == 0) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] *= inv;       // exp(s - max) parked by the sum pass
                    } else {
                        const bool full = c * 32 + 32 <= kvalid;
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float pe = ex2_approx(fmaf(v[e], p.c_log2, -mc)) * inv;
                            v[e] = (full || c * 32 + e < kvalid) ? pe : 0.f;
                        }
                    }
                    if (p.P_out && valid) {
                        float* po = p.P_out + srow * p.Ksel + key0 + c * 32;
#pragma unroll
                        for (int e = 0; e < 32; ++e) if (c * 32 + e < kvalid) po[e] = v[e];
                    }
                    if (p.drop_p > 0.f && valid) {
                        const DrawKey dkey = rng_resolve(p.seed, p.offset);
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const uint64_t idx = ((uint64_t)srow * p.Ksel + key0 + c * 32 + e);
                            v[e] *= drop_keep_scale(dkey.seed, dkey.offset, idx, p.drop_p);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int kgp = c * 4 + u;
                        if (kgp * 8 < KP) {
                            // hi = truncated bf16 (exact), lo = bf16 of the exact remainder: integer/FMA pipes only
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t ua = __float_as_uint(v[u * 8 + 2 * q]), ub = __float_as_uint(v[u * 8 + 2 * q + 1]);
                                const float la = v[u * 8 + 2 * q] - __uint_as_float(ua & 0xFFFF0000u);
                                const float lb = v[u * 8 + 2 * q + 1] - __uint_as_float(ub & 0xFFFF0000u);
                                hw[q] = __byte_perm(ua, ub, 0x7632);
                                lw[q] = __byte_perm(__float_as_uint(la), __float_as_uint(lb), 0x7632);
                            }
                            *reinterpret_cast<uint4*>(sP + (size_t)kgp * 2048 + rr * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                            *reinterpret_cast<uint4*>(sP + P_PLANE + (size_t)kgp * 2048 + rr * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (stamp) DBG_STAMP(0, it, 6);
                if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(p_full); }
            }
            if (MODE != 1) {
                // ---- item epilogue: O (TMEM lane = key) -> this split's partial; each half takes one 128-key block
                mbar_wait(o_full, item_no & 1);
                tc_fence_after();
                for (int mb = half; mb < mblocks; mb += AT_PARTS) {
                    const int key = mb * 128 + rr;
                    for (int c = 0; c < dk / 32; ++c) {
                        float v[32];
                        float v2[32];
                        tc_ld32(lane_addr + AT_O_COL + (uint32_t)(mb * 2 * dk + c * 32), v);
                        tc_ld32(lane_addr + AT_O_COL + (uint32_t)(mb * 2 * dk + dk + c * 32), v2);
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] += v2[e];
                        if (key < kvalid) {
                            float* o = p.O_part + (((int64_t)split * p.B + b) * p.Ksel + key0 + key) * p.d + j * dk + c * 32;
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                float4 w = ntiles > 0 ? make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                                *reinterpret_cast<float4*>(o + e) = w;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
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
    return (size_t)2 * KP * 256 + (size_t)2 * KP * dk * 2 + (size_t)4 * AT_TILE * dk * 2 + 128 + 4096;
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
        if (KC > 256) continue;
        if (((KC + 127) / 128) * 2 * dk > 256) continue;      // O accumulators: 2 dk TMEM columns per 128-key block from column 256
        const size_t smem = attn_tc_smem(KC, dk);
        if (smem > 227 * 1024) continue;
        // the second 128-key MMA block addresses 16 key groups from its base: must stay inside the CTA's window
        if (KC > 128 && (size_t)(16 + 16) * 2048 + (size_t)KC * 256 > smem) continue;
        pl.nkc = (int)((Ksel + KC - 1) / KC); pl.KC = KC; pl.KP = KC; pl.smem = smem;
        break;
    }
    if (pl.nkc == 0) return pl;
    const int64_t tiles = (N + AT_TILE - 1) / AT_TILE + 1;         // a bag may straddle one extra row tile
    const int64_t per = B * h * pl.nkc;
    // Row tiles per work item: the persistent CTAs take items round-robin, so the launch lasts ceil(items / SMs) rounds of one
    // item each.  Pick the split that minimises rounds x (tiles per item + per-item overhead: key split, O read-out) plus the
    // cost of folding `splits` partial outputs, all in units of one tile (~5.5 us at 208 keys x 64) — e.g. 16 slides x 8 heads x
    // 80 tiles: 10 tiles per item = 7 rounds of 10.6 instead of 5 rounds of 16.6.
    const double fold_per_split = (double)B * (double)Ksel * (double)d * 4.0 / 4.0e6 / 5.5;
    double best = 1e30;
    int best_t = (int)tiles;
    for (int64_t t = 1; t <= tiles; ++t) {
        const int64_t s = (tiles + t - 1) / t;
        if (s > 64) continue;
        const int64_t rounds = (per * s + sm_count() - 1) / sm_count();
        const double cost = (double)rounds * ((double)t + 0.6) + fold_per_split * (double)s;
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
    if (pl.nkc > 1) bytes += (int64_t)pl.nkc * B * h * N * 2 * 4;
    return bytes;
}

// Same contract as snuffy_sparse_attn_fwd, but Q and V arrive as the split-bf16 planes the Q|V projection wrote
// (planes over [B*N, ldk] columns, Q at column q_col0, V at column v_col0; ldk, q_col0, v_col0 multiples of 32).
static int launch_attn_tc(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0, int64_t v_col0,
                          const float* Kp, int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d,
                          float dropout_p, uint64_t seed, uint64_t offset, float* O, float* P_out, float* stats_out,
                          void* workspace, int64_t workspace_bytes, const int64_t* cu_seqlens, cudaStream_t stream) {
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
    p.c_log2 = (float)(1.4426950408889634 / sqrt((double)p.dk));
    p.O_part = reinterpret_cast<float*>(workspace);
    p.stats_part = p.O_part + (int64_t)pl.splits * B * Ksel * d;
    p.P_out = P_out; p.stats_out = stats_out;
    p.drop_p = dropout_p; p.seed = seed; p.offset = offset;
    p.cu_seqlens = cu_seqlens;
    if (pl.nkc == 1) {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_tc_kernel<0>), (int)pl.smem));
        attn_tc_kernel<0><<<pl.grid, AT_THREADS, pl.smem, stream>>>(p);
    } else {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_tc_kernel<1>), (int)pl.smem));
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_tc_kernel<2>), (int)pl.smem));
        attn_tc_kernel<1><<<pl.grid, AT_THREADS, pl.smem, stream>>>(p);
        attn_tc_kernel<2><<<pl.grid, AT_THREADS, pl.smem, stream>>>(p);
    }
    launch_fold_partials(p.O_part, pl.splits, B * Ksel * d / 4, O, stream);
    return check_launch("snuffy_sparse_attn_tc_fwd", pl.nkc == 1 ? 2 : 3);
}

int snuffy_sparse_attn_tc_fwd(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0, int64_t v_col0,
                              const float* Kp, int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d,
                              float dropout_p, uint64_t seed, uint64_t offset, float* O, float* P_out, float* stats_out,
                              void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    return launch_attn_tc(qv_planes, plane_stride, ldk, q_col0, v_col0, Kp, B, N, Ksel, h, d, dropout_p, seed, offset, O, P_out,
                          stats_out, workspace, workspace_bytes, nullptr, stream);
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
                          workspace, workspace_bytes, cu_seqlens, stream);
}

#ifdef ATTN_DEBUG_TIMING
int snuffy_attn_debug_read(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_attn_dbg, sizeof(g_attn_dbg)) == cudaSuccess ? 0 : 1;
}
#endif

}  // extern "C"
#pragma GCC visibility pop
