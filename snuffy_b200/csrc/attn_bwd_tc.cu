// Backward of the Snuffy sparse attention on the 5th-gen tensor cores, ONE kernel per layer (autograd of snuffy.py:160-168).
//
//   forward   S_j = Q_j Kp_j^T / sqrt(dk),  P = softmax_keys(S),  P~ = dropout(P),  O_j = P~^T V_j
//   backward  dV = P~ dO            [N, dk]      (per 128-query tile, nothing to reduce)
//             G  = V dO^T           [N, Ksel]    = dP~
//             dS = P o (D o G - delta) / sqrt(dk),   delta[n] = sum_k P~[n, k] G[n, k]
//             dQ = dS Kp            [N, dk]
//             dKp = dS^T Q          [Ksel, dk]   accumulated over the row tiles in TMEM, transposed: dKp^T = Q^T dS  (the
//                                                forward's O^T = V^T P with Q in place of V and dS in place of P)
// The forward saves only (row max, 1 / row sum) per (bag, head, row): S and P are recomputed per tile and, like G and dS, never
// leave the SM.  The previous formulation (five block-diagonal GEMM launches + two row kernels per bag, S / G / P~ / dS through
// HBM: ~0.9 GB and ~0.35 ms per cfg2 bag) is kept as the fallback for shapes this kernel does not serve.
//
// Work item = (bag, head, range of 128-query tiles).  All five products are split-bf16 (hi / lo planes, fp32 accumulate):
//   warp 0      Q producer          warp 14  V producer          warp 15  K / dO producer (ONE shared-memory buffer holds the
//   warp 1      MMA issuer                                                  head's key planes or its dO planes: both are needed
//   warps 2-13  row warps (one query row per thread, three warps per TMEM lane quadrant owning contiguous 8-key groups)
// Per tile (TMEM: A = [0, KP) scores then G, D = dk columns for dV then dQ, C = KP columns for dKp^T):
//   S = Q Kp^T -> A | rows: S -> registers, P from the saved statistics, P~ planes -> smem | buffer <- dO | G = V dO^T -> A |
//   dV = P~ dO -> D | rows: delta (first pass over G), dS in place of P (second pass), dV out, dS planes -> smem (over P~) |
//   buffer <- Kp | dQ = dS Kp -> D, dKp^T += Q^T dS -> C | rows: dQ out.
#include "tc_ptx.cuh"

namespace snuffy {

void launch_fold_partials(const float* part, int splits, int64_t n4, float* out, cudaStream_t stream);

constexpr int AB_PARTS = 3;
constexpr int AB_SOFT = 128 * AB_PARTS;
constexpr int AB_VWARP = 2 + 4 * AB_PARTS, AB_KWARP = AB_VWARP + 1;
constexpr int AB_THREADS = 64 + AB_SOFT + 64;
constexpr int AB_TILE = 128;
constexpr int AB_VG = 10;                 // 8-key groups one thread can own: KP <= 224 -> 28 groups over 3 parts

__device__ __forceinline__ float ex2_approx_b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t fast_div_b(uint32_t n, int d, uint32_t mul, uint32_t shr) {
    return d == 1 ? n : __umulhi(n, mul) >> shr;
}

#ifdef ATTN_DEBUG_TIMING
__device__ long long g_attnb_dbg[64 * 8];          // clock64 stamps of CTA 0, row warp 2 lane 0: [tile < 64][8 stamps]
#define DBGB(t, k) do { if (blockIdx.x == 0 && warp == 2 && lane == 0 && (t) < 64) g_attnb_dbg[(t) * 8 + (k)] = clock64(); } while (0)
#else
#define DBGB(t, k) do { } while (0)
#endif

struct AttnBwdParams {
    const __nv_bfloat16* planes; int64_t plane_stride;   // the forward's Q|V planes over [rows, ldk]
    int nkb, q_kb0, v_kb0;
    const float* Kp; const float* dO;                    // [B*Ksel, d] fp32
    __nv_bfloat16* k_planes; __nv_bfloat16* do_planes;   // workspace: per (bag, head) [hi | lo][dk/8][KP][8]
    const float* stats;                                  // [B, h, N, 2] (row max of the scaled scores, 1 / row sum)
    int B, N, Ksel, KP, h, dk, d;
    int splits, tiles_per_split;
    uint32_t div_mul[2], div_shr[2];                     // item -> (split, head, bag)
    uint32_t d_col, c_col;                               // TMEM columns of D and C
    float c_log2, scale;                                 // log2(e) / sqrt(dk), 1 / sqrt(dk)
    float drop_p; const uint8_t* drop_mask;              // keep bits as the forward drew them: [B, h, N, ceil(Ksel / 8)]
    float* dqv;                                          // [B*N, 2d]: dQ | dV
    float* dkp_part;                                     // [splits][B*Ksel][d]
};

// fp32 [B*Ksel, d] -> per (bag, head) split-bf16 planes [hi | lo][dk/8][KP][8] (zero rows beyond Ksel): blockIdx.y = 0 keys, 1 dO
__global__ void __launch_bounds__(256)
attn_bwd_planes_kernel(const AttnBwdParams p) {
    const int blk = blockIdx.x, j = blk % p.h, b = blk / p.h;
    const float* srcm = blockIdx.y ? p.dO : p.Kp;
    const int KPS = p.KP + 1;                                // padded group stride (see the kernel below)
    const size_t plane = (size_t)KPS * p.dk;
    __nv_bfloat16* out = (blockIdx.y ? p.do_planes : p.k_planes) + (size_t)blk * 2 * plane;
    const int units = KPS * (p.dk / 8);
    for (int idx = threadIdx.x; idx < units; idx += blockDim.x) {
        const int key = idx % KPS, kg = idx / KPS;
        bf16x8 hi, lo;
        if (key < p.Ksel) {
            const float* src = srcm + ((int64_t)b * p.Ksel + key) * p.d + j * p.dk + kg * 8;
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

// hi = truncated bf16 (exact), lo = bf16 of the exact remainder, for 8 values -> two 16-byte units
__device__ __forceinline__ void split_store8(const float (&w)[8], unsigned char* hi_dst, unsigned char* lo_dst) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t ua = __float_as_uint(w[2 * q]), ub = __float_as_uint(w[2 * q + 1]);
        const float la = w[2 * q] - __uint_as_float(ua & 0xFFFF0000u);
        const float lb = w[2 * q + 1] - __uint_as_float(ub & 0xFFFF0000u);
        hw[q] = __byte_perm(ua, ub, 0x7632);
        lw[q] = __byte_perm(__float_as_uint(la), __float_as_uint(lb), 0x7632);
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const AttnBwdParams p) {
    extern __shared__ __align__(1024) unsigned char ab_smem[];
    const int dk = p.dk, KP = p.KP;
    const uint32_t P_PLANE = (uint32_t)KP * 256u;            // [KP/8 key groups][128 queries][8 keys] bf16
    const uint32_t QV_PLANE = (uint32_t)AB_TILE * dk * 2u;   // [dk/8][128][8] bf16
    // key / dO planes [dk/8][KP + 1][8] bf16: the 8-column groups are (KP + 1) * 16 bytes apart, an odd number of 16-byte units,
    // so that the MN-major reads of dV = P~ dO / dQ = dS Kp (one 16-byte unit from each of the dk / 8 groups) spread over the banks
    const uint32_t KPS = (uint32_t)KP + 1u;
    const uint32_t KP_PLANE = KPS * dk * 2u;
    // order matters: the M = 128 stacked A operand of dKp^T = Q^T dS reads 16 groups of 2 KB from the start of sQ
    unsigned char* sKD = ab_smem;
    unsigned char* sQ = sKD + 2 * KP_PLANE;
    unsigned char* sV = sQ + 2 * QV_PLANE;
    unsigned char* sPS = sV + 2 * QV_PLANE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPS + 2 * P_PLANE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    float* sRed0 = reinterpret_cast<float*>(bars + 17);      // [tile parity][AB_PARTS][128] (one exchange barrier per tile: see attn_tc.cu)
    const uint32_t b0 = smem_u32(bars);
    const uint32_t q_full = b0, q_empty = b0 + 8, v_full = b0 + 16, v_empty = b0 + 24, kd_full = b0 + 32, kd_empty = b0 + 40,
                   a_full = b0 + 48, a_free = b0 + 56, ps_full = b0 + 64, ps_empty = b0 + 72, d_full = b0 + 80, d_free = b0 + 88,
                   c_full = b0 + 96, c_free = b0 + 104;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(kd_full, 1); mbar_init(kd_empty, 1); mbar_init(a_full, 1); mbar_init(a_free, 4 * AB_PARTS);
        mbar_init(ps_full, 4 * AB_PARTS); mbar_init(ps_empty, 1); mbar_init(d_full, 1); mbar_init(d_free, 4 * AB_PARTS);
        mbar_init(c_full, 1); mbar_init(c_free, 4 * AB_PARTS);
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

    const int ksteps = dk / 16, kksteps = KP / 16;
    const bool stacked = 2 * dk <= 128;
    const int chunks_per_head = dk / PLANE_KB;
    const int64_t chunk_elems = (int64_t)AB_TILE * PLANE_KB;
    const int items = p.B * p.h * p.splits;
    uint32_t it = 0, item_no = 0, kdc = 0;        // tiles / items / K-dO buffer fills seen so far (every role counts the same)

    for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_no) {
        const uint32_t r1 = fast_div_b((uint32_t)item, p.splits, p.div_mul[0], p.div_shr[0]);
        const int b = (int)fast_div_b(r1, p.h, p.div_mul[1], p.div_shr[1]);
        const int split = item - (int)r1 * p.splits, j = (int)r1 - b * p.h;
        const int64_t g_lo = (int64_t)b * p.N, g_hi = g_lo + p.N;
        const int64_t t_first = g_lo / AB_TILE, t_last = (g_hi + AB_TILE - 1) / AB_TILE;
        const int64_t t0 = t_first + (int64_t)split * p.tiles_per_split;
        const int64_t t1 = min(t_last, t0 + p.tiles_per_split);
        const int ntiles = (int)max((int64_t)0, t1 - t0);
        const size_t head_planes = ((size_t)b * p.h + j) * 2 * ((size_t)KPS * dk);

        if (warp == 0) {
            // ------------------------------------------------ Q producer
            for (int t = 0; t < ntiles; ++t, ++it) {
                const __nv_bfloat16* qsrc = p.planes + ((t0 + t) * p.nkb + p.q_kb0 + j * chunks_per_head) * chunk_elems;
                mbar_wait(q_empty, (it & 1) ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(q_full, 2 * QV_PLANE);
                    bulk_g2s(smem_u32(sQ), qsrc, QV_PLANE, q_full);
                    bulk_g2s(smem_u32(sQ) + QV_PLANE, qsrc + p.plane_stride, QV_PLANE, q_full);
                }
                __syncwarp();
            }
        } else if (warp == AB_VWARP) {
            // ------------------------------------------------ V producer
            for (int t = 0; t < ntiles; ++t, ++it) {
                const __nv_bfloat16* vsrc = p.planes + ((t0 + t) * p.nkb + p.v_kb0 + j * chunks_per_head) * chunk_elems;
                mbar_wait(v_empty, (it & 1) ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(v_full, 2 * QV_PLANE);
                    bulk_g2s(smem_u32(sV), vsrc, QV_PLANE, v_full);
                    bulk_g2s(smem_u32(sV) + QV_PLANE, vsrc + p.plane_stride, QV_PLANE, v_full);
                }
                __syncwarp();
            }
        } else if (warp == AB_KWARP) {
            // ------------------------------------------------ key / dO producer: Kp, then per tile dO, Kp
            const __nv_bfloat16* ksrc = p.k_planes + head_planes;
            const __nv_bfloat16* dsrc = p.do_planes + head_planes;
            for (int f = 0; f < 1 + 2 * ntiles; ++f, ++kdc) {
                const __nv_bfloat16* src = (f & 1) ? dsrc : ksrc;
                mbar_wait(kd_empty, (kdc & 1) ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(kd_full, 2 * KP_PLANE);
                    bulk_g2s(smem_u32(sKD), src, KP_PLANE, kd_full);
                    bulk_g2s(smem_u32(sKD) + KP_PLANE, src + (size_t)KPS * dk, KP_PLANE, kd_full);
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);              // K-major x K-major, N = KP
            const uint32_t idescD = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(dk >> 3) << 17) | (8u << 24); // K-major x MN-major, N = dk
            const uint32_t idescC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);
            // descriptors of k-step 0 (csrc/tc_ptx.cuh: (address, LBO, SBO)); a k-step advances the start-address field
            const uint64_t q_hi = make_smem_desc(smem_u32(sQ), 2048, 128), q_lo = make_smem_desc(smem_u32(sQ) + QV_PLANE, 2048, 128);
            const uint64_t v_hi = make_smem_desc(smem_u32(sV), 2048, 128), v_lo = make_smem_desc(smem_u32(sV) + QV_PLANE, 2048, 128);
            const uint64_t kd_hi = make_smem_desc(smem_u32(sKD), KPS * 16, 128), kd_lo = make_smem_desc(smem_u32(sKD) + KP_PLANE, KPS * 16, 128);
            // the same key / dO planes as the MN-major B operand [K = key, N = dv] of dQ = dS Kp and dV = P~ dO
            const uint64_t kdn_hi = make_smem_desc(smem_u32(sKD), 128, KPS * 16), kdn_lo = make_smem_desc(smem_u32(sKD) + KP_PLANE, 128, KPS * 16);
            // P~ / dS planes: K-major A operand [M = query, K = key] of dV / dQ, MN-major B operand [K = query, N = key] of dKp^T
            const uint64_t ps_hi = make_smem_desc(smem_u32(sPS), 2048, 128), ps_lo = make_smem_desc(smem_u32(sPS) + P_PLANE, 2048, 128);
            const uint64_t psn_hi = make_smem_desc(smem_u32(sPS), 128, 2048), psn_lo = make_smem_desc(smem_u32(sPS) + P_PLANE, 128, 2048);
            const uint64_t qn_hi = make_smem_desc(smem_u32(sQ), 128, 2048), qn_lo = make_smem_desc(smem_u32(sQ) + QV_PLANE, 128, 2048);
            const uint64_t qv_step = (2 * 2048) >> 4, kd_step = (uint64_t)(2 * KPS * 16) >> 4, mn_step = 256 >> 4;
            const uint32_t tA = tmem_base, tD = tmem_base + p.d_col, tC = tmem_base + p.c_col;

            mbar_wait(kd_full, kdc & 1); ++kdc;                                   // this head's key planes
            if (ntiles == 0) {
                if (item_no > 0) mbar_wait(c_free, (item_no - 1) & 1);            // keep c_full one phase apart from its consumers
                if (lane == 0) mbar_arrive(kd_empty);
            }
#pragma unroll 1
            for (int t = 0; t < ntiles; ++t, ++it) {
                // ---- S = Q Kp^T -> A
                mbar_wait(q_full, it & 1);
                mbar_wait(a_free, 1);                                             // fill 2 it: G of the previous tile consumed
                tc_fence_after();
                if (lane == 0) {
#pragma unroll 1
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t qo = qv_step * ks, ko = kd_step * ks;
                        tc_mma_bf16(tA, q_lo + qo, kd_hi + ko, idescS, ks ? 1u : 0u);
                        tc_mma_bf16(tA, q_hi + qo, kd_lo + ko, idescS, 1u);
                        tc_mma_bf16(tA, q_hi + qo, kd_hi + ko, idescS, 1u);
                    }
                    tc_commit(a_full);
                    tc_commit(kd_empty);                                          // keys -> dO
                }
                __syncwarp();
                // ---- G = V dO^T -> A (after the rows have taken S into registers)
                mbar_wait(kd_full, kdc & 1); ++kdc;
                mbar_wait(v_full, it & 1);
                mbar_wait(a_free, 0);                                             // fill 2 it + 1
                tc_fence_after();
                if (lane == 0) {
#pragma unroll 1
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t vo = qv_step * ks, ko = kd_step * ks;
                        tc_mma_bf16(tA, v_lo + vo, kd_hi + ko, idescS, ks ? 1u : 0u);
                        tc_mma_bf16(tA, v_hi + vo, kd_lo + ko, idescS, 1u);
                        tc_mma_bf16(tA, v_hi + vo, kd_hi + ko, idescS, 1u);
                    }
                    tc_commit(a_full);
                    tc_commit(v_empty);
                }
                __syncwarp();
                // ---- dV = P~ dO -> D
                mbar_wait(ps_full, 0);
                mbar_wait(d_free, 1);                                             // fill 2 it: dQ of the previous tile read out
                tc_fence_after();
                if (lane == 0) {
#pragma unroll 1
                    for (int ks = 0; ks < kksteps; ++ks) {         // rolled: any unrolling of this loop crashes nvcc 12.9 (cicc segfault)
                        const uint64_t po = qv_step * ks, ko = mn_step * ks;
                        tc_mma_bf16(tD, ps_lo + po, kdn_hi + ko, idescD, ks ? 1u : 0u);
                        tc_mma_bf16(tD, ps_hi + po, kdn_lo + ko, idescD, 1u);
                        tc_mma_bf16(tD, ps_hi + po, kdn_hi + ko, idescD, 1u);
                    }
                    tc_commit(d_full);
                    tc_commit(ps_empty);
                    tc_commit(kd_empty);                                          // dO -> keys
                }
                __syncwarp();
                // ---- dQ = dS Kp -> D,  dKp^T += Q^T dS -> C
                mbar_wait(kd_full, kdc & 1); ++kdc;
                mbar_wait(ps_full, 1);
                mbar_wait(d_free, 0);                                             // fill 2 it + 1: dV read out
                if (t == 0 && item_no > 0) mbar_wait(c_free, (item_no - 1) & 1);  // the previous item's dKp has been read out
                tc_fence_after();
                if (lane == 0) {
#pragma unroll 1
                    for (int ks = 0; ks < kksteps; ++ks) {         // rolled: any unrolling of this loop crashes nvcc 12.9 (cicc segfault)
                        const uint64_t po = qv_step * ks, ko = mn_step * ks;
                        tc_mma_bf16(tD, ps_lo + po, kdn_hi + ko, idescD, ks ? 1u : 0u);
                        tc_mma_bf16(tD, ps_hi + po, kdn_lo + ko, idescD, 1u);
                        tc_mma_bf16(tD, ps_hi + po, kdn_hi + ko, idescD, 1u);
                    }
                    tc_commit(d_full);
#pragma unroll 1
                    for (int ks = 0; ks < AB_TILE / 16; ++ks) {
                        const uint64_t o = mn_step * ks;
                        const uint32_t acc0 = (t == 0 && ks == 0) ? 0u : 1u;
                        if (stacked) {           // lanes [0, dk): Q_hi^T (dS_lo + dS_hi), lanes [dk, 2 dk): Q_lo^T (dS_lo + dS_hi)
                            tc_mma_bf16(tC, qn_hi + o, psn_lo + o, idescC, acc0);
                            tc_mma_bf16(tC, qn_hi + o, psn_hi + o, idescC, 1u);
                        } else {
                            tc_mma_bf16(tC, qn_lo + o, psn_hi + o, idescC, acc0);
                            tc_mma_bf16(tC, qn_hi + o, psn_lo + o, idescC, 1u);
                            tc_mma_bf16(tC, qn_hi + o, psn_hi + o, idescC, 1u);
                        }
                    }
                    tc_commit(ps_empty);
                    tc_commit(q_empty);
                    if (t == ntiles - 1) tc_commit(kd_empty);                     // keys -> the next item's keys
                }
                __syncwarp();
            }
            if (lane == 0) tc_commit(c_full);
            __syncwarp();
        } else {
            // ------------------------------------------------ row warps
            const int quad = warp & 3, part = (warp - 2) >> 2;
            const int rr = quad * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
            const int ngroups = KP / 8, gbase = ngroups / AB_PARTS, grem = ngroups % AB_PARTS;
            const int g0 = part * gbase + min(part, grem), ng = gbase + (part < grem ? 1 : 0);
            const uint32_t bar_id = 1 + quad;
            // this thread's share of the dk output columns of dV / dQ (8-column groups)
            const int cgroups = dk / 8, cbase = cgroups / AB_PARTS, crem = cgroups % AB_PARTS;
            const int c0 = part * cbase + min(part, crem), nc = cbase + (part < crem ? 1 : 0);
            const int mbytes = (p.Ksel + 7) / 8;
            const float keep_scale = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
            for (int t = 0; t < ntiles; ++t, ++it) {
                const int64_t g = (t0 + t) * AB_TILE + rr;
                const bool valid = g >= g_lo && g < g_hi;
                const int n = (int)(g - g_lo);
                const int64_t srow = ((int64_t)b * p.h + j) * p.N + n;
                float2 st = make_float2(0.f, 0.f);
                if (valid) st = __ldg(reinterpret_cast<const float2*>(p.stats) + srow);
                // exp(s / sqrt(dk) - max) / sum = exp2(s c_log2 - max log2 e) * inv; other bags' rows / padding: exactly 0
                const float mc = valid ? st.x * 1.4426950408889634f : INFINITY, inv = valid ? st.y : 0.f;
                // keep bits of this thread's keys (byte = one 8-key group), fetched while the S MMAs run
                uint32_t kbits[(AB_VG + 3) / 4] = {};
                if (DROP && valid) {
                    const uint8_t* mrow = p.drop_mask + srow * mbytes + g0;
#pragma unroll
                    for (int gi = 0; gi < AB_VG; ++gi)
                        if (gi < ng && g0 + gi < mbytes) kbits[gi >> 2] |= (uint32_t)__ldg(mrow + gi) << (8 * (gi & 3));
                }
                // ---- (a) scores -> registers -> P (kept) -> P~ planes
                mbar_wait(a_full, 0);
                tc_fence_after();
                DBGB(it, 0);
                float v[AB_VG * 8];
                tc_ld32(lane_addr + (uint32_t)(g0 * 8), *reinterpret_cast<float(*)[32]>(v));
                if (ng > 4) tc_ld32(lane_addr + (uint32_t)(g0 * 8 + 32), *reinterpret_cast<float(*)[32]>(v + 32));
                if (ng > 8) tc_ld8(lane_addr + (uint32_t)(g0 * 8 + 64), *reinterpret_cast<float(*)[8]>(v + 64));
                if (ng > 9) tc_ld8(lane_addr + (uint32_t)(g0 * 8 + 72), *reinterpret_cast<float(*)[8]>(v + 72));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_free);
                // padding keys (Q . 0 = 0) get -inf once: exp2 turns them into exact zeros, the loops need no masks
                if ((g0 + ng) * 8 > p.Ksel) {
#pragma unroll
                    for (int gi = 0; gi < AB_VG; ++gi) {
                        if (gi < ng && (g0 + gi) * 8 + 8 > p.Ksel) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) if ((g0 + gi) * 8 + e >= p.Ksel) v[gi * 8 + e] = -INFINITY;
                        }
                    }
                }
#pragma unroll
                for (int gi = 0; gi < AB_VG; ++gi) {
                    if (gi < ng) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[gi * 8 + e] = ex2_approx_b(fmaf(v[gi * 8 + e], p.c_log2, -mc)) * inv;
                    }
                }
                mbar_wait(ps_empty, 1);                          // fill 2 it: dKp^T of the previous tile has consumed dS
#pragma unroll
                for (int gi = 0; gi < AB_VG; ++gi) {
                    if (gi < ng) {
                        const int kgp = g0 + gi;
                        float w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            w[e] = v[gi * 8 + e];
                            if (DROP) w[e] = (kbits[gi >> 2] >> (8 * (gi & 3) + e)) & 1u ? w[e] * keep_scale : 0.f;
                        }
                        split_store8(w, sPS + (size_t)kgp * 2048 + rr * 16, sPS + P_PLANE + (size_t)kgp * 2048 + rr * 16);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(ps_full);
                DBGB(it, 1);
                // ---- (c1) delta[n] = sum_k P~[n, k] G[n, k]: first pass over G (this thread's keys), one exchange between the parts.
                // (Taken from the very G values that enter dS below, so that a saturated row cancels exactly.)
                mbar_wait(a_full, 1);
                tc_fence_after();
                DBGB(it, 2);
                float dpart = 0.f;
#pragma unroll
                for (int gi = 0; gi < AB_VG; ++gi) {
                    if (gi < ng) {
                        float gv[8];
                        tc_ld8(lane_addr + (uint32_t)((g0 + gi) * 8), gv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float pt = v[gi * 8 + e];
                            if (DROP) pt = (kbits[gi >> 2] >> (8 * (gi & 3) + e)) & 1u ? pt * keep_scale : 0.f;
                            dpart = fmaf(pt, gv[e], dpart);
                        }
                    }
                }
                float* sRed = sRed0 + (it & 1) * AB_PARTS * 128;
                sRed[part * 128 + rr] = dpart;
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(AB_PARTS * 32) : "memory");
                float delta = 0.f;
#pragma unroll
                for (int q = 0; q < AB_PARTS; ++q) delta += sRed[q * 128 + rr];
                DBGB(it, 3);
                // ---- (c2) dS = P o (D o G - delta) / sqrt(dk), in place of P (second pass over G)
#pragma unroll
                for (int gi = 0; gi < AB_VG; ++gi) {
                    if (gi < ng) {
                        float gv[8];
                        tc_ld8(lane_addr + (uint32_t)((g0 + gi) * 8), gv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float dp = gv[e];
                            if (DROP) dp = (kbits[gi >> 2] >> (8 * (gi & 3) + e)) & 1u ? dp * keep_scale : 0.f;
                            v[gi * 8 + e] *= (dp - delta) * p.scale;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_free);
                DBGB(it, 4);
                // ---- (b) dV out (its MMAs ran beside the two passes above)
                float* drow = p.dqv + g * (2 * (int64_t)p.d) + j * dk;
                mbar_wait(d_full, 0);
                tc_fence_after();
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    if (ci < nc) {
                        float o[8];
                        tc_ld8(lane_addr + p.d_col + (uint32_t)((c0 + ci) * 8), o);
                        if (valid) st_global_v8(drow + p.d + (c0 + ci) * 8, o);
                    }
                }
                for (int ci = 3; ci < nc; ++ci) {                // head sizes above 64: the rest, rolled
                    float o[8];
                    tc_ld8(lane_addr + p.d_col + (uint32_t)((c0 + ci) * 8), o);
                    if (valid) st_global_v8(drow + p.d + (c0 + ci) * 8, o);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_free);
                // ---- dS planes over P~ (dV has consumed it)
                mbar_wait(ps_empty, 0);                          // fill 2 it + 1
#pragma unroll
                for (int gi = 0; gi < AB_VG; ++gi) {
                    if (gi < ng) {
                        const int kgp = g0 + gi;
                        split_store8(*reinterpret_cast<const float(*)[8]>(v + gi * 8), sPS + (size_t)kgp * 2048 + rr * 16,
                                     sPS + P_PLANE + (size_t)kgp * 2048 + rr * 16);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(ps_full);
                DBGB(it, 5);
                // ---- (d) dQ out
                mbar_wait(d_full, 1);
                tc_fence_after();
                DBGB(it, 6);
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    if (ci < nc) {
                        float o[8];
                        tc_ld8(lane_addr + p.d_col + (uint32_t)((c0 + ci) * 8), o);
                        if (valid) st_global_v8(drow + (c0 + ci) * 8, o);
                    }
                }
                for (int ci = 3; ci < nc; ++ci) {
                    float o[8];
                    tc_ld8(lane_addr + p.d_col + (uint32_t)((c0 + ci) * 8), o);
                    if (valid) st_global_v8(drow + (c0 + ci) * 8, o);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_free);
                DBGB(it, 7);
            }
            // ---- item epilogue: dKp^T (TMEM lane = dv, column = key) -> this split's partial (like the forward's O^T read-out)
            mbar_wait(c_full, item_no & 1);
            tc_fence_after();
            asm volatile("bar.sync %0, %1;" ::"r"(5), "r"(AB_SOFT) : "memory");      // as in attn_tc.cu: the order c_full implies, made visible
            {
                const int L = rr;
                float* stage = reinterpret_cast<float*>(sPS);        // [key][dk] fp32; every MMA that read sPS has retired
                float* dst = p.dkp_part + (((int64_t)split * p.B + b) * p.Ksel) * p.d + j * dk;
                const bool keep = ntiles > 0;
                const int nch = (p.Ksel + 31) / 32;
                if (stacked) {
                    if (L < dk) {
                        for (int c = part; c < nch; c += AB_PARTS) {
                            float o[32];
                            tc_ld32(lane_addr + p.c_col + (uint32_t)(c * 32), o);
#pragma unroll
                            for (int e = 0; e < 32; ++e) stage[(c * 32 + e) * dk + L] = o[e];
                        }
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(5), "r"(AB_SOFT) : "memory");
                    if (L >= dk && L < 2 * dk) {
                        for (int c = part; c < nch; c += AB_PARTS) {
                            float o[32];
                            tc_ld32(lane_addr + p.c_col + (uint32_t)(c * 32), o);
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = keep ? o[e] + stage[(c * 32 + e) * dk + (L - dk)] : 0.f;
                            float* dg = dst + (int64_t)(c * 32) * p.d + (L - dk);
#pragma unroll
                            for (int e = 0; e < 32; ++e) if (c * 32 + e < p.Ksel) dg[(int64_t)e * p.d] = o[e];
                        }
                    }
                } else if (L < dk) {
                    for (int c = part; c < nch; c += AB_PARTS) {
                        float o[32];
                        tc_ld32(lane_addr + p.c_col + (uint32_t)(c * 32), o);
                        float* dg = dst + (int64_t)(c * 32) * p.d + L;
#pragma unroll
                        for (int e = 0; e < 32; ++e) if (c * 32 + e < p.Ksel) dg[(int64_t)e * p.d] = keep ? o[e] : 0.f;
                    }
                }
                tc_fence_before();
                asm volatile("bar.sync %0, %1;" ::"r"(5), "r"(AB_SOFT) : "memory");       // the staging area is the P~ / dS planes
                if (lane == 0) mbar_arrive(c_free);
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

struct AttnBwdPlan { int KP, splits, tiles_per_split, grid; uint32_t d_col, c_col; size_t smem; bool ok; };

static void fast_div_setup_b(int d, uint32_t& mul, uint32_t& shr) {
    if (d <= 1) { mul = 0; shr = 0; return; }
    uint32_t lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;
    const uint64_t pw = 31 + lg;
    mul = (uint32_t)((((uint64_t)1 << pw) + (uint64_t)d - 1) / (uint64_t)d);
    shr = (uint32_t)(pw - 32);
}

static AttnBwdPlan plan_attn_bwd(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d) {
    AttnBwdPlan pl{};
    if (h <= 0 || d % h || B < 1 || N < 1) return pl;
    const int dk = (int)(d / h);
    if (dk % 32 || dk > 128 || Ksel < 1 || Ksel > 224) return pl;
    const int KP = (int)((Ksel + 15) / 16 * 16);
    pl.KP = KP;
    pl.d_col = (uint32_t)((KP + 31) / 32 * 32);
    pl.c_col = pl.d_col + (uint32_t)((dk + 31) / 32 * 32);
    if (pl.c_col + (uint32_t)KP > 512u) return pl;                     // TMEM: scores / G, dV / dQ, dKp^T
    size_t stage = (size_t)2 * KP * 256;                               // P~ / dS planes, also the fp32 staging of the dKp read-out
    if (2 * dk <= 128) { const size_t st = (((size_t)(KP + 32) * dk * 4) + 127) & ~(size_t)127; if (st > stage) stage = st; }
    pl.smem = (size_t)2 * (KP + 1) * dk * 2 + (size_t)4 * AB_TILE * dk * 2 + stage + 17 * 8 + 16 + 2 * AB_PARTS * 128 * 4;
    if (stage != (size_t)2 * KP * 256) return pl;                      // keep the carve-up of the kernel (bars follow 2 P planes)
    if (pl.smem > 227 * 1024) return pl;
    const int64_t tiles = (N + AB_TILE - 1) / AB_TILE + 1;
    const int64_t per = B * h;
    double best = 1e30;
    int best_t = (int)tiles;
    for (int64_t t = 1; t <= tiles; ++t) {
        const int64_t s = (tiles + t - 1) / t;
        if (s > 64) continue;
        const int64_t rounds = (per * s + sm_count() - 1) / sm_count();
        const double cost = (double)rounds * ((double)t + 1.0) + 0.05 * (double)s;     // per-item overhead ~ one tile; fold per split
        if (cost < best - 1e-9) { best = cost; best_t = (int)t; }
    }
    pl.tiles_per_split = best_t;
    pl.splits = (int)((tiles + best_t - 1) / best_t);
    const int64_t items = per * pl.splits;
    pl.grid = (int)(items < sm_count() ? items : sm_count());
    pl.ok = true;
    return pl;
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// bytes of workspace for snuffy_sparse_attn_bwd_tc, or -1 when the shape is not served (then use the block-diagonal GEMM
// formulation: snuffy_gemm_tc_blockdiag + snuffy_attn_seg_bwd).  Needs dk % 32 == 0, dk <= 128, Ksel <= 224.
int64_t snuffy_sparse_attn_bwd_tc_workspace(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d) {
    const AttnBwdPlan pl = plan_attn_bwd(B, N, Ksel, h, d);
    if (!pl.ok) return -1;
    return (int64_t)pl.splits * B * Ksel * d * 4 + (int64_t)2 * B * (pl.KP + 1) * d * 4 + 256;
}

// Backward of snuffy_sparse_attn_tc_fwd in one kernel: dQV [B*N, 2d] (dQ | dV), dKp [B*Ksel, d].
// qv_planes: the forward's Q|V planes; Kp, dO: fp32 [B*Ksel, d]; stats [B, h, N, 2] from the forward (stats_out);
// dropout_p > 0: drop_mask = the keep bits snuffy_sparse_attn_tc_fwd wrote for the same draw.
int snuffy_sparse_attn_bwd_tc(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0, int64_t v_col0,
                              const float* Kp, const float* dO, const float* stats, int64_t B, int64_t N, int64_t Ksel,
                              int64_t h, int64_t d, float dropout_p, const uint8_t* drop_mask, float* dQV, float* dKp,
                              void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    SNUFFY_REQUIRE(qv_planes && Kp && dO && stats && dQV && dKp && workspace, "snuffy_sparse_attn_bwd_tc: null pointer");
    const AttnBwdPlan pl = plan_attn_bwd(B, N, Ksel, h, d);
    SNUFFY_REQUIRE(pl.ok, "snuffy_sparse_attn_bwd_tc: unsupported shape (h=%lld d=%lld Ksel=%lld)", (long long)h, (long long)d,
                   (long long)Ksel);
    SNUFFY_REQUIRE(ldk % 32 == 0 && q_col0 % 32 == 0 && v_col0 % 32 == 0 && q_col0 + d <= ldk && v_col0 + d <= ldk,
                   "snuffy_sparse_attn_bwd_tc: Q/V column ranges must be 32-aligned inside the planes");
    SNUFFY_REQUIRE(workspace_bytes >= snuffy_sparse_attn_bwd_tc_workspace(B, N, Ksel, h, d),
                   "snuffy_sparse_attn_bwd_tc: workspace too small");
    SNUFFY_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "snuffy_sparse_attn_bwd_tc: dropout_p out of range");
    SNUFFY_REQUIRE(dropout_p == 0.f || drop_mask, "snuffy_sparse_attn_bwd_tc: dropout needs the forward's keep-bit mask");
    SNUFFY_REQUIRE((uintptr_t)Kp % 16 == 0 && (uintptr_t)dO % 16 == 0 && (uintptr_t)dQV % 16 == 0 && (uintptr_t)dKp % 16 == 0 &&
                   (uintptr_t)workspace % 16 == 0 && (uintptr_t)stats % 8 == 0 && d % 8 == 0,
                   "snuffy_sparse_attn_bwd_tc: pointers must be 16-byte aligned");
    AttnBwdParams p{};
    p.planes = reinterpret_cast<const __nv_bfloat16*>(qv_planes); p.plane_stride = plane_stride;
    p.nkb = (int)(ldk / 32); p.q_kb0 = (int)(q_col0 / 32); p.v_kb0 = (int)(v_col0 / 32);
    p.Kp = Kp; p.dO = dO; p.stats = stats;
    p.B = (int)B; p.N = (int)N; p.Ksel = (int)Ksel; p.KP = pl.KP; p.h = (int)h; p.dk = (int)(d / h); p.d = (int)d;
    p.splits = pl.splits; p.tiles_per_split = pl.tiles_per_split;
    fast_div_setup_b(p.splits, p.div_mul[0], p.div_shr[0]);
    fast_div_setup_b(p.h, p.div_mul[1], p.div_shr[1]);
    p.d_col = pl.d_col; p.c_col = pl.c_col;
    p.c_log2 = (float)(1.4426950408889634 / sqrt((double)p.dk));
    p.scale = (float)(1.0 / sqrt((double)p.dk));
    p.drop_p = dropout_p; p.drop_mask = drop_mask;
    p.dqv = dQV;
    p.dkp_part = reinterpret_cast<float*>(workspace);
    p.k_planes = reinterpret_cast<__nv_bfloat16*>(p.dkp_part + (int64_t)pl.splits * B * Ksel * d);
    p.do_planes = p.k_planes + (int64_t)B * (pl.KP + 1) * d * 2;
    attn_bwd_planes_kernel<<<dim3((unsigned)(B * h), 2), 256, 0, stream>>>(p);
    if (dropout_p > 0.f) {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_bwd_tc_kernel<true>), (int)pl.smem));
        attn_bwd_tc_kernel<true><<<pl.grid, AB_THREADS, pl.smem, stream>>>(p);
    } else {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_bwd_tc_kernel<false>), (int)pl.smem));
        attn_bwd_tc_kernel<false><<<pl.grid, AB_THREADS, pl.smem, stream>>>(p);
    }
    launch_fold_partials(p.dkp_part, pl.splits, B * Ksel * d / 4, dKp, stream);
    return check_launch("snuffy_sparse_attn_bwd_tc", 3);
}

#ifdef ATTN_DEBUG_TIMING
int snuffy_attn_bwd_debug_read(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_attnb_dbg, sizeof(g_attnb_dbg)) == cudaSuccess ? 0 : 1;
}
#endif

}  // extern "C"
#pragma GCC visibility pop
