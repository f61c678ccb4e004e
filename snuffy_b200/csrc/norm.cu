// LayerNorm kernels of the aggregator (eps 1e-5, affine; snuffy.py:80,97,107,110,86):
//   ln_rows       : u = LN(y) for every row, where y is "x with the selected rows replaced"
//                   (read through row_map, never materialised; snuffy.py:152-155), written as fp32
//                   and/or as the split-bf16 operand planes the tcgen05 GEMM consumes (common.cuh).
//   ln_mean_head  : final LN -> mean over ALL N tokens -> linear head (snuffy.py:86,71), one pass
//                   over x, deterministic two-level reduction, no atomics on data.
#include "common.cuh"
#include "../../include/snuffy_b200.h"

namespace snuffy {

constexpr float LN_EPS = 1e-5f;

// One warp per row, 8 rows (= one 8-row core-matrix group of the plane layout) per CTA.
// Each lane keeps 8 consecutive elements per iteration: e = (it*32 + lane)*8.
template <int MAXIT>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_map, const float* __restrict__ alt,
               const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int d,
               float* __restrict__ out_f32, __nv_bfloat16* __restrict__ planes, int64_t plane_stride,
               float* __restrict__ stats, int apply_ln, int rc, const float* __restrict__ score_w = nullptr,
               const float* __restrict__ score_b = nullptr, float* __restrict__ score_out = nullptr, int score_c = 0) {
    extern __shared__ __align__(16) unsigned char ln_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * 8;
    const int64_t row = row0 + warp;
    const int nunits = (int)plane_kblocks(d) * 4;          // 16-byte units per row incl. K padding
    // staging [nunits][8 rows (+1 pad)] of 16-byte units: the pad rotates the 16-byte slot by the unit index, so a warp's
    // 32 stores (same row, consecutive units) spread over all banks instead of serialising on four
    bf16x8* s_hi = reinterpret_cast<bf16x8*>(ln_smem);
    bf16x8* s_lo = s_hi + (size_t)nunits * 9;

    float v[MAXIT][8];
    const bool live = row < rows;
    const float* src = nullptr;
    if (live) {
        src = x + row * (int64_t)d;
        if (row_map) {
            const int32_t slot = row_map[row];
            if (slot >= 0) src = alt + (int64_t)slot * d;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
        const int e = (it * 32 + lane) * 8;
        if (live && e < d) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + e));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + e + 4));
            v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
            v[it][4] = b.x; v[it][5] = b.y; v[it][6] = b.z; v[it][7] = b.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[it][j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[it][j] = 0.f;
        }
    }
    if (score_out) {
        // fused instance scorer c = y W^T + b (snuffy.py:39-41) on the RAW row while it is in registers: layer 0 then reads
        // the bag once for both the scores and the normalised planes.  Same per-lane partition and accumulation order as
        // scores_kernel (select.cu) -> bit-identical scores.
        for (int c = 0; c < score_c; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int e = (it * 32 + lane) * 8;
                if (live && e < d) {
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(score_w + (int64_t)c * d + e));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(score_w + (int64_t)c * d + e + 4));
                    acc = fmaf(v[it][0], w0.x, acc); acc = fmaf(v[it][1], w0.y, acc);
                    acc = fmaf(v[it][2], w0.z, acc); acc = fmaf(v[it][3], w0.w, acc);
                    acc = fmaf(v[it][4], w1.x, acc); acc = fmaf(v[it][5], w1.y, acc);
                    acc = fmaf(v[it][6], w1.z, acc); acc = fmaf(v[it][7], w1.w, acc);
                }
            }
            acc = warp_sum(acc);
            if (live && lane == 0) score_out[row * score_c + c] = acc + (score_b ? score_b[c] : 0.f);
        }
    }
    // apply_ln: 0 = plain convert (optionally scaled per column by gamma: folds a LayerNorm gain into a weight),
    //           1 = LayerNorm with affine, 2 = normalise only (z = (y - mean) * rstd; the affine lives in the weights)
    float mean = 0.f, rstd = 1.f;
    if (apply_ln) {
        mean = warp_sum(sum) / (float)d;
        float sq = 0.f;
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int e = (it * 32 + lane) * 8;
            if (e < d) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float t = v[it][j] - mean; sq = fmaf(t, t, sq); }
            }
        }
        rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
        if (stats && live && lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
        const int e = (it * 32 + lane) * 8;
        const int unit = it * 32 + lane;
        if (unit < nunits) {
            float u[8];
            if (live && e < d) {
                if (apply_ln == 1) {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + e));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + e + 4));
                    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[j] = (v[it][j] - mean) * rstd * g[j] + bb[j];
                } else if (apply_ln == 2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[j] = (v[it][j] - mean) * rstd;
                } else if (gamma) {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e + 4));
                    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[j] = v[it][j] * g[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[j] = v[it][j];
                }
                if (out_f32) {
                    float* o = out_f32 + row * (int64_t)d + e;
                    *reinterpret_cast<float4*>(o) = make_float4(u[0], u[1], u[2], u[3]);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(u[4], u[5], u[6], u[7]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = 0.f;
            }
            if (planes) {
                bf16x8 h, l;
#pragma unroll
                for (int j = 0; j < 8; ++j) split_bf16(u[j], h.v[j], l.v[j]);
                s_hi[unit * 9 + warp] = h;
                s_lo[unit * 9 + warp] = l;
            }
        }
    }
    if (planes) {
        __syncthreads();
        // each unit's 8 rows are 128 contiguous bytes in the plane: coalesced copy-out
        for (int s = threadIdx.x; s < nunits * 8; s += blockDim.x) {
            const int unit = s >> 3, rr = s & 7;
            const int64_t off = plane_unit_offset(row0, (int64_t)unit * 8, d, rc) + rr * 8;
            *reinterpret_cast<bf16x8*>(planes + off) = s_hi[unit * 9 + rr];
            *reinterpret_cast<bf16x8*>(planes + plane_stride + off) = s_lo[unit * 9 + rr];
        }
    }
}

// generic-d fallback (d % 8 != 0): fp32 output only
__global__ void __launch_bounds__(256)
ln_rows_scalar_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_map, const float* __restrict__ alt,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int d,
                      float* __restrict__ out_f32, float* __restrict__ stats, int apply_ln) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* src = x + row * (int64_t)d;
    if (row_map) { const int32_t slot = row_map[row]; if (slot >= 0) src = alt + (int64_t)slot * d; }
    float mean = 0.f, rstd = 1.f;
    if (apply_ln) {
        float sum = 0.f;
        for (int e = lane; e < d; e += 32) sum += src[e];
        mean = warp_sum(sum) / (float)d;
        float sq = 0.f;
        for (int e = lane; e < d; e += 32) { const float t = src[e] - mean; sq = fmaf(t, t, sq); }
        rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
        if (stats && lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    }
    for (int e = lane; e < d; e += 32)
        out_f32[row * (int64_t)d + e] = apply_ln == 1 ? (src[e] - mean) * rstd * gamma[e] + beta[e]
                                      : apply_ln == 2 ? (src[e] - mean) * rstd : (gamma ? src[e] * gamma[e] : src[e]);
}

// Re-normalise K scattered rows in place inside an existing plane set: row k of `src` [B*K, d] (the updated selected
// rows X_S') goes to plane row b*N + idx[b, k].  Lets LN1 and LN2 share one set of z planes: only the Ksel rows the
// attention sub-layer changed are rewritten (snuffy.py:152-155 + 110) instead of a second pass over all N rows.
__global__ void __launch_bounds__(256)
ln_rows_scatter_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t N, int64_t K, int64_t total,
                       int d, const float* __restrict__ gamma, const float* __restrict__ beta, int apply_ln,
                       __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t slot = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (slot >= total) return;
    const int64_t r = idx[slot];
    if (r < 0 || r >= N) return;
    const int64_t dst = (slot / K) * N + r;
    const float* s = src + slot * (int64_t)d;
    float mean = 0.f, rstd = 1.f;
    if (apply_ln) {
        float sum = 0.f;
        for (int e = lane * 4; e < d; e += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s + e));
            sum += (a.x + a.y) + (a.z + a.w);
        }
        mean = warp_sum(sum) / (float)d;
        float sq = 0.f;
        for (int e = lane * 4; e < d; e += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s + e));
            const float t0 = a.x - mean, t1 = a.y - mean, t2 = a.z - mean, t3 = a.w - mean;
            sq += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
        }
        rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
    }
    const int nunits = (int)plane_kblocks(d) * 4;
    for (int unit = lane; unit < nunits; unit += 32) {
        const int e = unit * 8;
        bf16x8 h, l;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float u = 0.f;
            if (e + j < d) {
                u = s[e + j];
                if (apply_ln) u = (u - mean) * rstd;
                if (apply_ln == 1) u = u * gamma[e + j] + beta[e + j];
            }
            split_bf16(u, h.v[j], l.v[j]);
        }
        const int64_t off = plane_unit_offset(dst, e, d, 128);
        *reinterpret_cast<bf16x8*>(planes + off) = h;
        *reinterpret_cast<bf16x8*>(planes + plane_stride + off) = l;
    }
}

// ------------------------------------------------------------------ final LN + mean + head
// grid (chunks, B).  Each CTA normalises its rows (warp per row), sums the normalised rows,
// writes one partial [d] and takes a ticket; the last CTA of the bag folds the partials in a
// fixed order, applies the affine (mean(z*g+b) = g*mean(z)+b), divides by N and runs the C x d head.
// tail shared by both variants: write this CTA's partial, take a ticket, the last CTA of the bag folds + runs the head
__device__ __forceinline__ void ln_mean_head_tail(float* hs, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  const float* __restrict__ Wh, const float* __restrict__ bh, int64_t N, int d,
                                                  int C, float* __restrict__ partials, unsigned int* __restrict__ tickets,
                                                  float* __restrict__ pooled, float* __restrict__ bag_out) {
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = gridDim.x, chunk = blockIdx.x, bag = blockIdx.y;
    float* part = partials + ((int64_t)bag * chunks + chunk) * d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += hs[(size_t)w * d + e];
        part[e] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&tickets[bag], 1u);
        s_last = (t == (unsigned)chunks - 1u);
        if (s_last) tickets[bag] = 0;                      // leave the workspace reusable
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float inv_n = 1.f / (float)N;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
        // fixed summation order (deterministic); 8 independent loads in flight: this serial tail is what a single bag waits on
        const float* col = partials + (int64_t)bag * chunks * d + e;
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int ch = 0;
        for (; ch + 8 <= chunks; ch += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += __ldcg(col + (int64_t)(ch + u) * d);
        }
        for (; ch < chunks; ++ch) a[0] += __ldcg(col + (int64_t)ch * d);
        const float s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        const float p = gamma[e] * (s * inv_n) + beta[e];
        hs[e] = p;
        if (pooled) pooled[(int64_t)bag * d + e] = p;
    }
    __syncthreads();
    for (int j = warp; j < C; j += 8) {
        float acc = 0.f;
        for (int e = lane; e < d; e += 32) acc = fmaf(hs[e], Wh[(int64_t)j * d + e], acc);
        acc = warp_sum(acc);
        if (lane == 0) bag_out[(int64_t)bag * C + j] = acc + (bh ? bh[j] : 0.f);
    }
}

// d <= MAXIT * 128: the row and the column accumulators live in registers (one 128-bit load per element, two rows in
// flight per warp); the generic kernel below re-reads the row from L1 and accumulates in shared memory.
template <int MAXIT>
__global__ void __launch_bounds__(256)
ln_mean_head_reg_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ Wh, const float* __restrict__ bh, int64_t N, int d, int C,
                        float* __restrict__ partials, unsigned int* __restrict__ tickets, float* __restrict__ stats,
                        float* __restrict__ pooled, float* __restrict__ bag_out, const int64_t* __restrict__ cu_seqlens) {
    extern __shared__ __align__(16) float hs[];            // [8 warps][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = gridDim.x, chunk = blockIdx.x, bag = blockIdx.y;
    int64_t row_base = (int64_t)bag * N;                    // packed variable-length bags: rows [cu[bag], cu[bag+1])
    if (cu_seqlens) { row_base = cu_seqlens[bag]; N = cu_seqlens[bag + 1] - row_base; }
    const int64_t per = (N + chunks - 1) / chunks;
    const int64_t r0 = chunk * per, r1 = min(N, r0 + per);
    const float* xb = x + row_base * d;
    if (stats) stats += row_base * 2 - (int64_t)bag * N * 2;
    const float inv_d = 1.f / (float)d;
    float4 acc[MAXIT];
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + warp; r < r1; r += 16) {
        const bool two = r + 8 < r1;
        float4 a[MAXIT], b[MAXIT];
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int e = (it * 32 + lane) * 4;
            a[it] = e < d ? ld_stream(reinterpret_cast<const float4*>(xb + r * d + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[it] = (two && e < d) ? ld_stream(reinterpret_cast<const float4*>(xb + (r + 8) * d + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            sa += (a[it].x + a[it].y) + (a[it].z + a[it].w);
            sb += (b[it].x + b[it].y) + (b[it].z + b[it].w);
        }
        const float ma = warp_sum(sa) * inv_d, mb = warp_sum(sb) * inv_d;
        float qa = 0.f, qb = 0.f;
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            if ((it * 32 + lane) * 4 < d) {
                a[it].x -= ma; a[it].y -= ma; a[it].z -= ma; a[it].w -= ma;
                b[it].x -= mb; b[it].y -= mb; b[it].z -= mb; b[it].w -= mb;
                qa += (a[it].x * a[it].x + a[it].y * a[it].y) + (a[it].z * a[it].z + a[it].w * a[it].w);
                qb += (b[it].x * b[it].x + b[it].y * b[it].y) + (b[it].z * b[it].z + b[it].w * b[it].w);
            }
        }
        const float ra = rsqrtf(warp_sum(qa) * inv_d + LN_EPS);
        const float rb = two ? rsqrtf(warp_sum(qb) * inv_d + LN_EPS) : 0.f;
        if (stats && lane == 0) {
            stats[((int64_t)bag * N + r) * 2] = ma; stats[((int64_t)bag * N + r) * 2 + 1] = ra;
            if (two) { stats[((int64_t)bag * N + r + 8) * 2] = mb; stats[((int64_t)bag * N + r + 8) * 2 + 1] = rb; }
        }
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            acc[it].x += a[it].x * ra + b[it].x * rb; acc[it].y += a[it].y * ra + b[it].y * rb;
            acc[it].z += a[it].z * ra + b[it].z * rb; acc[it].w += a[it].w * ra + b[it].w * rb;
        }
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
        const int e = (it * 32 + lane) * 4;
        if (e < d) *reinterpret_cast<float4*>(hs + (size_t)warp * d + e) = acc[it];
    }
    __syncthreads();
    ln_mean_head_tail(hs, gamma, beta, Wh, bh, N, d, C, partials, tickets, pooled, bag_out);
}

__global__ void __launch_bounds__(256)
ln_mean_head_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ Wh, const float* __restrict__ bh, int64_t N, int d, int C,
                    float* __restrict__ partials, unsigned int* __restrict__ tickets, float* __restrict__ stats,
                    float* __restrict__ pooled, float* __restrict__ bag_out, const int64_t* __restrict__ cu_seqlens) {
    extern __shared__ __align__(16) float hs[];            // [8 warps][d] then reused
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = gridDim.x, chunk = blockIdx.x, bag = blockIdx.y;
    int64_t row_base = (int64_t)bag * N;
    if (cu_seqlens) { row_base = cu_seqlens[bag]; N = cu_seqlens[bag + 1] - row_base; }
    const int64_t per = (N + chunks - 1) / chunks;
    const int64_t r0 = chunk * per, r1 = min(N, r0 + per);
    const float* xb = x + row_base * d;
    if (stats) stats += row_base * 2 - (int64_t)bag * N * 2;
    float* wacc = hs + (size_t)warp * d;
    for (int e = lane; e < d; e += 32) wacc[e] = 0.f;
    for (int64_t r = r0 + warp; r < r1; r += 8) {
        const float* src = xb + r * d;
        float sum = 0.f;
        for (int e = lane * 4; e < d; e += 128) {
            const float4 a = ld_stream(reinterpret_cast<const float4*>(src + e));
            sum += (a.x + a.y) + (a.z + a.w);
        }
        const float mean = warp_sum(sum) / (float)d;
        float sq = 0.f;
        for (int e = lane * 4; e < d; e += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + e));   // L1/L2 hit
            const float t0 = a.x - mean, t1 = a.y - mean, t2 = a.z - mean, t3 = a.w - mean;
            sq += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
        }
        const float rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
        if (stats && lane == 0) { stats[((int64_t)bag * N + r) * 2] = mean; stats[((int64_t)bag * N + r) * 2 + 1] = rstd; }
        for (int e = lane * 4; e < d; e += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + e));
            float4 w = *reinterpret_cast<float4*>(wacc + e);
            w.x += (a.x - mean) * rstd; w.y += (a.y - mean) * rstd;
            w.z += (a.z - mean) * rstd; w.w += (a.w - mean) * rstd;
            *reinterpret_cast<float4*>(wacc + e) = w;
        }
    }
    __syncthreads();
    ln_mean_head_tail(hs, gamma, beta, Wh, bh, N, d, C, partials, tickets, pooled, bag_out);
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// LN over rows of y = (row_map ? x with mapped rows taken from alt : x); apply_ln = 0 turns it into a
// plain gather/convert (used to split raw activations).  Outputs (each optional): fp32 [rows, d];
// split-bf16 planes (hi at `planes`, lo at `planes + plane_stride`, tiled with `plane_rc` rows per chunk:
// 128 for an A operand, snuffy_gemm_tc_block_n(rows) for a weight); stats [rows, 2] = (mean, rstd).
int snuffy_ln_rows_fwd(const float* x, const int32_t* row_map, const float* alt, const float* gamma,
                       const float* beta, int64_t rows, int64_t d, int apply_ln, float* out_f32, void* planes,
                       int64_t plane_stride, int plane_rc, float* stats, cudaStream_t stream) {
    SNUFFY_REQUIRE(x && rows >= 0 && d > 0, "snuffy_ln_rows_fwd: bad arguments");
    SNUFFY_REQUIRE(apply_ln >= 0 && apply_ln <= 2, "snuffy_ln_rows_fwd: apply_ln must be 0, 1 or 2");
    SNUFFY_REQUIRE(apply_ln != 1 || (gamma && beta), "snuffy_ln_rows_fwd: LayerNorm needs gamma and beta");
    SNUFFY_REQUIRE(!row_map || alt, "snuffy_ln_rows_fwd: row_map given without the replacement rows");
    SNUFFY_REQUIRE(!planes || plane_rc == 128 || plane_rc == 256, "snuffy_ln_rows_fwd: plane_rc must be 128 or 256");
    if (rows == 0) return 0;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    const bool vec = d % 8 == 0 && ((uintptr_t)x % 16 == 0) && (!alt || (uintptr_t)alt % 16 == 0) &&
                     (!out_f32 || (uintptr_t)out_f32 % 16 == 0);
    if (!vec) {
        SNUFFY_REQUIRE(!planes && out_f32, "snuffy_ln_rows_fwd: d=%lld not a multiple of 8 supports fp32 output only",
                       (long long)d);
        ln_rows_scalar_kernel<<<grid, 256, 0, stream>>>(x, row_map, alt, gamma, beta, rows, (int)d, out_f32, stats,
                                                        apply_ln);
        return check_launch("snuffy_ln_rows_fwd");
    }
    SNUFFY_REQUIRE(d <= 4096, "snuffy_ln_rows_fwd: d=%lld > 4096 unsupported", (long long)d);
    const size_t smem = planes ? (size_t)plane_kblocks(d) * 4 * 9 * 16 * 2 : 0;
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
#define LN_LAUNCH(MAXIT)                                                                                        \
    do {                                                                                                        \
        if (smem > 48 * 1024)                                                                                   \
            SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&ln_rows_kernel<MAXIT>), \
                                             (int)smem));                                                      \
        ln_rows_kernel<MAXIT><<<grid, 256, smem, stream>>>(x, row_map, alt, gamma, beta, rows, (int)d, out_f32,  \
                                                           pl, plane_stride, stats, apply_ln, plane_rc);        \
    } while (0)
    const int iters = (int)((plane_kblocks(d) * 4 + 31) / 32);
    if (iters <= 2) LN_LAUNCH(2);
    else if (iters <= 4) LN_LAUNCH(4);
    else if (iters <= 8) LN_LAUNCH(8);
    else LN_LAUNCH(16);
#undef LN_LAUNCH
    return check_launch("snuffy_ln_rows_fwd");
}

// Layer-0 fusion (inference): ONE pass over the bag produces the instance scores c = x W^T + b (snuffy.py:39-41) and the
// normalised operand planes z = (x - mean) * rstd that LN1 / LN2 share (snuffy.py:107,110).  c [rows, C]; needs d % 8 == 0.
int snuffy_scores_ln_planes_fwd(const float* x, const float* W, const float* bias, int64_t rows, int64_t d, int64_t C,
                                float* c, void* planes, int64_t plane_stride, float* stats, cudaStream_t stream) {
    SNUFFY_REQUIRE(x && W && c && planes && rows >= 1 && C >= 1, "snuffy_scores_ln_planes_fwd: bad arguments");
    SNUFFY_REQUIRE(d % 8 == 0 && d <= 4096 && (uintptr_t)x % 16 == 0 && (uintptr_t)W % 16 == 0,
                   "snuffy_scores_ln_planes_fwd: needs d %% 8 == 0, d <= 4096 and 16-byte aligned rows (d=%lld)", (long long)d);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    const size_t smem = (size_t)plane_kblocks(d) * 4 * 9 * 16 * 2;
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
#define LNS_LAUNCH(MAXIT)                                                                                              \
    do {                                                                                                               \
        if (smem > 48 * 1024)                                                                                          \
            SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&ln_rows_kernel<MAXIT>), (int)smem)); \
        ln_rows_kernel<MAXIT><<<grid, 256, smem, stream>>>(x, nullptr, nullptr, nullptr, nullptr, rows, (int)d, nullptr, pl,  \
                                                           plane_stride, stats, 2, 128, W, bias, c, (int)C);            \
    } while (0)
    const int iters = (int)((plane_kblocks(d) * 4 + 31) / 32);
    if (iters <= 2) LNS_LAUNCH(2);
    else if (iters <= 4) LNS_LAUNCH(4);
    else if (iters <= 8) LNS_LAUNCH(8);
    else LNS_LAUNCH(16);
#undef LNS_LAUNCH
    return check_launch("snuffy_scores_ln_planes_fwd");
}

// planes rows b*N + idx[b,k] <- split(LN(src[b*K + k]))   (apply_ln as in snuffy_ln_rows_fwd; A-operand planes, RC = 128)
int snuffy_ln_rows_scatter_planes(const float* src, const int64_t* idx, int64_t B, int64_t N, int64_t K, int64_t d,
                                  const float* gamma, const float* beta, int apply_ln, void* planes, int64_t plane_stride,
                                  cudaStream_t stream) {
    SNUFFY_REQUIRE(src && idx && planes && B >= 1 && N >= 1 && K >= 0, "snuffy_ln_rows_scatter_planes: bad arguments");
    SNUFFY_REQUIRE(d % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)planes % 16 == 0 && plane_stride % 8 == 0,
                   "snuffy_ln_rows_scatter_planes: needs d %% 8 == 0 and 16-byte aligned buffers");
    SNUFFY_REQUIRE(apply_ln >= 0 && apply_ln <= 2 && (apply_ln != 1 || (gamma && beta)),
                   "snuffy_ln_rows_scatter_planes: bad LayerNorm mode");
    if (K == 0) return 0;
    const int64_t total = B * K;
    ln_rows_scatter_kernel<<<(unsigned)((total + 7) / 8), 256, 0, stream>>>(
        src, idx, N, K, total, (int)d, gamma, beta, apply_ln, reinterpret_cast<__nv_bfloat16*>(planes), plane_stride);
    return check_launch("snuffy_ln_rows_scatter_planes");
}

// workspace floats needed by snuffy_ln_mean_head_fwd (partials) -- tickets are B uint32 (zeroed once by the caller)
int64_t snuffy_ln_mean_head_chunks(int64_t B, int64_t N) {
    int64_t chunks = (2 * (int64_t)sm_count() + B - 1) / (B > 0 ? B : 1);
    const int64_t max_chunks = (N + 7) / 8;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    return chunks;
}

// bag[B, C] = head( mean_n LN_f(x[b, n, :]) )        snuffy.py:86 + 71
static int launch_ln_mean_head(const float* x, const float* gamma, const float* beta, const float* Wh, const float* bh,
                               int64_t B, int64_t N, int64_t d, int64_t C, float* partials, uint32_t* tickets,
                               float* stats, float* pooled, float* bag_out, const int64_t* cu_seqlens, cudaStream_t stream) {
    SNUFFY_REQUIRE(x && gamma && beta && Wh && partials && tickets && bag_out, "snuffy_ln_mean_head_fwd: null pointer");
    SNUFFY_REQUIRE(B >= 1 && N >= 1 && C >= 1 && d % 4 == 0 && (uintptr_t)x % 16 == 0,
                   "snuffy_ln_mean_head_fwd: needs d %% 4 == 0 and 16-byte aligned rows (d=%lld)", (long long)d);
    const int64_t chunks = snuffy_ln_mean_head_chunks(B, N);
    const size_t smem = (size_t)8 * d * sizeof(float);
    SNUFFY_REQUIRE(smem <= 200 * 1024, "snuffy_ln_mean_head_fwd: d=%lld too large", (long long)d);
    if (smem > 48 * 1024)
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&ln_mean_head_kernel), (int)smem));
    dim3 grid((unsigned)chunks, (unsigned)B);
    if (d <= 512) {
        if (smem > 48 * 1024)
            SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&ln_mean_head_reg_kernel<4>), (int)smem));
        ln_mean_head_reg_kernel<4><<<grid, 256, smem, stream>>>(x, gamma, beta, Wh, bh, N, (int)d, (int)C, partials, tickets,
                                                                stats, pooled, bag_out, cu_seqlens);
    } else if (d <= 1024) {
        if (smem > 48 * 1024)
            SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&ln_mean_head_reg_kernel<8>), (int)smem));
        ln_mean_head_reg_kernel<8><<<grid, 256, smem, stream>>>(x, gamma, beta, Wh, bh, N, (int)d, (int)C, partials, tickets,
                                                                stats, pooled, bag_out, cu_seqlens);
    } else {
        ln_mean_head_kernel<<<grid, 256, smem, stream>>>(x, gamma, beta, Wh, bh, N, (int)d, (int)C, partials, tickets,
                                                         stats, pooled, bag_out, cu_seqlens);
    }
    return check_launch("snuffy_ln_mean_head_fwd");
}

int snuffy_ln_mean_head_fwd(const float* x, const float* gamma, const float* beta, const float* Wh, const float* bh,
                            int64_t B, int64_t N, int64_t d, int64_t C, float* partials, uint32_t* tickets,
                            float* stats, float* pooled, float* bag_out, cudaStream_t stream) {
    return launch_ln_mean_head(x, gamma, beta, Wh, bh, B, N, d, C, partials, tickets, stats, pooled, bag_out, nullptr, stream);
}

// packed variable-length bags: x [T, d], bag b = rows [cu_seqlens[b], cu_seqlens[b+1]); max_n sizes the grid / partials
// (snuffy_ln_mean_head_chunks(B, max_n)).
int snuffy_ln_mean_head_varlen_fwd(const float* x, const int64_t* cu_seqlens, const float* gamma, const float* beta,
                                   const float* Wh, const float* bh, int64_t B, int64_t max_n, int64_t d, int64_t C,
                                   float* partials, uint32_t* tickets, float* pooled, float* bag_out, cudaStream_t stream) {
    SNUFFY_REQUIRE(cu_seqlens, "snuffy_ln_mean_head_varlen_fwd: null cu_seqlens");
    return launch_ln_mean_head(x, gamma, beta, Wh, bh, B, max_n, d, C, partials, tickets, nullptr, pooled, bag_out, cu_seqlens,
                               stream);
}

}  // extern "C"
#pragma GCC visibility pop

// ------------------------------------------------------------------ transposed operand planes (weight gradients)
// planes of X^T for an fp32 X [R, C]: plane row = column c of X, k = row r of X (the contraction of dW = dY^T X runs over
// the rows).  Optional element-wise prologue so the transposed operand is produced straight from what the forward saved:
//   mode 0: x                       mode 1: LayerNorm(x-with-mapped-rows) from saved (mean, rstd) + affine
//   mode 2: dropout(act(x))  (x = saved pre-activation)
namespace snuffy {
struct PlanesTParams {
    const float* x; int64_t ldx; int64_t R; int C; int rc; int mode;
    const float* stats; const float* gamma; const float* beta; const int32_t* row_map; const float* alt;
    int act; float drop_p; uint64_t seed, offset;
    __nv_bfloat16* planes; int64_t plane_stride;
};

__global__ void __launch_bounds__(256)
planes_t_kernel(const PlanesTParams p) {
    __shared__ __align__(16) float tile[32][132];        // 16-byte aligned rows: float4 stores, conflict-free column reads
    const int t = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = t + i * 256, row = idx >> 5, c4 = (idx & 31) * 4;
        const int64_t r = r0 + row;
        const int c = c0 + c4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < p.R && c < p.C) {
            const float* src = p.x + r * p.ldx;
            if (p.mode == 1 && p.row_map) { const int32_t slot = p.row_map[r]; if (slot >= 0) src = p.alt + (int64_t)slot * p.ldx; }
            if (c + 3 < p.C) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src + c));
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            } else {
                for (int j = 0; j < 4; ++j) if (c + j < p.C) v[j] = src[c + j];
            }
            if (p.mode == 1) {
                const float mean = p.stats[r * 2], rstd = p.stats[r * 2 + 1];
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < p.C) v[j] = (v[j] - mean) * rstd * p.gamma[c + j] + p.beta[c + j];
            } else if (p.mode == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = act_apply(p.act, v[j]);
                    if (p.drop_p > 0.f) {
                        const DrawKey key = rng_resolve(p.seed, p.offset);
                        a *= drop_keep_scale(key.seed, key.offset, (uint64_t)(r * p.C + c + j), p.drop_p);
                    }
                    v[j] = (c + j < p.C) ? a : 0.f;
                }
            }
        }
        *reinterpret_cast<float4*>(&tile[row][c4]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    const int64_t nkb = plane_kblocks(p.R);
    const int64_t kb = blockIdx.y;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int u = t + i * 256, kg = u >> 7, cl = u & 127;
        const int64_t prow = c0 + cl;                              // plane row = column of X (rows >= C are zero padding)
        bf16x8 h, l;
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16(tile[kg * 8 + j][cl], h.v[j], l.v[j]);
        const int64_t rt = prow / p.rc, rr = prow % p.rc;
        const int64_t off = ((rt * nkb + kb) * 4 + kg) * (int64_t)p.rc * 8 + rr * 8;
        *reinterpret_cast<bf16x8*>(p.planes + off) = h;
        *reinterpret_cast<bf16x8*>(p.planes + p.plane_stride + off) = l;
    }
}

// rows [row0, row1) of a plane set := 0 (both planes): the tail of the last 16-row group when a product contracts over rows
__global__ void __launch_bounds__(256)
planes_zero_rows_kernel(__nv_bfloat16* planes, int64_t plane_stride, int64_t K, int rc, int64_t row0, int64_t row1) {
    const int64_t units = (row1 - row0) * plane_kblocks(K) * 4;
    const int64_t u = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (u >= units) return;
    const int64_t row = row0 + u % (row1 - row0), kunit = u / (row1 - row0);
    const int64_t off = plane_unit_offset(row, kunit * 8, K, rc);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(planes + off) = z;
    *reinterpret_cast<uint4*>(planes + plane_stride + off) = z;
}

// ---------------------------------------------------------------------------------------------------------------------
// All derived weight operands of a layer in ONE launch.  A training step re-derives them from the updated parameters every
// step: per layer the planes of Wq|Wv, W1, W2, Wk, Wo and of their transposes, the fused Q|V bias: ten conversions, two
// concatenations and the zero fills of the padded plane sets were ~19 launches of 2-8 us.  Each CTA converts one
// 128-plane-row x 32-k block of one job and writes the padding itself.
struct PlaneJobs {
    snuffy_plane_job_t job[SNUFFY_MAX_PLANE_JOBS];
    int32_t cta0[SNUFFY_MAX_PLANE_JOBS + 1];               // first CTA of each job
    int32_t n;
};

__global__ void __launch_bounds__(256)
weight_planes_batch_kernel(const __grid_constant__ PlaneJobs P) {
    __shared__ __align__(16) float tile[32][132];
    const int t = threadIdx.x, bid = blockIdx.x;
    int j = 0;
    while (j + 1 < P.n && bid >= P.cta0[j + 1]) ++j;
    const snuffy_plane_job_t& jb = P.job[j];
    const int local = bid - P.cta0[j];
    if (jb.kind == 2) {                                     // fp32 copy (fused bias): 1024 floats per CTA
        const int64_t total = jb.rows * jb.cols;
        float* dst = reinterpret_cast<float*>(jb.dst) + jb.dst_row0;
        for (int64_t i = (int64_t)local * 1024 + t; i < min(total, (int64_t)(local + 1) * 1024); i += 256) dst[i] = jb.src[i];
        return;
    }
    const int rc = jb.plane_rc;
    const int64_t prows = jb.kind == 0 ? jb.rows : jb.cols;      // plane rows this job fills
    const int64_t kdim = jb.kind == 0 ? jb.cols : jb.rows;       // contraction length of the plane set
    const int64_t nkb = plane_kblocks(kdim), nkb_set = plane_kblocks(jb.k_total), kb_set0 = jb.dst_k0 / PLANE_KB;
    const int64_t kb = local % nkb, rb = local / nkb;
    __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(jb.dst);
    if (jb.kind == 1) {                                     // transposed: k runs over the rows of src
        const int64_t r0 = kb * 32, c0 = rb * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = t + i * 256, row = idx >> 5, c4 = (idx & 31) * 4;
            const int64_t r = r0 + row, c = c0 + c4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (r < jb.rows && c < jb.cols) {
                const float* src = jb.src + r * jb.ld;
                if (c + 3 < jb.cols) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(src + c));
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                } else {
                    for (int q = 0; q < 4; ++q) if (c + q < jb.cols) v[q] = src[c + q];
                }
            }
            *reinterpret_cast<float4*>(&tile[row][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int u = t + i * 256, kg = u >> 7, rl = u & 127;
        const int64_t prl = rb * 128 + rl;                  // plane row inside the job
        bf16x8 h, l;
        if (jb.kind == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) split_bf16(tile[kg * 8 + q][rl], h.v[q], l.v[q]);
        } else {
            const int64_t k0 = kb * 32 + kg * 8;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (prl < prows && k0 < kdim) {
                const float* src = jb.src + prl * jb.ld + k0;
                if (k0 + 7 < kdim) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
                    for (int q = 0; q < 8; ++q) if (k0 + q < kdim) v[q] = src[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) split_bf16(v[q], h.v[q], l.v[q]);
        }
        const int64_t prow = jb.dst_row0 + prl;
        const int64_t rt = prow / rc, rr = prow % rc;
        const int64_t off = ((rt * nkb_set + kb_set0 + kb) * 4 + kg) * (int64_t)rc * 8 + rr * 8;
        *reinterpret_cast<bf16x8*>(planes + off) = h;
        *reinterpret_cast<bf16x8*>(planes + jb.plane_stride + off) = l;
    }
}
}  // namespace snuffy

#pragma GCC visibility push(default)
extern "C" int snuffy_planes_t_fwd(const float* x, int64_t ldx, int64_t R, int64_t C, int plane_rc, int mode,
                                   const float* stats, const float* gamma, const float* beta, const int32_t* row_map,
                                   const float* alt, int act, float dropout_p, uint64_t seed, uint64_t offset, void* planes,
                                   int64_t plane_stride, cudaStream_t stream) {
    using namespace snuffy;
    SNUFFY_REQUIRE(x && planes && R >= 1 && C >= 1, "snuffy_planes_t_fwd: bad arguments");
    SNUFFY_REQUIRE(plane_rc == 128 || plane_rc == 256, "snuffy_planes_t_fwd: plane_rc must be 128 or 256");
    SNUFFY_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || (stats && gamma && beta)) && (!row_map || alt),
                   "snuffy_planes_t_fwd: inconsistent prologue arguments");
    SNUFFY_REQUIRE(ldx % 4 == 0 && (uintptr_t)x % 16 == 0 && (!alt || (uintptr_t)alt % 16 == 0) && (uintptr_t)planes % 16 == 0 &&
                       plane_stride % 8 == 0, "snuffy_planes_t_fwd: rows must be 16-byte aligned");
    PlanesTParams p{x, ldx, R, (int)C, plane_rc, mode, stats, gamma, beta, row_map, alt, act, dropout_p, seed, offset,
                    reinterpret_cast<__nv_bfloat16*>(planes), plane_stride};
    const int64_t cpad = plane_rtiles(C, plane_rc) * plane_rc;
    dim3 grid((unsigned)(cpad / 128), (unsigned)plane_kblocks(R));
    SNUFFY_REQUIRE(grid.y <= 65535, "snuffy_planes_t_fwd: too many rows for one launch");
    planes_t_kernel<<<grid, 256, 0, stream>>>(p);
    return check_launch("snuffy_planes_t_fwd");
}
extern "C" int snuffy_plane_job_bytes(void) { return (int)sizeof(snuffy_plane_job_t); }
extern "C" int snuffy_weight_planes_batch(const snuffy_plane_job_t* jobs, int64_t n_jobs, cudaStream_t stream) {
    using namespace snuffy;
    SNUFFY_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= SNUFFY_MAX_PLANE_JOBS, "snuffy_weight_planes_batch: 1..16 jobs per launch");
    PlaneJobs P;
    int64_t ctas = 0;
    for (int64_t j = 0; j < n_jobs; ++j) {
        const snuffy_plane_job_t& jb = jobs[j];
        SNUFFY_REQUIRE(jb.src && jb.dst && jb.rows >= 1 && jb.cols >= 1 && jb.kind >= 0 && jb.kind <= 2,
                       "snuffy_weight_planes_batch: bad job");
        P.job[j] = jb;
        P.cta0[j] = (int32_t)ctas;
        if (jb.kind == 2) {
            SNUFFY_REQUIRE(jb.ld == jb.cols, "snuffy_weight_planes_batch: copy jobs take contiguous rows");
            ctas += (jb.rows * jb.cols + 1023) / 1024;
            continue;
        }
        SNUFFY_REQUIRE(jb.plane_rc == 128 || jb.plane_rc == 256, "snuffy_weight_planes_batch: plane_rc must be 128 or 256");
        SNUFFY_REQUIRE(jb.dst_row0 % jb.plane_rc == 0, "snuffy_weight_planes_batch: jobs start on a row tile");
        SNUFFY_REQUIRE(jb.dst_k0 % PLANE_KB == 0 && jb.dst_k0 + (jb.kind == 0 ? jb.cols : jb.rows) <= jb.k_total,
                       "snuffy_weight_planes_batch: jobs start on a k block inside the plane set");
        SNUFFY_REQUIRE(jb.ld % 4 == 0 && (uintptr_t)jb.src % 16 == 0 && (uintptr_t)jb.dst % 16 == 0 && jb.plane_stride % 8 == 0,
                       "snuffy_weight_planes_batch: rows must be 16-byte aligned");
        const int64_t prows = jb.kind == 0 ? jb.rows : jb.cols, kdim = jb.kind == 0 ? jb.cols : jb.rows;
        ctas += plane_rtiles(prows, jb.plane_rc) * (jb.plane_rc / 128) * plane_kblocks(kdim);
    }
    P.cta0[n_jobs] = (int32_t)ctas;
    P.n = (int32_t)n_jobs;
    SNUFFY_REQUIRE(ctas < (1ll << 30), "snuffy_weight_planes_batch: too much work for one launch");
    weight_planes_batch_kernel<<<(unsigned)ctas, 256, 0, stream>>>(P);
    return check_launch("snuffy_weight_planes_batch");
}
extern "C" int snuffy_planes_zero_rows(void* planes, int64_t plane_stride, int64_t K, int plane_rc, int64_t row0, int64_t row1,
                                       cudaStream_t stream) {
    using namespace snuffy;
    SNUFFY_REQUIRE(planes && K >= 1 && (plane_rc == 128 || plane_rc == 256) && row0 >= 0 && row1 >= row0,
                   "snuffy_planes_zero_rows: bad arguments");
    if (row1 == row0) return 0;
    const int64_t units = (row1 - row0) * plane_kblocks(K) * 4;
    planes_zero_rows_kernel<<<(unsigned)((units + 255) / 256), 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(planes),
                                                                               plane_stride, K, plane_rc, row0, row1);
    return check_launch("snuffy_planes_zero_rows");
}
#pragma GCC visibility pop
