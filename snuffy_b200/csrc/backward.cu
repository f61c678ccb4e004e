// Backward kernels of the aggregator that are not plain contractions (train.py:259 `loss.backward()` runs through
// the drop-in modules; the contractions themselves go through snuffy_gemm_f32 / snuffy_gemm_f32_batched):
//   ln_rows_bwd        LayerNorm backward (snuffy.py:107,110,86) incl. dgamma / dbeta, optional residual-gradient add,
//                      optional broadcast upstream gradient (the mean-pool of snuffy.py:71 sends the same row to all N)
//   act_bwd            dh = da * dropout_mask * act'(h_pre), a = dropout(act(h_pre))   (snuffy.py:216-225)
//   colsum             out[c, :] = sum_rows w[row, c] * X[row, :]   (bias grads, instance-classifier grad)
//   softmax recompute / dS   the row-local pieces of the attention backward (snuffy.py:160-168)
//   scatter_add_rows   dx[S] += dxs  (backward of the raw-row gather snuffy.py:131,145-147)
#include "common.cuh"

namespace snuffy {

void launch_fold_partials(const float* part, int splits, int64_t n4, float* out, cudaStream_t stream);


// ------------------------------------------------------------------ LayerNorm backward
// warp per row; per-warp dgamma/dbeta accumulators in shared memory (fixed summation order -> deterministic);
// each CTA writes one partial [2][d] that fold_partials sums.
__global__ void __launch_bounds__(256)
ln_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy_bcast, int64_t rows_per_bag, float bscale,
                   const float* __restrict__ x, const int32_t* __restrict__ row_map, const float* __restrict__ alt,
                   const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ add,
                   int64_t rows, int d, float* __restrict__ dx, float* __restrict__ partials) {
    extern __shared__ __align__(16) float lb_smem[];        // [warps][2][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* accg = lb_smem + (size_t)warp * 2 * d;
    float* accb = accg + d;
    for (int e = lane; e < d; e += 32) { accg[e] = 0.f; accb[e] = 0.f; }
    __syncwarp();
    const float inv_d = 1.f / (float)d;
    for (int64_t row = (int64_t)blockIdx.x * nwarps + warp; row < rows; row += (int64_t)gridDim.x * nwarps) {
        const float* src = x + row * (int64_t)d;
        if (row_map) { const int32_t slot = row_map[row]; if (slot >= 0) src = alt + (int64_t)slot * d; }
        const float* g = dy ? dy + row * (int64_t)d : dy_bcast + (row / rows_per_bag) * (int64_t)d;
        const float gs = dy ? 1.f : bscale;
        const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
        float s1 = 0.f, s2 = 0.f;
        for (int e = lane * 4; e < d; e += 128) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(src + e));
            const float4 gv = __ldg(reinterpret_cast<const float4*>(g + e));
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + e));
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gg[4] = {gv.x * gs, gv.y * gs, gv.z * gs, gv.w * gs};
            const float gm[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = (xs[j] - mean) * rstd, t = gg[j] * gm[j];
                s1 += t; s2 = fmaf(t, z, s2);
                accg[e + j] = fmaf(gg[j], z, accg[e + j]);
                accb[e + j] += gg[j];
            }
        }
        s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
        if (dx) {
            for (int e = lane * 4; e < d; e += 128) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(src + e));
                const float4 gv = __ldg(reinterpret_cast<const float4*>(g + e));
                const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + e));
                float4 o;
                o.x = rstd * (gv.x * gs * ga.x - s1 - (xv.x - mean) * rstd * s2);
                o.y = rstd * (gv.y * gs * ga.y - s1 - (xv.y - mean) * rstd * s2);
                o.z = rstd * (gv.z * gs * ga.z - s1 - (xv.z - mean) * rstd * s2);
                o.w = rstd * (gv.w * gs * ga.w - s1 - (xv.w - mean) * rstd * s2);
                if (add) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(add + row * (int64_t)d + e));
                    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                }
                *reinterpret_cast<float4*>(dx + row * (int64_t)d + e) = o;
            }
        }
    }
    __syncthreads();
    float* part = partials + (int64_t)blockIdx.x * 2 * d;
    for (int e = threadIdx.x; e < 2 * d; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += lb_smem[(size_t)w * 2 * d + e];
        part[e] = s;
    }
}

// ------------------------------------------------------------------ activation / dropout backward (elementwise)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ hpre, const float* __restrict__ da, int act, float drop_p, uint64_t seed,
               uint64_t offset, int64_t total4, float* __restrict__ dh, float* __restrict__ a_out) {
    { const DrawKey key_ = rng_resolve(seed, offset); seed = key_.seed; offset = key_.offset; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        float h[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
        if (hpre) { const float4 v = __ldg(reinterpret_cast<const float4*>(hpre) + i); h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w; }
        if (da) { const float4 v = __ldg(reinterpret_cast<const float4*>(da) + i); g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w; }
        float od[4], oa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float m = drop_p > 0.f ? drop_keep_scale(seed, offset, (uint64_t)(i * 4 + j), drop_p) : 1.f;
            od[j] = g[j] * m * (hpre ? act_grad(act, h[j]) : 1.f);
            oa[j] = (hpre ? act_apply(act, h[j]) : 0.f) * m;
        }
        if (dh) reinterpret_cast<float4*>(dh)[i] = make_float4(od[0], od[1], od[2], od[3]);
        if (a_out) reinterpret_cast<float4*>(a_out)[i] = make_float4(oa[0], oa[1], oa[2], oa[3]);
    }
}

// ------------------------------------------------------------------ residual + dropout (stand-alone SublayerConnection)
// out = x + y * dropout_mask    (snuffy.py:108,110 when SublayerConnection.forward is called on its own)
__global__ void __launch_bounds__(256)
residual_dropout_kernel(const float* __restrict__ x, const float* __restrict__ y, float drop_p, uint64_t seed, uint64_t offset,
                        int64_t total, float* __restrict__ out) {
    { const DrawKey key_ = rng_resolve(seed, offset); seed = key_.seed; offset = key_.offset; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float m = drop_p > 0.f ? drop_keep_scale(seed, offset, (uint64_t)i, drop_p) : 1.f;
        out[i] = __ldg(x + i) + __ldg(y + i) * m;
    }
}

// ------------------------------------------------------------------ weighted column sums
// partial[chunk][c][e] = sum_{rows in chunk} w[row*C + c] * X[row*ldx + e]   (w == null -> weight 1, C = 1)
// grid (row chunks, column slabs of 256): thread = one column (coalesced row reads), 4 independent row accumulators
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, int64_t rows, int d, int C,
              int64_t rows_per_chunk, float* __restrict__ partials) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    const int e = blockIdx.y * 256 + threadIdx.x;
    if (e >= d) return;
    for (int c = 0; c < C; ++c) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int64_t r = r0;
        for (; r + 3 < r1; r += 4) {
            const float w0 = w ? __ldg(w + r * C + c) : 1.f, w1 = w ? __ldg(w + (r + 1) * C + c) : 1.f;
            const float w2 = w ? __ldg(w + (r + 2) * C + c) : 1.f, w3 = w ? __ldg(w + (r + 3) * C + c) : 1.f;
            a0 = fmaf(w0, __ldg(X + r * ldx + e), a0);
            a1 = fmaf(w1, __ldg(X + (r + 1) * ldx + e), a1);
            a2 = fmaf(w2, __ldg(X + (r + 2) * ldx + e), a2);
            a3 = fmaf(w3, __ldg(X + (r + 3) * ldx + e), a3);
        }
        for (; r < r1; ++r) a0 = fmaf(w ? __ldg(w + r * C + c) : 1.f, __ldg(X + r * ldx + e), a0);
        partials[((int64_t)blockIdx.x * C + c) * d + e] = (a0 + a1) + (a2 + a3);
    }
}

// Vector variant (d % 4 == 0, 16-byte aligned rows): CTA = 64 column quads x 4 row lanes, two float4 loads in flight per
// thread (the scalar kernel keeps too few bytes in flight to reach HBM speed), row lanes folded in a fixed order.
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, int64_t rows, int d, int C,
                  int64_t rows_per_chunk, float* __restrict__ partials) {
    __shared__ float4 red[4][64];
    const int q = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    const int e = blockIdx.y * 256 + q * 4;
    for (int c = 0; c < C; ++c) {
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (e < d) {
            int64_t r = r0 + rl;
            for (; r + 4 < r1; r += 8) {
                const float w0 = w ? __ldg(w + r * C + c) : 1.f, w1 = w ? __ldg(w + (r + 4) * C + c) : 1.f;
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(X + r * ldx + e));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(X + (r + 4) * ldx + e));
                a0.x = fmaf(w0, v0.x, a0.x); a0.y = fmaf(w0, v0.y, a0.y); a0.z = fmaf(w0, v0.z, a0.z); a0.w = fmaf(w0, v0.w, a0.w);
                a1.x = fmaf(w1, v1.x, a1.x); a1.y = fmaf(w1, v1.y, a1.y); a1.z = fmaf(w1, v1.z, a1.z); a1.w = fmaf(w1, v1.w, a1.w);
            }
            if (r < r1) {
                const float w0 = w ? __ldg(w + r * C + c) : 1.f;
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(X + r * ldx + e));
                a0.x = fmaf(w0, v0.x, a0.x); a0.y = fmaf(w0, v0.y, a0.y); a0.z = fmaf(w0, v0.z, a0.z); a0.w = fmaf(w0, v0.w, a0.w);
            }
        }
        red[rl][q] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
        __syncthreads();
        if (rl == 0 && e < d) {
            float4 s = red[0][q];
#pragma unroll
            for (int k = 1; k < 4; ++k) { s.x += red[k][q].x; s.y += red[k][q].y; s.z += red[k][q].z; s.w += red[k][q].w; }
            *reinterpret_cast<float4*>(partials + ((int64_t)blockIdx.x * C + c) * d + e) = s;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
fold_scalar_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] += __ldcg(part + (int64_t)(s + u) * n + i);
    }
    for (; s < splits; ++s) a[0] += __ldcg(part + (int64_t)s * n + i);
    out[i] = (a[0] + a[1]) + (a[2] + a[3]);
}

// ------------------------------------------------------------------ attention backward, row-local pieces
// S holds the scaled scores Q Kp^T / sqrt(dk) of every (bag, head): [B*h*N, Ksel] rows; stats [B*h*N, 2] = (max, 1/sum).
//   mode 0:  Pd[r, k] = exp(S - max) / sum * dropout_mask            (P~, the operand of dV = P~ dO)
//   mode 1:  G[r, k] <- (P~ G - P * sum_k(P~ G)) / sqrt(dk)          with G = V dO^T on entry: dS on exit
// warp per row.
__global__ void __launch_bounds__(256)
attn_rows_bwd_kernel(const float* __restrict__ S, const float* __restrict__ stats, int64_t nrows, int Ksel, int64_t N,
                     int mode, float inv_scale, float drop_p, uint64_t seed, uint64_t offset, float* __restrict__ Pd,
                     float* __restrict__ G) {
    { const DrawKey key_ = rng_resolve(seed, offset); seed = key_.seed; offset = key_.offset; }
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const float m = stats[row * 2], inv = stats[row * 2 + 1];
    const float* s = S + row * (int64_t)Ksel;
    // dropout index of element (b, j, n, key) = ((b*h + j)*N + n)*Ksel + key = row*Ksel + key  (matches the forward)
    const uint64_t base = (uint64_t)row * (uint64_t)Ksel;
    if (mode == 0) {
        for (int k = lane; k < Ksel; k += 32) {
            float p = expf(s[k] - m) * inv;
            if (drop_p > 0.f) p *= drop_keep_scale(seed, offset, base + k, drop_p);
            Pd[row * (int64_t)Ksel + k] = p;
        }
        return;
    }
    float* g = G + row * (int64_t)Ksel;
    float delta = 0.f;
    for (int k = lane; k < Ksel; k += 32) {
        float p = expf(s[k] - m) * inv;
        if (drop_p > 0.f) p *= drop_keep_scale(seed, offset, base + k, drop_p);
        delta = fmaf(p, g[k], delta);
    }
    delta = warp_sum(delta);
    for (int k = lane; k < Ksel; k += 32) {
        const float p = expf(s[k] - m) * inv;
        const float mk = drop_p > 0.f ? drop_keep_scale(seed, offset, base + k, drop_p) : 1.f;
        g[k] = p * (mk * g[k] - delta) * inv_scale;
    }
}

// ------------------------------------------------------------------ dx[b, idx[b,k], :] += src[b*K + k, :]
__global__ void __launch_bounds__(128)
scatter_add_rows_kernel(float* __restrict__ dx, const int64_t* __restrict__ idx, const float* __restrict__ src, int64_t N,
                        int64_t K, int d) {
    const int64_t slot = blockIdx.x;
    const int64_t b = slot / K;
    const int64_t r = idx[slot];
    if (r < 0 || r >= N) return;
    float* dst = dx + (b * N + r) * (int64_t)d;
    const float* s = src + slot * (int64_t)d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) dst[e] += s[e];
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// CTAs (= partial rows) snuffy_ln_rows_bwd uses; partials needs blocks * 2 * d floats
int64_t snuffy_ln_rows_bwd_blocks(int64_t rows) {
    int64_t b = (rows + 7) / 8;
    const int64_t cap = 4 * (int64_t)sm_count();           // 8 warps per CTA keep too few loads in flight: several CTAs per SM
    if (b > cap) b = cap;
    return b < 1 ? 1 : b;
}

// LayerNorm backward over rows of y = (row_map ? x with mapped rows from alt : x) with saved stats [rows, 2]:
//   dx[row] = add[row] + rstd * (g - mean(g) - z * mean(g z)),  g = dy * gamma, z = (y - mean) * rstd
//   dgamma_dbeta[0:d] = sum_rows dy * z,  dgamma_dbeta[d:2d] = sum_rows dy
// Upstream gradient: dy [rows, d], or (dy == null) the broadcast dy_bcast[row / rows_per_bag, :] * bscale.
int snuffy_ln_rows_bwd(const float* dy, const float* dy_bcast, int64_t rows_per_bag, float bscale, const float* x,
                       const int32_t* row_map, const float* alt, const float* stats, const float* gamma,
                       const float* add, int64_t rows, int64_t d, float* dx, float* dgamma_dbeta, float* partials,
                       cudaStream_t stream) {
    SNUFFY_REQUIRE((dy || dy_bcast) && x && stats && gamma && partials, "snuffy_ln_rows_bwd: null pointer");
    SNUFFY_REQUIRE(!row_map || alt, "snuffy_ln_rows_bwd: row_map given without the replacement rows");
    SNUFFY_REQUIRE(d % 4 == 0 && d >= 4 && rows >= 1 && (dy || rows_per_bag >= 1), "snuffy_ln_rows_bwd: needs d %% 4 == 0 (d=%lld)",
                   (long long)d);
    SNUFFY_REQUIRE((uintptr_t)x % 16 == 0 && (!dy || (uintptr_t)dy % 16 == 0) && (!dy_bcast || (uintptr_t)dy_bcast % 16 == 0) &&
                       (!alt || (uintptr_t)alt % 16 == 0) && (!add || (uintptr_t)add % 16 == 0) &&
                       (!dx || (uintptr_t)dx % 16 == 0) && (uintptr_t)gamma % 16 == 0,
                   "snuffy_ln_rows_bwd: pointers must be 16-byte aligned");
    int warps = 8;
    while (warps > 1 && (size_t)warps * 2 * d * 4 > 200 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 2 * d * 4;
    SNUFFY_REQUIRE(smem <= 200 * 1024, "snuffy_ln_rows_bwd: d=%lld too large", (long long)d);
    if (smem > 48 * 1024)
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&ln_rows_bwd_kernel), (int)smem));
    const int64_t blocks = snuffy_ln_rows_bwd_blocks(rows);
    ln_rows_bwd_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(dy, dy_bcast, rows_per_bag, bscale, x, row_map, alt,
                                                                      stats, gamma, add, rows, (int)d, dx, partials);
    if (!dgamma_dbeta) return check_launch("snuffy_ln_rows_bwd");     // the caller folds the partials (snuffy_fold_partials)
    launch_fold_partials(partials, (int)blocks, 2 * d / 4, dgamma_dbeta, stream);
    return check_launch("snuffy_ln_rows_bwd", 2);
}

// out[n] = sum over `splits` partial rows of part[splits, n] (n % 4 == 0), in a fixed order.  The deferred second half of
// snuffy_ln_rows_bwd (dgamma_dbeta == NULL there): the parameter gradients are not on the backward pass's critical chain, so a
// caller may issue this fold on another stream.
int snuffy_fold_partials(const float* partials, int64_t splits, int64_t n, float* out, cudaStream_t stream) {
    SNUFFY_REQUIRE(partials && out && splits >= 1 && n >= 4 && n % 4 == 0 && (uintptr_t)partials % 16 == 0 && (uintptr_t)out % 16 == 0,
                   "snuffy_fold_partials: bad arguments");
    launch_fold_partials(partials, (int)splits, n / 4, out, stream);
    return check_launch("snuffy_fold_partials");
}

// dh = da * mask * act'(hpre) and/or a_out = act(hpre) * mask over `total` contiguous elements (total % 4 == 0).
// hpre == null: plain dropout-mask application (act = identity).  The mask of element i is the one the forward GEMM
// epilogue drew for flat index i with the same (p, seed, offset).
int snuffy_act_bwd(const float* hpre, const float* da, int act, float dropout_p, uint64_t seed, uint64_t offset,
                   int64_t total, float* dh, float* a_out, cudaStream_t stream) {
    SNUFFY_REQUIRE((dh || a_out) && (!dh || da) && (!a_out || hpre), "snuffy_act_bwd: inconsistent pointers");
    SNUFFY_REQUIRE(total % 4 == 0, "snuffy_act_bwd: element count must be a multiple of 4");
    if (total == 0) return 0;
    const int64_t t4 = total / 4;
    int64_t blocks = (t4 + 255) / 256;
    const int64_t cap = 16 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    act_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(hpre, da, act, dropout_p, seed, offset, t4, dh, a_out);
    return check_launch("snuffy_act_bwd");
}

// out = x + y * dropout_mask over `total` contiguous elements (mask of element i as in snuffy_act_bwd with hpre == null)
int snuffy_residual_dropout(const float* x, const float* y, float dropout_p, uint64_t seed, uint64_t offset, int64_t total,
                            float* out, cudaStream_t stream) {
    SNUFFY_REQUIRE(x && y && out && total >= 0, "snuffy_residual_dropout: bad arguments");
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = 16 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    residual_dropout_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, y, dropout_p, seed, offset, total, out);
    return check_launch("snuffy_residual_dropout");
}

int64_t snuffy_colsum_chunks(int64_t rows) {
    int64_t c = (rows + 31) / 32;
    const int64_t cap = 2 * (int64_t)sm_count();
    if (c > cap) c = cap;
    return c < 1 ? 1 : c;
}

// out[c, e] = sum_rows w[row, c] * X[row, e]  (w == null: plain column sums, C must be 1).  partials: chunks*C*d floats.
int snuffy_colsum(const float* X, int64_t ldx, const float* w, int64_t rows, int64_t d, int64_t C, float* out,
                  float* partials, cudaStream_t stream) {
    SNUFFY_REQUIRE(X && out && partials && rows >= 1 && d >= 1 && C >= 1 && (w || C == 1), "snuffy_colsum: bad arguments");
    const int64_t chunks = snuffy_colsum_chunks(rows);
    const int64_t rpc = (rows + chunks - 1) / chunks;
    dim3 grid((unsigned)chunks, (unsigned)((d + 255) / 256));
    if (d % 4 == 0 && ldx % 4 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)partials % 16 == 0)
        colsum_vec_kernel<<<grid, 256, 0, stream>>>(X, ldx, w, rows, (int)d, (int)C, rpc, partials);
    else
        colsum_kernel<<<grid, 256, 0, stream>>>(X, ldx, w, rows, (int)d, (int)C, rpc, partials);
    const int64_t n = C * d;
    if (fold_wide_pays((int)chunks, n))
        fold_wide_kernel<float><<<(unsigned)((n + 15) / 16), 256, 0, stream>>>(partials, (int)chunks, n, out);
    else
        fold_scalar_kernel<<<(unsigned)((n + 63) / 64), 64, 0, stream>>>(partials, (int)chunks, n, out);
    return check_launch("snuffy_colsum", 2);
}

// Row-local pieces of the sparse-attention backward on materialised [B*h*N, Ksel] score matrices (see kernel comment).
int snuffy_attn_rows_bwd(const float* S, const float* stats, int64_t nrows, int64_t Ksel, int64_t N, int mode, float scale,
                         float dropout_p, uint64_t seed, uint64_t offset, float* Pd, float* G, cudaStream_t stream) {
    SNUFFY_REQUIRE(S && stats && nrows >= 1 && Ksel >= 1 && (mode == 0 ? Pd != nullptr : G != nullptr),
                   "snuffy_attn_rows_bwd: bad arguments");
    attn_rows_bwd_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, stream>>>(S, stats, nrows, (int)Ksel, N, mode, 1.f / scale,
                                                                         dropout_p, seed, offset, Pd, G);
    return check_launch("snuffy_attn_rows_bwd");
}

int snuffy_scatter_add_rows(float* dx, const int64_t* idx, const float* src, int64_t B, int64_t N, int64_t K, int64_t d,
                            cudaStream_t stream) {
    SNUFFY_REQUIRE(dx && idx && src && B >= 1 && N >= 1 && K >= 0 && d >= 1, "snuffy_scatter_add_rows: bad arguments");
    if (K == 0) return 0;
    scatter_add_rows_kernel<<<(unsigned)(B * K), 128, 0, stream>>>(dx, idx, src, N, K, (int)d);
    return check_launch("snuffy_scatter_add_rows");
}

}  // extern "C"
#pragma GCC visibility pop

// ------------------------------------------------------------------ DSMIL: softmax over instances, backward
// A [N, C] = softmax over n (dsmil.py:86).  dS[n, c] = A[n, c] * (dA[n, c] - sum_n' A[n', c] dA[n', c]) / scale.
// One CTA per class (C is 1..3 and this is N*C elements: latency-sized).
namespace snuffy {
__global__ void __launch_bounds__(1024)
softmax_cols_bwd_kernel(const float* __restrict__ A, const float* __restrict__ dA, int64_t N, int C, float inv_scale,
                        float* __restrict__ dS) {
    __shared__ float red[32];
    __shared__ float s_delta;
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) acc = fmaf(A[n * C + c], dA[n * C + c], acc);
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        float v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) s_delta = v;
    }
    __syncthreads();
    const float delta = s_delta;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x)
        dS[n * C + c] = A[n * C + c] * (dA[n * C + c] - delta) * inv_scale;
}
}  // namespace snuffy

#pragma GCC visibility push(default)
extern "C" int snuffy_softmax_cols_bwd(const float* A, const float* dA, int64_t N, int64_t C, float scale, float* dS,
                                       cudaStream_t stream) {
    SNUFFY_REQUIRE(A && dA && dS && N >= 1 && C >= 1 && scale > 0.f, "snuffy_softmax_cols_bwd: bad arguments");
    snuffy::softmax_cols_bwd_kernel<<<(unsigned)C, 1024, 0, stream>>>(A, dA, N, (int)C, 1.f / scale, dS);
    return snuffy::check_launch("snuffy_softmax_cols_bwd");
}
#pragma GCC visibility pop


// ------------------------------------------------------------------ attention backward on tensor cores: head-block operands
// All heads of one bag are contracted by ONE dense tcgen05 GEMM against a block-structured operand
//   Kbd[(j, k), c] = Kp[k, c] if column c belongs to head j else 0           [h*Ksel, d]
// so that S_all[n, (j, k)] = Q[n, :] . Kbd[(j, k), :] = Q_j[n] . Kp_j[k].  The zero blocks cost 8x the useful FLOPs, which
// the tensor cores absorb (35 us per product at cfg2) where the head-batched SIMT kernels took 200 us each.
namespace snuffy {
__global__ void __launch_bounds__(256)
block_diag_rows_kernel(const float* __restrict__ src, int Ksel, int h, int d, float* __restrict__ out) {
    const int64_t total4 = (int64_t)h * Ksel * d / 4;
    const int dk = d / h;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % d);
        const int64_t r = (i * 4) / d;                     // row (j, k)
        const int j = (int)(r / Ksel), k = (int)(r % Ksel);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c / dk == j) v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)k * d + c));
        reinterpret_cast<float4*>(out)[i] = v;
    }
}
__global__ void __launch_bounds__(256)
block_diag_extract_kernel(const float* __restrict__ bd, int Ksel, int h, int d, float* __restrict__ out) {
    const int64_t total = (int64_t)Ksel * d;
    const int dk = d / h;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % d), k = (int)(i / d), j = c / dk;
    out[i] = bd[((int64_t)j * Ksel + k) * d + c];
}

// Row-local pieces on the all-heads layout S_all [N, h*Ksel] (segment j of row n = head j), RAW scores (unscaled):
//   mode 0:  Pd = exp(S/scale - max) / sum * dropout_mask          mode 1:  G <- (Pd G - P sum_k(Pd G)) / scale
// stats [h, N, 2] of this bag; the dropout index of (bag, j, n, key) matches the forward: ((bag*h + j)*N + n)*Ksel + key.
__global__ void __launch_bounds__(256)
attn_seg_bwd_kernel(const float* __restrict__ S, const float* __restrict__ stats, int64_t N, int h, int Ksel, int bag, int mode,
                    float inv_scale, float drop_p, uint64_t seed, uint64_t offset, float* __restrict__ Pd, float* __restrict__ G) {
    { const DrawKey key_ = rng_resolve(seed, offset); seed = key_.seed; offset = key_.offset; }
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (n, j)
    if (seg >= N * h) return;
    const int64_t n = seg / h;
    const int j = (int)(seg % h);
    const int64_t srow = (int64_t)j * N + n;
    const float m = stats[srow * 2], inv = stats[srow * 2 + 1];
    const int64_t off = n * (int64_t)h * Ksel + (int64_t)j * Ksel;
    const uint64_t base = ((uint64_t)((int64_t)bag * h + j) * (uint64_t)N + (uint64_t)n) * (uint64_t)Ksel;
    const float* s = S + off;
    if (mode == 0) {
        for (int k = lane; k < Ksel; k += 32) {
            float p = expf(s[k] * inv_scale - m) * inv;
            if (drop_p > 0.f) p *= drop_keep_scale(seed, offset, base + k, drop_p);
            Pd[off + k] = p;
        }
        return;
    }
    float* g = G + off;
    float delta = 0.f;
    for (int k = lane; k < Ksel; k += 32) {
        float p = expf(s[k] * inv_scale - m) * inv;
        if (drop_p > 0.f) p *= drop_keep_scale(seed, offset, base + k, drop_p);
        delta = fmaf(p, g[k], delta);
    }
    delta = warp_sum(delta);
    for (int k = lane; k < Ksel; k += 32) {
        const float p = expf(s[k] * inv_scale - m) * inv;
        const float mk = drop_p > 0.f ? drop_keep_scale(seed, offset, base + k, drop_p) : 1.f;
        g[k] = p * (mk * g[k] - delta) * inv_scale;
    }
}

// Same math, one 8-key unit per lane (Ksel % 8 == 0, Ksel <= 256): S and G are read once (the probabilities stay in registers
// between the two passes of mode 1) and the result is ALSO written as split-bf16 A-operand planes over [rows_pad, h*Ksel] — the
// operand of the next tensor-core product (dV = P~ dObd, dQ = dS Kbd) — instead of a separate fp32 -> planes pass.  Warps of a CTA
// take consecutive rows of one head, so the 16-byte plane units of a CTA are contiguous.  Rows [N, rows_pad) are zero-filled.
__global__ void __launch_bounds__(256)
attn_seg_bwd_planes_kernel(const float* __restrict__ S, const float* __restrict__ stats, int64_t N, int64_t rows_pad, int h,
                           int Ksel, int bag, int mode, float inv_scale, float drop_p, uint64_t seed, uint64_t offset,
                           float* __restrict__ Pd, float* __restrict__ G, __nv_bfloat16* __restrict__ planes,
                           int64_t plane_stride) {
    { const DrawKey key_ = rng_resolve(seed, offset); seed = key_.seed; offset = key_.offset; }
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (j, n), n fastest
    if (seg >= rows_pad * h) return;
    const int j = (int)(seg / rows_pad);
    const int64_t n = seg % rows_pad;
    const bool active = lane * 8 < Ksel;
    const int64_t K = (int64_t)h * Ksel;
    float out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = 0.f;
    if (n < N) {                                                                          // warp-uniform
        const int64_t srow = (int64_t)j * N + n;
        const float m = stats[srow * 2], inv = stats[srow * 2 + 1];
        const int64_t off = n * K + (int64_t)j * Ksel + lane * 8;
        const uint64_t base = ((uint64_t)((int64_t)bag * h + j) * (uint64_t)N + (uint64_t)n) * (uint64_t)Ksel + (uint64_t)(lane * 8);
        float p[8], pd[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { p[e] = 0.f; pd[e] = 0.f; }
        if (active) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(S + off)), s1 = __ldg(reinterpret_cast<const float4*>(S + off + 4));
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                p[e] = expf(sv[e] * inv_scale - m) * inv;
                pd[e] = drop_p > 0.f ? p[e] * drop_keep_scale(seed, offset, base + e, drop_p) : p[e];
            }
        }
        if (mode == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) out[e] = pd[e];
            if (Pd && active) {
                *reinterpret_cast<float4*>(Pd + off) = make_float4(pd[0], pd[1], pd[2], pd[3]);
                *reinterpret_cast<float4*>(Pd + off + 4) = make_float4(pd[4], pd[5], pd[6], pd[7]);
            }
        } else {
            float gv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (active) {
                const float4 g0 = *reinterpret_cast<const float4*>(G + off), g1 = *reinterpret_cast<const float4*>(G + off + 4);
                gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
            }
            float delta = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) delta = fmaf(pd[e], gv[e], delta);
            delta = warp_sum(delta);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                // pd = p * mask  ->  p * (mask * g - delta) = pd * g - p * delta
                out[e] = (pd[e] * gv[e] - p[e] * delta) * inv_scale;
            }
            if (active) {
                *reinterpret_cast<float4*>(G + off) = make_float4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<float4*>(G + off + 4) = make_float4(out[4], out[5], out[6], out[7]);
            }
        }
    }
    if (active) {
        bf16x8 hi, lo;
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16(out[e], hi.v[e], lo.v[e]);
        __nv_bfloat16* dst = planes + plane_unit_offset(n, (int64_t)j * Ksel + lane * 8, K, 128);
        *reinterpret_cast<bf16x8*>(dst) = hi;
        *reinterpret_cast<bf16x8*>(dst + plane_stride) = lo;
    }
}
}  // namespace snuffy

#pragma GCC visibility push(default)
extern "C" {
// out [h*Ksel, d]: row (j, k) = src[k, :] restricted to the columns of head j (zeros elsewhere); src [Ksel, d]
int snuffy_block_diag_rows(const float* src, int64_t Ksel, int64_t h, int64_t d, float* out, cudaStream_t stream) {
    SNUFFY_REQUIRE(src && out && Ksel >= 1 && h >= 1 && d % h == 0 && (d / h) % 4 == 0 && (uintptr_t)src % 16 == 0 &&
                       (uintptr_t)out % 16 == 0, "snuffy_block_diag_rows: bad arguments");
    const int64_t total4 = h * Ksel * d / 4;
    int64_t blocks = (total4 + 255) / 256;
    if (blocks > 8 * (int64_t)snuffy::sm_count()) blocks = 8 * (int64_t)snuffy::sm_count();
    snuffy::block_diag_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, (int)Ksel, (int)h, (int)d, out);
    return snuffy::check_launch("snuffy_block_diag_rows");
}
// out [Ksel, d]: out[k, c] = bd[(head(c), k), c]   (the diagonal blocks of a [h*Ksel, d] matrix)
int snuffy_block_diag_extract(const float* bd, int64_t Ksel, int64_t h, int64_t d, float* out, cudaStream_t stream) {
    SNUFFY_REQUIRE(bd && out && Ksel >= 1 && h >= 1 && d % h == 0, "snuffy_block_diag_extract: bad arguments");
    snuffy::block_diag_extract_kernel<<<(unsigned)((Ksel * d + 255) / 256), 256, 0, stream>>>(bd, (int)Ksel, (int)h, (int)d, out);
    return snuffy::check_launch("snuffy_block_diag_extract");
}
int snuffy_attn_seg_bwd(const float* S, const float* stats, int64_t N, int64_t h, int64_t Ksel, int64_t bag, int mode,
                        float scale, float dropout_p, uint64_t seed, uint64_t offset, float* Pd, float* G, void* planes,
                        int64_t plane_stride, cudaStream_t stream) {
    SNUFFY_REQUIRE(S && stats && N >= 1 && h >= 1 && Ksel >= 1 && (mode == 0 || mode == 1), "snuffy_attn_seg_bwd: bad arguments");
    if (planes) {
        SNUFFY_REQUIRE(Ksel % 8 == 0 && Ksel <= 256 && (h * Ksel) % 32 == 0 && (mode == 0 || G != nullptr) &&
                           plane_stride >= snuffy::plane_elems(N, h * Ksel, 128) && (uintptr_t)S % 16 == 0 &&
                           (uintptr_t)planes % 16 == 0 && (!G || (uintptr_t)G % 16 == 0) && (!Pd || (uintptr_t)Pd % 16 == 0),
                       "snuffy_attn_seg_bwd: plane output needs Ksel %% 8 == 0, Ksel <= 256, (h*Ksel) %% 32 == 0 and "
                       "16-byte aligned buffers");
        const int64_t rows_pad = (N + 127) / 128 * 128;
        snuffy::attn_seg_bwd_planes_kernel<<<(unsigned)((rows_pad * h + 7) / 8), 256, 0, stream>>>(
            S, stats, N, rows_pad, (int)h, (int)Ksel, (int)bag, mode, 1.f / scale, dropout_p, seed, offset, Pd, G,
            reinterpret_cast<__nv_bfloat16*>(planes), plane_stride);
        return snuffy::check_launch("snuffy_attn_seg_bwd");
    }
    SNUFFY_REQUIRE(mode == 0 ? Pd != nullptr : G != nullptr, "snuffy_attn_seg_bwd: null output");
    snuffy::attn_seg_bwd_kernel<<<(unsigned)((N * h + 7) / 8), 256, 0, stream>>>(S, stats, N, (int)h, (int)Ksel, (int)bag, mode,
                                                                               1.f / scale, dropout_p, seed, offset, Pd, G);
    return snuffy::check_launch("snuffy_attn_seg_bwd");
}
}  // extern "C"
#pragma GCC visibility pop
