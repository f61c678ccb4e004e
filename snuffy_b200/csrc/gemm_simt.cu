// fp32 SIMT GEMM: the exact-fp32 path (precision mode "fp32"), the small [Ksel x d] projections
// (key / output projection, snuffy.py:188,205), DSMIL's q-MLP, and every backward contraction.
//
//   C[m, n] = epilogue( sum_k A(m, k) * B(n, k) )
//   A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]   (one of the two strides of each operand is 1)
//   epilogue: + bias[n] -> activation -> + residual row (optionally redirected through row_map)
//
// 128x128x16 tiles, 256 threads, 8x8 micro-tile per thread, register-prefetched double buffering.
#include "common.cuh"

namespace snuffy {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_PAD = 4;

struct SimtEpilogue {
    const float* bias;        // [N] or null
    int act;                  // Act id
    const float* resid;       // [M, ldr] or null, added after the activation
    const int32_t* row_map;   // optional: resid row m comes from resid_alt[row_map[m]] when >= 0
    const float* resid_alt;
    int64_t ldr;
    float* preact;            // optional [M, ldc]: value before the activation (saved for backward)
    float alpha;              // scales the accumulator before bias
    float drop_p; uint64_t seed, offset;   // dropout on the activated value (before the residual)
};

// Batched / split-K extension (backward contractions): blockIdx.z = batch * ksplit + split, batch = (outer, inner)
// with separate element strides per operand (e.g. outer = bag, inner = head slice of a [rows, d] matrix).  With
// ksplit > 1 every split writes its raw alpha-scaled partial to a dense [ksplit][nbatch][M][N] workspace that
// splitk_fold_kernel sums in a fixed order (deterministic), so long contractions over the N patches (dW = dY^T X,
// dKp = dS^T Q) fill the machine instead of a handful of CTAs.
struct SimtBatch {
    int nb_inner, ksplit;
    int64_t sa_o, sa_i, sb_o, sb_i, sc_o, sc_i;
    int64_t k_per_split;
    float* part;
};

// K_CONTIG: the operand's k stride is 1 (row-major [rows, K]); otherwise its row stride is 1.
template <bool K_CONTIG>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t s_row, int64_t s_k, int64_t row0,
                                          int64_t nrows, int64_t k0, int64_t K, bool vec, float (&reg)[8]) {
    const int t = threadIdx.x;
    if (K_CONTIG) {
        // tile [128 rows][16 k]: thread -> rows t/4 and t/4+64, k quad t%4
        const int kq = (t & 3) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t r = row0 + (t >> 2) + 64 * i;
            const int64_t k = k0 + kq;
            if (r < nrows && vec && k + 3 < K) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(P + r * s_row + k));
                reg[i * 4 + 0] = v.x; reg[i * 4 + 1] = v.y; reg[i * 4 + 2] = v.z; reg[i * 4 + 3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    reg[i * 4 + j] = (r < nrows && k + j < K) ? __ldg(P + r * s_row + (k + j) * s_k) : 0.f;
            }
        }
    } else {
        // tile [16 k][128 rows], rows contiguous: thread -> k t/32 and t/32+8, row quad t%32
        const int rq = (t & 31) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t k = k0 + (t >> 5) + 8 * i;
            const int64_t r = row0 + rq;
            if (k < K && vec && r + 3 < nrows) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(P + k * s_k + r));
                reg[i * 4 + 0] = v.x; reg[i * 4 + 1] = v.y; reg[i * 4 + 2] = v.z; reg[i * 4 + 3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    reg[i * 4 + j] = (k < K && r + j < nrows) ? __ldg(P + k * s_k + (r + j) * s_row) : 0.f;
            }
        }
    }
}

template <bool K_CONTIG>
__device__ __forceinline__ void store_tile(float (*S)[SG_BM + SG_PAD], const float (&reg)[8]) {
    const int t = threadIdx.x;
    if (K_CONTIG) {
        const int kq = (t & 3) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) S[kq + j][(t >> 2) + 64 * i] = reg[i * 4 + j];
    } else {
        const int rq = (t & 31) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            *reinterpret_cast<float4*>(&S[(t >> 5) + 8 * i][rq]) =
                make_float4(reg[i * 4 + 0], reg[i * 4 + 1], reg[i * 4 + 2], reg[i * 4 + 3]);
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2)     // <= 128 registers: two CTAs per SM hide the short-K latency of the backward GEMMs
gemm_simt_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbn,
                 int64_t sbk, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K, bool vecA,
                 bool vecB, SimtEpilogue ep, SimtBatch bt) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + SG_PAD];
    __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + SG_PAD];
    const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
    const int split = (int)(blockIdx.z % bt.ksplit), zb = (int)(blockIdx.z / bt.ksplit);
    const int zo = zb / bt.nb_inner, zi = zb % bt.nb_inner;
    A += zo * bt.sa_o + zi * bt.sa_i;
    B += zo * bt.sb_o + zi * bt.sb_i;
    C += zo * bt.sc_o + zi * bt.sc_i;
    const int64_t k_begin = (int64_t)split * bt.k_per_split;
    if (bt.ksplit > 1) K = min(K, k_begin + bt.k_per_split);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    load_tile<A_KC>(A, sam, sak, m0, M, k_begin, K, vecA, ra);
    load_tile<B_KC>(B, sbn, sbk, n0, N, k_begin, K, vecB, rb);
    store_tile<A_KC>(As[0], ra);
    store_tile<B_KC>(Bs[0], rb);
    __syncthreads();

    const int64_t ktiles = K > k_begin ? (K - k_begin + SG_BK - 1) / SG_BK : 0;
    for (int64_t kt = 0; kt < ktiles; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < ktiles) {
            load_tile<A_KC>(A, sam, sak, m0, M, k_begin + (kt + 1) * SG_BK, K, vecA, ra);
            load_tile<B_KC>(B, sbn, sbk, n0, N, k_begin + (kt + 1) * SG_BK, K, vecB, rb);
        }
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            store_tile<A_KC>(As[cur ^ 1], ra);
            store_tile<B_KC>(Bs[cur ^ 1], rb);
        }
        __syncthreads();
    }

    if (bt.ksplit > 1) {
        // raw partial: part[split][batch][m][n]
        float* P = bt.part + ((int64_t)split * (gridDim.z / bt.ksplit) + zb) * M * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (m >= M) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t n = n0 + h * 64 + tx * 4 + j;
                    if (n < N) P[m * N + n] = acc[i][h * 4 + j] * ep.alpha;
                }
        }
        return;
    }
    const bool vec_out = (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0) &&
                         (!ep.resid || ((ep.ldr % 4 == 0) && ((uintptr_t)ep.resid % 16 == 0) &&
                                        (!ep.resid_alt || (uintptr_t)ep.resid_alt % 16 == 0))) &&
                         (!ep.preact || (uintptr_t)ep.preact % 16 == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
        const float* rrow = nullptr;
        if (ep.resid) {
            rrow = ep.resid + m * ep.ldr;
            if (ep.row_map) { const int32_t slot = ep.row_map[m]; if (slot >= 0) rrow = ep.resid_alt + (int64_t)slot * ep.ldr; }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t n = n0 + h * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = acc[i][h * 4 + j] * ep.alpha;
                if (ep.bias && n + j < N) s += ep.bias[n + j];
                v[j] = s;
            }
            if (vec_out && n + 3 < N) {
                if (ep.preact) *reinterpret_cast<float4*>(ep.preact + m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = act_apply(ep.act, v[j]);
                if (ep.drop_p > 0.f) {
                    const DrawKey key = rng_resolve(ep.seed, ep.offset);
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] *= drop_keep_scale(key.seed, key.offset, (uint64_t)(m * N + n + j), ep.drop_p);
                }
                if (rrow) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(rrow + n));
                    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
                }
                *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < N) {
                        if (ep.preact) ep.preact[m * ldc + n + j] = v[j];
                        float o = act_apply(ep.act, v[j]);
                        if (ep.drop_p > 0.f) {
                            const DrawKey key = rng_resolve(ep.seed, ep.offset);
                            o *= drop_keep_scale(key.seed, key.offset, (uint64_t)(m * N + n + j), ep.drop_p);
                        }
                        if (rrow) o += rrow[n + j];
                        C[m * ldc + n + j] = o;
                    }
                }
            }
        }
    }
}

// C[batch][m][n] (strided) = sum_split part[split][batch][m][n] + bias[n]
__global__ void __launch_bounds__(256)
splitk_fold_kernel(const float* __restrict__ part, int ksplit, int nbatch, int64_t M, int64_t N, float* __restrict__ C,
                   int64_t ldc, int nb_inner, int64_t sc_o, int64_t sc_i, const float* __restrict__ bias) {
    const int64_t total = (int64_t)nbatch * M * N;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s) acc += __ldcg(part + (int64_t)s * total + i);
    const int64_t n = i % N, m = (i / N) % M;
    const int zb = (int)(i / (M * N));
    if (bias) acc += bias[n];
    C[(zb / nb_inner) * sc_o + (zb % nb_inner) * sc_i + m * ldc + n] = acc;
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

static int launch_gemm_f32(const float* A, int64_t lda, int a_kc, const float* B, int64_t ldb, int b_kc, float* C,
                           int64_t ldc, int64_t M, int64_t N, int64_t K, const SimtEpilogue& ep, const SimtBatch& bt,
                           int nbatch, cudaStream_t stream, const char* who) {
    const int64_t sam = a_kc ? lda : 1, sak = a_kc ? 1 : lda;
    const int64_t sbn = b_kc ? ldb : 1, sbk = b_kc ? 1 : ldb;
    const bool vecA = (lda % 4 == 0) && ((uintptr_t)A % 16 == 0) && bt.sa_o % 4 == 0 && bt.sa_i % 4 == 0;
    const bool vecB = (ldb % 4 == 0) && ((uintptr_t)B % 16 == 0) && bt.sb_o % 4 == 0 && bt.sb_i % 4 == 0;
    dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)(nbatch * bt.ksplit));
    SNUFFY_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: M=%lld / batch too large for one launch", who, (long long)M);
    if (a_kc && b_kc)
        gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep, bt);
    else if (a_kc && !b_kc)
        gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep, bt);
    else if (!a_kc && b_kc)
        gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep, bt);
    else
        gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep, bt);
    return 0;
}

// C[M,N] (ldc) = dropout(act(alpha * A.B^T + bias)) + resid.   a_kc / b_kc: 1 when the operand is stored [rows, K]
// row-major with leading dimension lda/ldb, 0 when it is stored [K, rows] row-major (i.e. transposed).
int snuffy_gemm_f32(const float* A, int64_t lda, int a_kc, const float* B, int64_t ldb, int b_kc, float* C,
                    int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, const float* bias, int act,
                    const float* resid, int64_t ldr, const int32_t* row_map, const float* resid_alt, float* preact,
                    float dropout_p, uint64_t seed, uint64_t offset, cudaStream_t stream) {
    SNUFFY_REQUIRE(A && B && C, "snuffy_gemm_f32: null pointer");
    SNUFFY_REQUIRE(M >= 0 && N >= 0 && K >= 0, "snuffy_gemm_f32: negative dimension");
    SNUFFY_REQUIRE(!row_map || (resid && resid_alt), "snuffy_gemm_f32: row_map needs resid and resid_alt");
    if (M == 0 || N == 0) return 0;
    SimtEpilogue ep{bias, act, resid, row_map, resid_alt, ldr, preact, alpha, dropout_p, seed, offset};
    SimtBatch bt{1, 1, 0, 0, 0, 0, 0, 0, 0, nullptr};
    if (int rc = launch_gemm_f32(A, lda, a_kc, B, ldb, b_kc, C, ldc, M, N, K, ep, bt, 1, stream, "snuffy_gemm_f32")) return rc;
    return check_launch("snuffy_gemm_f32");
}

// Batched / split-K form used by the backward pass: for batch z = (zo, zi), zo < nb_outer, zi < nb_inner
//   C_z[M,N] = alpha * A_z . B_z^T (+ bias),   X_z = X + zo * s?_o + zi * s?_i   (element strides)
// ksplit > 1 splits the contraction over K across CTAs; `workspace` then needs snuffy_gemm_f32_batched_workspace bytes.
int64_t snuffy_gemm_f32_batched_workspace(int64_t nbatch, int64_t M, int64_t N, int64_t ksplit) {
    return ksplit > 1 ? nbatch * M * N * ksplit * 4 + 16 : 0;
}

// a ksplit that fills ~2 waves of the machine for this problem (1 when the grid is already large or K is short)
int64_t snuffy_gemm_f32_auto_ksplit(int64_t nbatch, int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = nbatch * ((M + SG_BM - 1) / SG_BM) * ((N + SG_BN - 1) / SG_BN);
    if (tiles <= 0) return 1;
    int64_t ks = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
    const int64_t max_ks = K / 256;                        // at least 256 k per split
    if (ks > max_ks) ks = max_ks;
    if (ks > 64) ks = 64;
    return ks < 1 ? 1 : ks;
}

int snuffy_gemm_f32_batched(const float* A, int64_t lda, int a_kc, const float* B, int64_t ldb, int b_kc, float* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, const float* bias,
                            int64_t nb_outer, int64_t nb_inner, int64_t sa_o, int64_t sa_i, int64_t sb_o, int64_t sb_i,
                            int64_t sc_o, int64_t sc_i, int64_t ksplit, void* workspace, int64_t workspace_bytes,
                            cudaStream_t stream) {
    SNUFFY_REQUIRE(A && B && C, "snuffy_gemm_f32_batched: null pointer");
    SNUFFY_REQUIRE(M >= 0 && N >= 0 && K >= 0 && nb_outer >= 1 && nb_inner >= 1 && ksplit >= 1,
                   "snuffy_gemm_f32_batched: bad dimensions");
    if (M == 0 || N == 0) return 0;
    const int64_t nbatch = nb_outer * nb_inner;
    SNUFFY_REQUIRE(ksplit == 1 || (workspace && workspace_bytes >= snuffy_gemm_f32_batched_workspace(nbatch, M, N, ksplit)),
                   "snuffy_gemm_f32_batched: split-K needs a workspace");
    SimtEpilogue ep{ksplit == 1 ? bias : nullptr, ACT_NONE, nullptr, nullptr, nullptr, 0, nullptr, alpha, 0.f, 0, 0};
    SimtBatch bt{(int)nb_inner, (int)ksplit, sa_o, sa_i, sb_o, sb_i, sc_o, sc_i, 0, reinterpret_cast<float*>(workspace)};
    if (ksplit > 1) {
        int64_t kps = (K + ksplit - 1) / ksplit;
        bt.k_per_split = (kps + SG_BK - 1) / SG_BK * SG_BK;
    }
    if (int rc = launch_gemm_f32(A, lda, a_kc, B, ldb, b_kc, C, ldc, M, N, K, ep, bt, (int)nbatch, stream,
                                 "snuffy_gemm_f32_batched")) return rc;
    if (ksplit > 1) {
        const int64_t total = nbatch * M * N;
        splitk_fold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(bt.part, (int)ksplit, (int)nbatch, M, N, C, ldc,
                                                                               (int)nb_inner, sc_o, sc_i, bias);
        return check_launch("snuffy_gemm_f32_batched", 2);
    }
    return check_launch("snuffy_gemm_f32_batched");
}

}  // extern "C"
#pragma GCC visibility pop
