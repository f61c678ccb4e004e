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
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbn,
                 int64_t sbk, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K, bool vecA,
                 bool vecB, SimtEpilogue ep) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + SG_PAD];
    __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + SG_PAD];
    const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    load_tile<A_KC>(A, sam, sak, m0, M, 0, K, vecA, ra);
    load_tile<B_KC>(B, sbn, sbk, n0, N, 0, K, vecB, rb);
    store_tile<A_KC>(As[0], ra);
    store_tile<B_KC>(Bs[0], rb);
    __syncthreads();

    const int64_t ktiles = (K + SG_BK - 1) / SG_BK;
    for (int64_t kt = 0; kt < ktiles; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < ktiles) {
            load_tile<A_KC>(A, sam, sak, m0, M, (kt + 1) * SG_BK, K, vecA, ra);
            load_tile<B_KC>(B, sbn, sbk, n0, N, (kt + 1) * SG_BK, K, vecB, rb);
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
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] *= drop_keep_scale(ep.seed, ep.offset, (uint64_t)(m * N + n + j), ep.drop_p);
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
                        if (ep.drop_p > 0.f) o *= drop_keep_scale(ep.seed, ep.offset, (uint64_t)(m * N + n + j), ep.drop_p);
                        if (rrow) o += rrow[n + j];
                        C[m * ldc + n + j] = o;
                    }
                }
            }
        }
    }
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

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
    const int64_t sam = a_kc ? lda : 1, sak = a_kc ? 1 : lda;
    const int64_t sbn = b_kc ? ldb : 1, sbk = b_kc ? 1 : ldb;
    const bool vecA = (lda % 4 == 0) && ((uintptr_t)A % 16 == 0);
    const bool vecB = (ldb % 4 == 0) && ((uintptr_t)B % 16 == 0);
    dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM));
    SNUFFY_REQUIRE(grid.y <= 65535, "snuffy_gemm_f32: M=%lld too large for one launch", (long long)M);
    if (a_kc && b_kc)
        gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep);
    else if (a_kc && !b_kc)
        gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep);
    else if (!a_kc && b_kc)
        gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep);
    else
        gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, vecA, vecB, ep);
    return check_launch("snuffy_gemm_f32");
}

}  // extern "C"
#pragma GCC visibility pop
