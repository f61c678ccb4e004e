// DSMIL bag classifier (dsmil.py:72-92):
//   S[n,c] = Q[n,:] . q_max[c,:] / sqrt(128)   ;  A = softmax over the INSTANCE axis n (per class)
//   Bm[c,:] = sum_n A[n,c] V[n,:]              ;  logits = Conv1d(C, C, kernel=d)(Bm)
// Q / V / q_max come from the shared GEMM; the critical instance (arg-max row per class) from
// snuffy_select_topk with K = 1.  Two streaming passes over N, deterministic reductions.
#include "common.cuh"

namespace snuffy {

constexpr int DS_MAXC = 8;

// pass 1: raw scaled scores into A, per-CTA (max, sum-exp) partials per class.  grid.x = chunks
__global__ void __launch_bounds__(256)
dsmil_scores_kernel(const float* __restrict__ Q, const float* __restrict__ qmax, int64_t N, int dq, int C,
                    float scale, float* __restrict__ A, float* __restrict__ stat_part) {
    extern __shared__ __align__(16) float ds_smem[];       // qmax [C][dq]
    __shared__ float red_m[8][DS_MAXC], red_l[8][DS_MAXC];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < C * dq; i += 256) ds_smem[i] = qmax[i];
    __syncthreads();
    const int chunks = gridDim.x;
    const int64_t per = (N + chunks - 1) / chunks;
    const int64_t r0 = blockIdx.x * per, r1 = min(N, r0 + per);
    float m[DS_MAXC], l[DS_MAXC];
#pragma unroll
    for (int c = 0; c < DS_MAXC; ++c) { m[c] = -INFINITY; l[c] = 0.f; }
    for (int64_t n = r0 + warp; n < r1; n += 8) {
        float acc[DS_MAXC];
#pragma unroll
        for (int c = 0; c < DS_MAXC; ++c) acc[c] = 0.f;
        for (int e = lane; e < dq; e += 32) {
            const float qv = Q[n * dq + e];
#pragma unroll
            for (int c = 0; c < DS_MAXC; ++c)
                if (c < C) acc[c] = fmaf(qv, ds_smem[c * dq + e], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < DS_MAXC; ++c) {
            if (c < C) {
                const float s = warp_sum(acc[c]) / scale;
                if (lane == 0) A[n * C + c] = s;
                const float mn = fmaxf(m[c], s);
                l[c] = l[c] * expf(m[c] - mn) + expf(s - mn);
                m[c] = mn;
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < DS_MAXC; ++c) { red_m[warp][c] = m[c]; red_l[warp][c] = l[c]; }
    }
    __syncthreads();
    if (t < C) {
        float mg = -INFINITY, lg = 0.f;
        for (int w = 0; w < 8; ++w) mg = fmaxf(mg, red_m[w][t]);
        for (int w = 0; w < 8; ++w) lg += (red_m[w][t] == -INFINITY) ? 0.f : red_l[w][t] * expf(red_m[w][t] - mg);
        stat_part[((int64_t)blockIdx.x * C + t) * 2] = mg;
        stat_part[((int64_t)blockIdx.x * C + t) * 2 + 1] = lg;
    }
}

// pass 2: normalise A in place, Bm partial = sum over this CTA's rows.  grid.x = chunks (same split)
__global__ void __launch_bounds__(256)
dsmil_pool_kernel(const float* __restrict__ V, int64_t N, int d, int C, int stat_chunks,
                  const float* __restrict__ stat_part, float* __restrict__ A, float* __restrict__ B_part,
                  float* __restrict__ stats_out) {
    __shared__ float s_m[DS_MAXC], s_inv[DS_MAXC];
    const int t = threadIdx.x;
    if (t < C) {
        float mg = -INFINITY, lg = 0.f;
        for (int k = 0; k < stat_chunks; ++k) mg = fmaxf(mg, stat_part[((int64_t)k * C + t) * 2]);
        for (int k = 0; k < stat_chunks; ++k) {
            const float mk = stat_part[((int64_t)k * C + t) * 2];
            if (mk != -INFINITY) lg += stat_part[((int64_t)k * C + t) * 2 + 1] * expf(mk - mg);
        }
        s_m[t] = mg; s_inv[t] = 1.f / lg;
        if (stats_out && blockIdx.x == 0) { stats_out[t * 2] = mg; stats_out[t * 2 + 1] = 1.f / lg; }
    }
    __syncthreads();
    const int chunks = gridDim.x;
    const int64_t per = (N + chunks - 1) / chunks;
    const int64_t r0 = blockIdx.x * per, r1 = min(N, r0 + per);
    // thread owns columns t + 256*i
    for (int e0 = 0; e0 < d; e0 += 256 * 4) {
        float acc[DS_MAXC][4];
#pragma unroll
        for (int c = 0; c < DS_MAXC; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
        for (int64_t n = r0; n < r1; ++n) {
            float a[DS_MAXC];
#pragma unroll
            for (int c = 0; c < DS_MAXC; ++c) a[c] = (c < C) ? expf(__ldg(A + n * C + c) - s_m[c]) * s_inv[c] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = e0 + t + 256 * i;
                if (e < d) {
                    const float v = __ldg(V + n * d + e);
#pragma unroll
                    for (int c = 0; c < DS_MAXC; ++c) acc[c][i] = fmaf(a[c], v, acc[c][i]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < DS_MAXC; ++c)
            if (c < C)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = e0 + t + 256 * i;
                    if (e < d) B_part[((int64_t)blockIdx.x * C + c) * d + e] = acc[c][i];
                }
    }
    __syncthreads();
    // normalise this chunk's rows of A in place (after every thread finished reading the raw scores)
    for (int64_t i = r0 * C + t; i < r1 * C; i += 256) {
        const int c = (int)(i % C);
        A[i] = expf(A[i] - s_m[c]) * s_inv[c];
    }
}

// Bm = sum of partials; logits[o] = sum_{c,e} W[o,c,e] Bm[c,e] + bias[o].  single CTA
__global__ void __launch_bounds__(256)
dsmil_head_kernel(const float* __restrict__ B_part, int chunks, int C, int d, const float* __restrict__ W,
                  const float* __restrict__ bias, float* __restrict__ Bm, float* __restrict__ logits) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < C * d; i += 256) {
        // fixed order, 8 loads in flight: a single bag waits on this serial tail
        float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int k = 0;
        for (; k + 8 <= chunks; k += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a8[u] += __ldcg(B_part + (int64_t)(k + u) * C * d + i);
        }
        for (; k < chunks; ++k) a8[0] += __ldcg(B_part + (int64_t)k * C * d + i);
        Bm[i] = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
    }
    __syncthreads();
    for (int o = warp; o < C; o += 8) {
        float acc = 0.f;
        for (int i = lane; i < C * d; i += 32) acc = fmaf(W[(int64_t)o * C * d + i], Bm[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) logits[o] = acc + (bias ? bias[o] : 0.f);
    }
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

int64_t snuffy_dsmil_workspace(int64_t N, int64_t d, int64_t C) {
    const int64_t chunks = 2 * (int64_t)sm_count();
    return (chunks * C * 2 + chunks * C * d) * 4 + 256;
}

// A[N, C] (softmax over instances), Bm[C, d], logits[C] from Q[N, dq], q_max[C, dq], V[N, d], fcc weight [C, C, d].
int snuffy_dsmil_pool_fwd(const float* Q, const float* qmax, const float* V, const float* Wfcc, const float* bfcc,
                          int64_t N, int64_t d, int64_t dq, int64_t C, float* A, float* Bm, float* logits,
                          float* stats_out, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    SNUFFY_REQUIRE(Q && qmax && V && Wfcc && A && Bm && logits && workspace, "snuffy_dsmil_pool_fwd: null pointer");
    SNUFFY_REQUIRE(C >= 1 && C <= DS_MAXC, "snuffy_dsmil_pool_fwd: supports 1..%d classes (got %lld)", DS_MAXC, (long long)C);
    SNUFFY_REQUIRE(N >= 1 && d >= 1 && dq >= 1, "snuffy_dsmil_pool_fwd: empty problem");
    SNUFFY_REQUIRE(workspace_bytes >= snuffy_dsmil_workspace(N, d, C), "snuffy_dsmil_pool_fwd: workspace too small");
    int chunks = 2 * sm_count();
    if (chunks > (N + 7) / 8) chunks = (int)((N + 7) / 8);
    float* stat_part = reinterpret_cast<float*>(workspace);
    float* B_part = stat_part + (int64_t)2 * sm_count() * C * 2;
    // the reference divides by an fp32-rounded sqrt of the query width (dsmil.py:85)
    const float scale = sqrtf((float)dq);
    dsmil_scores_kernel<<<chunks, 256, (size_t)C * dq * 4, stream>>>(Q, qmax, N, (int)dq, (int)C, scale, A, stat_part);
    dsmil_pool_kernel<<<chunks, 256, 0, stream>>>(V, N, (int)d, (int)C, chunks, stat_part, A, B_part, stats_out);
    dsmil_head_kernel<<<1, 256, 0, stream>>>(B_part, chunks, (int)C, (int)d, Wfcc, bfcc, Bm, logits);
    return check_launch("snuffy_dsmil_pool_fwd", 3);
}

}  // extern "C"
#pragma GCC visibility pop
