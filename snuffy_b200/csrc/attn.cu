// Snuffy sparse attention (snuffy.py:160-168 + head split 187-201):
//   queries and values are ALL N (LayerNormed, projected) patches, keys are the Ksel selected raw patches;
//   S_j = Q_j Kp_j^T / sqrt(dk)  [N, Ksel];  P_j = softmax over the Ksel axis (per query row);
//   O_j = P_j^T V_j  [Ksel, dk]  -- the TRANSPOSED aggregation: output rows are the keys.
//
// This file is the fp32 SIMT implementation (any head size, any Ksel).  The [h, N, Ksel] score/probability
// tensors are never materialised unless the caller asks for A: softmax is row-local, the reduction over the
// long N axis is a plain sum that is split over CTAs and folded by a deterministic second stage.
//
//   grid.x = row split, grid.y = head * v-slices, grid.z = bag * key-chunks
//   PASS 0: one key chunk covers all keys -> stats in-kernel, P and O in one pass
//   PASS 1: several key chunks -> per-chunk (max, sum) partial statistics only
//   PASS 2: several key chunks -> merge the partial statistics, P and O for this chunk
#include "common.cuh"

namespace snuffy {

constexpr int AT_DV = 64;   // value columns per CTA

struct AttnParams {
    const float* Q; int64_t ldq;     // Q(b,n,c) = Q[(b*N+n)*ldq + c]
    const float* V; int64_t ldv;
    const float* Kp;                 // [B*Ksel, d]
    int N, Ksel, h, dk, d, B;
    int nkc, nvs, rows_per_split;
    float scale;                     // fp32-rounded sqrt(dk); scores are DIVIDED by it (snuffy.py:163)
    float* stats_part;               // [B, h, nkc, N, 2]
    float* O_part;                   // [splits, B*Ksel, d]
    float* P_out;                    // [B, h, N, Ksel] or null
    float drop_p; uint64_t seed, offset;   // attention dropout (train mode), Philox keyed per element
    float* stats_out;                // optional [B, h, N, 2] final (max, 1/sum) saved for backward
};

__device__ __forceinline__ float drop_scale(const AttnParams& p, int b, int j, int n, int key) {
    const uint64_t idx = (((uint64_t)(b * p.h + j) * p.N + n) * p.Ksel + key);
    const DrawKey draw = rng_resolve(p.seed, p.offset);
    return drop_keep_scale(draw.seed, draw.offset, idx, p.drop_p);
}

template <int PASS, int KC>
__global__ void __launch_bounds__(256)
attn_simt_kernel(const AttnParams p) {
    constexpr int TN = 32;                      // query rows per tile
    constexpr int RW = TN / 8;                  // rows per warp
    constexpr int SC = KC / 32;                 // contiguous keys per lane in the S phase
    constexpr int OK = KC / 16;                 // keys per thread in the PV phase
    extern __shared__ __align__(16) float at_smem[];
    const int dk = p.dk;
    // transposed key chunk [dk][KC]; 4-key groups are XOR-swizzled by the k index so that the one-off
    // transposing store does not serialise on a single bank while the float4 reads stay conflict-free
    float* sK = at_smem;
    float* sQ = sK + (size_t)dk * KC;           // [TN][dk]
    float* sV = sQ + (size_t)TN * dk;           // [TN][AT_DV]
    float* sP = sV + (size_t)TN * AT_DV;        // [TN][KC]

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int split = blockIdx.x;
    const int j = blockIdx.y / p.nvs, vs = blockIdx.y % p.nvs;
    const int b = blockIdx.z / p.nkc, kc = blockIdx.z % p.nkc;
    const int key0 = kc * KC;
    const int nkeys = min(KC, p.Ksel - key0);
    const int dk4 = dk >> 2;

    for (int idx = t; idx < KC * dk4; idx += 256) {
        const int key = idx / dk4, k4 = (idx % dk4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key < nkeys)
            v = __ldg(reinterpret_cast<const float4*>(p.Kp + ((int64_t)b * p.Ksel + key0 + key) * p.d + j * dk + k4));
        const int ks = key ^ (((k4 >> 2) & 7) << 2);   // same swizzle for k4..k4+3
        sK[(k4 + 0) * KC + ks] = v.x; sK[(k4 + 1) * KC + ks] = v.y;
        sK[(k4 + 2) * KC + ks] = v.z; sK[(k4 + 3) * KC + ks] = v.w;
    }

    float oacc[OK][4];
#pragma unroll
    for (int a = 0; a < OK; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) oacc[a][c] = 0.f;
    const int ky = t >> 4, dx = t & 15;

    const int rbeg = split * p.rows_per_split;
    const int rend = min(p.N, rbeg + p.rows_per_split);
    for (int r0 = rbeg; r0 < rend; r0 += TN) {
        __syncthreads();
        for (int idx = t; idx < TN * dk4; idx += 256) {
            const int row = idx / dk4, k4 = (idx % dk4) * 4;
            const int n = r0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < rend) v = ld_stream(reinterpret_cast<const float4*>(p.Q + ((int64_t)b * p.N + n) * p.ldq + j * dk + k4));
            *reinterpret_cast<float4*>(sQ + row * dk + k4) = v;
        }
        if (PASS != 1) {
            for (int idx = t; idx < TN * (AT_DV / 4); idx += 256) {
                const int row = idx / (AT_DV / 4), c4 = (idx % (AT_DV / 4)) * 4;
                const int n = r0 + row, col = vs * AT_DV + c4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < rend && col < dk)
                    v = ld_stream(reinterpret_cast<const float4*>(p.V + ((int64_t)b * p.N + n) * p.ldv + j * dk + col));
                *reinterpret_cast<float4*>(sV + row * AT_DV + c4) = v;
            }
        }
        __syncthreads();

        // ---- S tile: rows warp*RW.., keys lane*SC .. lane*SC+SC-1
        float s[RW][SC];
#pragma unroll
        for (int i = 0; i < RW; ++i)
#pragma unroll
            for (int c = 0; c < SC; ++c) s[i][c] = 0.f;
        for (int k = 0; k < dk; k += 4) {
            float q[RW][4];
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                const float4 qv = *reinterpret_cast<const float4*>(sQ + (warp * RW + i) * dk + k);
                q[i][0] = qv.x; q[i][1] = qv.y; q[i][2] = qv.z; q[i][3] = qv.w;
            }
            const int sw = ((k >> 2) & 7) << 2;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float kv[SC];
                if (SC >= 4) {
#pragma unroll
                    for (int c = 0; c < SC; c += 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(sK + (k + kk) * KC + ((lane * SC + c) ^ sw));
                        kv[c] = t4.x; kv[c + 1] = t4.y; kv[c + 2] = t4.z; kv[c + 3] = t4.w;
                    }
                } else {
                    const float2 t2 = *reinterpret_cast<const float2*>(sK + (k + kk) * KC + ((lane * SC) ^ sw));
                    kv[0] = t2.x; kv[1] = t2.y;
                }
#pragma unroll
                for (int c = 0; c < SC; ++c)
#pragma unroll
                    for (int i = 0; i < RW; ++i) s[i][c] = fmaf(q[i][kk], kv[c], s[i][c]);
            }
        }
        // ---- row statistics and probabilities
#pragma unroll
        for (int i = 0; i < RW; ++i) {
            const int n = r0 + warp * RW + i;
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                s[i][c] = (lane * SC + c < nkeys) ? s[i][c] / p.scale : -INFINITY;
                m = fmaxf(m, s[i][c]);
            }
            m = warp_max(m);
            float inv = 0.f;
            if (PASS != 2) {
                float l = 0.f;
#pragma unroll
                for (int c = 0; c < SC; ++c) l += (lane * SC + c < nkeys) ? expf(s[i][c] - m) : 0.f;
                l = warp_sum(l);
                if (PASS == 1) {
                    if (lane == 0 && n < rend) {
                        float* sp = p.stats_part + ((((int64_t)b * p.h + j) * p.nkc + kc) * p.N + n) * 2;
                        sp[0] = m; sp[1] = l;
                    }
                    continue;
                }
                inv = 1.f / l;
            } else {
                float mg = -INFINITY, lg = 0.f;
                if (n < rend) {
                    const float* sp = p.stats_part + (((int64_t)b * p.h + j) * p.nkc * p.N + n) * 2;
                    for (int c2 = 0; c2 < p.nkc; ++c2) mg = fmaxf(mg, __ldcg(sp + (int64_t)c2 * p.N * 2));
                    for (int c2 = 0; c2 < p.nkc; ++c2)
                        lg += __ldcg(sp + (int64_t)c2 * p.N * 2 + 1) * expf(__ldcg(sp + (int64_t)c2 * p.N * 2) - mg);
                    m = mg; inv = 1.f / lg;
                } else {
                    m = 0.f; inv = 0.f;
                }
            }
            if (p.stats_out && vs == 0 && kc == 0 && lane == 0 && n < rend) {
                float* so = p.stats_out + (((int64_t)b * p.h + j) * p.N + n) * 2;
                so[0] = m; so[1] = inv;
            }
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                const int key = lane * SC + c;
                float pv = (key < nkeys && n < rend) ? expf(s[i][c] - m) * inv : 0.f;
                sP[(warp * RW + i) * KC + key] = pv;
            }
        }
        if (PASS == 1) continue;
        __syncthreads();
        if (p.P_out && vs == 0) {
            for (int idx = t; idx < TN * KC; idx += 256) {
                const int row = idx / KC, key = idx % KC;
                const int n = r0 + row;
                if (n < rend && key < nkeys)
                    p.P_out[(((int64_t)b * p.h + j) * p.N + n) * p.Ksel + key0 + key] = sP[idx];
            }
        }
        if (p.drop_p > 0.f) {
            // dropout acts on P AFTER it is reported (snuffy.py:164-168): rescale the tile in place
            for (int idx = t; idx < TN * KC; idx += 256) {
                const int row = idx / KC, key = idx % KC;
                const int n = r0 + row;
                if (n < rend && key < nkeys) sP[idx] *= drop_scale(p, b, j, n, key0 + key);
            }
            __syncthreads();
        }
        // ---- O[key, dv] += sum_n P[n, key] V[n, dv]
#pragma unroll 4
        for (int n = 0; n < TN; ++n) {
            const float4 vv = *reinterpret_cast<const float4*>(sV + n * AT_DV + dx * 4);
            float pk[OK];
#pragma unroll
            for (int a = 0; a < OK; a += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(sP + n * KC + ky * OK + a);
                pk[a] = t4.x; pk[a + 1] = t4.y; pk[a + 2] = t4.z; pk[a + 3] = t4.w;
            }
#pragma unroll
            for (int a = 0; a < OK; ++a) {
                oacc[a][0] = fmaf(pk[a], vv.x, oacc[a][0]);
                oacc[a][1] = fmaf(pk[a], vv.y, oacc[a][1]);
                oacc[a][2] = fmaf(pk[a], vv.z, oacc[a][2]);
                oacc[a][3] = fmaf(pk[a], vv.w, oacc[a][3]);
            }
        }
    }
    if (PASS == 1) return;
    const int col = vs * AT_DV + dx * 4;
    if (col < dk) {
#pragma unroll
        for (int a = 0; a < OK; ++a) {
            const int key = ky * OK + a;
            if (key < nkeys) {
                float* o = p.O_part + (((int64_t)split * p.B + b) * p.Ksel + key0 + key) * p.d + j * dk + col;
                *reinterpret_cast<float4*>(o) = make_float4(oacc[a][0], oacc[a][1], oacc[a][2], oacc[a][3]);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
fold_partials_kernel(const float* __restrict__ part, int splits, int64_t n4, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    // fixed order (deterministic), four independent accumulators so that the loads of different splits overlap
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)(s + u) * n4 + i);
            a[u].x += v.x; a[u].y += v.y; a[u].z += v.z; a[u].w += v.w;
        }
    }
    for (; s < splits; ++s) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)s * n4 + i);
        a[0].x += v.x; a[0].y += v.y; a[0].z += v.z; a[0].w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = make_float4((a[0].x + a[1].x) + (a[2].x + a[3].x), (a[0].y + a[1].y) + (a[2].y + a[3].y),
                                                    (a[0].z + a[1].z) + (a[2].z + a[3].z), (a[0].w + a[1].w) + (a[2].w + a[3].w));
}

void launch_fold_partials(const float* part, int splits, int64_t n4, float* out, cudaStream_t stream) {
    if (fold_wide_pays(splits, n4)) {
        fold_wide_kernel<float4><<<(unsigned)((n4 + 15) / 16), 256, 0, stream>>>(reinterpret_cast<const float4*>(part), splits, n4,
                                                                                 reinterpret_cast<float4*>(out));
        return;
    }
    const int threads = n4 >= 64 * 1024 ? 256 : 64;          // small folds: more CTAs, the loop over splits is the latency
    fold_partials_kernel<<<(unsigned)((n4 + threads - 1) / threads), threads, 0, stream>>>(part, splits, n4, out);
}

struct AttnPlan { int KC, nkc, nvs, splits, rows_per_split; size_t smem; };

static AttnPlan plan_attn(int B, int N, int Ksel, int h, int dk) {
    AttnPlan pl{};
    const int TN = 32;
    auto smem_of = [&](int KC) {
        return ((size_t)dk * KC + (size_t)TN * dk + (size_t)TN * AT_DV + (size_t)TN * KC) * sizeof(float);
    };
    // largest key chunk that still lets two CTAs share an SM; otherwise the largest that fits at all
    const size_t two_cta = 112 * 1024, one_cta = 220 * 1024;
    int KC = 64;
    if (Ksel > 64 && smem_of(128) <= two_cta) KC = 128;
    if (Ksel > 128 && smem_of(256) <= two_cta) KC = 256;
    if (smem_of(KC) > one_cta) KC = 64;
    pl.KC = KC;
    pl.smem = smem_of(KC);
    pl.nkc = (Ksel + KC - 1) / KC;
    pl.nvs = (dk + AT_DV - 1) / AT_DV;
    const int per = B * h * pl.nkc * pl.nvs;
    int splits = (2 * sm_count() + per - 1) / per;
    const int max_splits = (N + TN - 1) / TN;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rps = (N + splits - 1) / splits;
    rps = (rps + TN - 1) / TN * TN;
    pl.rows_per_split = rps;
    pl.splits = (N + rps - 1) / rps;
    return pl;
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// workspace bytes for snuffy_sparse_attn_fwd (O partials + partial statistics)
int64_t snuffy_sparse_attn_workspace(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d) {
    if (h <= 0 || d <= 0 || d % h) return -1;
    const AttnPlan pl = plan_attn((int)B, (int)N, (int)Ksel, (int)h, (int)(d / h));
    int64_t bytes = (int64_t)pl.splits * B * Ksel * d * 4;
    if (pl.nkc > 1) bytes += B * h * pl.nkc * N * 2 * 4;
    return bytes + 256;
}

// O[B*Ksel, d] = concat_j softmax_keys(Q_j Kp_j^T / sqrt(dk))^T V_j ; optional P_out[B, h, N, Ksel]
// (probabilities BEFORE dropout, as the reference returns them), optional stats_out[B, h, N, 2] = (max, 1/sum).
int snuffy_sparse_attn_fwd(const float* Q, int64_t ldq, const float* V, int64_t ldv, const float* Kp, int64_t B,
                           int64_t N, int64_t Ksel, int64_t h, int64_t d, float dropout_p, uint64_t seed,
                           uint64_t offset, float* O, float* P_out, float* stats_out, void* workspace,
                           int64_t workspace_bytes, cudaStream_t stream) {
    SNUFFY_REQUIRE(Q && V && Kp && O && workspace, "snuffy_sparse_attn_fwd: null pointer");
    SNUFFY_REQUIRE(h > 0 && d % h == 0, "snuffy_sparse_attn_fwd: d=%lld not divisible by h=%lld", (long long)d, (long long)h);
    const int dk = (int)(d / h);
    SNUFFY_REQUIRE(dk % 4 == 0 && ldq % 4 == 0 && ldv % 4 == 0 && (uintptr_t)Q % 16 == 0 && (uintptr_t)V % 16 == 0 &&
                       (uintptr_t)Kp % 16 == 0 && (uintptr_t)O % 16 == 0,
                   "snuffy_sparse_attn_fwd: head size %d must be a multiple of 4 and pointers 16-byte aligned", dk);
    SNUFFY_REQUIRE(B >= 1 && N >= 1 && Ksel >= 1, "snuffy_sparse_attn_fwd: empty problem");
    SNUFFY_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "snuffy_sparse_attn_fwd: dropout_p out of range");
    const AttnPlan pl = plan_attn((int)B, (int)N, (int)Ksel, (int)h, dk);
    SNUFFY_REQUIRE(pl.smem <= 220 * 1024, "snuffy_sparse_attn_fwd: head size %d too large", dk);
    SNUFFY_REQUIRE(workspace_bytes >= snuffy_sparse_attn_workspace(B, N, Ksel, h, d),
                   "snuffy_sparse_attn_fwd: workspace too small");
    AttnParams p{};
    p.Q = Q; p.ldq = ldq; p.V = V; p.ldv = ldv; p.Kp = Kp;
    p.N = (int)N; p.Ksel = (int)Ksel; p.h = (int)h; p.dk = dk; p.d = (int)d; p.B = (int)B;
    p.nkc = pl.nkc; p.nvs = pl.nvs; p.rows_per_split = pl.rows_per_split;
    p.scale = (float)sqrt((double)dk);
    p.O_part = reinterpret_cast<float*>(workspace);
    p.stats_part = p.O_part + (int64_t)pl.splits * B * Ksel * d;
    p.P_out = P_out; p.stats_out = stats_out;
    p.drop_p = dropout_p; p.seed = seed; p.offset = offset;
    dim3 grid((unsigned)pl.splits, (unsigned)(h * pl.nvs), (unsigned)(B * pl.nkc));
#define ATTN_LAUNCH(PASS, KC)                                                                                   \
    do {                                                                                                        \
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_simt_kernel<PASS, KC>), \
                                         (int)pl.smem));                                                        \
        attn_simt_kernel<PASS, KC><<<grid, 256, pl.smem, stream>>>(p);                                          \
    } while (0)
#define ATTN_DISPATCH(PASS)                       \
    do {                                          \
        if (pl.KC == 256) ATTN_LAUNCH(PASS, 256); \
        else if (pl.KC == 128) ATTN_LAUNCH(PASS, 128); \
        else ATTN_LAUNCH(PASS, 64);               \
    } while (0)
    if (pl.nkc == 1) {
        ATTN_DISPATCH(0);
    } else {
        ATTN_DISPATCH(1);
        ATTN_DISPATCH(2);
    }
#undef ATTN_DISPATCH
#undef ATTN_LAUNCH
    const int64_t n4 = B * Ksel * d / 4;
    fold_partials_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(p.O_part, pl.splits, n4, O);
    return check_launch("snuffy_sparse_attn_fwd", pl.nkc == 1 ? 2 : 3);
}

}  // extern "C"
#pragma GCC visibility pop
