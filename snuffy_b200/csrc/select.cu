// Instance scoring + patch selection:
//   scores   c = x W^T + b                       snuffy.py:39-41            (fp32 GEMV, HBM-bound)
//   top-K    first K of the descending order      snuffy.py:128-129          (exact, ties -> lower index)
//   random-K uniform sample of the complement     snuffy.py:136-143          (device Philox, no host trip)
//   unique   ascending distinct indices           snuffy_multiclass.py:140
//   gather   raw rows x[S]                        snuffy.py:131,145-147
//
// Selection never sorts the N scores: it radix-selects the K-th largest of the
// 64-bit composites (order-preserving score bits << 32 | ~index), which are all
// distinct, so the tie rule "lower index first" is part of the key and exactly K
// elements are >= the threshold.  The K winners are then bitonic-sorted so the
// output order equals the reference's sort order on tie-free inputs.
#include "common.cuh"

namespace snuffy {

// ------------------------------------------------------------------ scores (a1)
template <int CC>
__global__ void __launch_bounds__(256)
scores_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
              float* __restrict__ c, int64_t rows, int d, int C, int c0, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * (int64_t)d;
    float acc[CC];
#pragma unroll
    for (int j = 0; j < CC; ++j) acc[j] = 0.f;
    if (vec_ok == 2) {
        // d % 8 == 0: each lane owns 8 consecutive elements per 256-element slab, accumulated in index order -- the SAME
        // partition and order as ln_rows_kernel's fused scorer (norm.cu), so both produce bit-identical scores
        for (int e = lane * 8; e < d; e += 256) {
            const float4 x0 = ld_stream(reinterpret_cast<const float4*>(xr + e));
            const float4 x1 = ld_stream(reinterpret_cast<const float4*>(xr + e + 4));
#pragma unroll
            for (int j = 0; j < CC; ++j) {
                if (c0 + j < C) {
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(c0 + j) * d + e));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(c0 + j) * d + e + 4));
                    acc[j] = fmaf(x0.x, w0.x, acc[j]); acc[j] = fmaf(x0.y, w0.y, acc[j]);
                    acc[j] = fmaf(x0.z, w0.z, acc[j]); acc[j] = fmaf(x0.w, w0.w, acc[j]);
                    acc[j] = fmaf(x1.x, w1.x, acc[j]); acc[j] = fmaf(x1.y, w1.y, acc[j]);
                    acc[j] = fmaf(x1.z, w1.z, acc[j]); acc[j] = fmaf(x1.w, w1.w, acc[j]);
                }
            }
        }
    } else if (vec_ok) {
        for (int e = lane * 4; e < d; e += 128) {
            const float4 xv = ld_stream(reinterpret_cast<const float4*>(xr + e));
#pragma unroll
            for (int j = 0; j < CC; ++j) {
                if (c0 + j < C) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(c0 + j) * d + e));
                    acc[j] = fmaf(xv.x, wv.x, acc[j]);
                    acc[j] = fmaf(xv.y, wv.y, acc[j]);
                    acc[j] = fmaf(xv.z, wv.z, acc[j]);
                    acc[j] = fmaf(xv.w, wv.w, acc[j]);
                }
            }
        }
    } else {
        for (int e = lane; e < d; e += 32) {
            const float xv = xr[e];
#pragma unroll
            for (int j = 0; j < CC; ++j)
                if (c0 + j < C) acc[j] = fmaf(xv, __ldg(W + (int64_t)(c0 + j) * d + e), acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < CC; ++j) {
        const float s = warp_sum(acc[j]);
        if (lane == 0 && c0 + j < C) c[row * C + c0 + j] = s + (bias ? bias[c0 + j] : 0.f);
    }
}

// ------------------------------------------------------------------ selection
constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX_K = 4096;
constexpr int SEL_KEY_CACHE = 45056;   // uint32 keys cached in shared memory (176 KB)

__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint32_t philox_word(uint64_t seed, uint64_t offset, uint32_t bag, uint32_t i) {
    // Philox4x32-10, counter = (i, bag, offset_lo, offset_hi), key = seed
    uint32_t c0 = i, c1 = bag, c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// MODE 0: key = score of (bag, class).  MODE 1: key = Philox word, 0 for flagged (taken) rows.
// One CTA selects the K best keys of one (bag, class) column: shared by the stand-alone kernel and the fused selection kernel.
template <int MODE>
__device__ __forceinline__ void select_cta(const float* __restrict__ scores, int N, int C, int K, int Kpad, int cache_keys,
                                           uint8_t* __restrict__ flags, uint64_t seed, uint64_t offset,
                                           int64_t* __restrict__ o, const int64_t* __restrict__ cu_seqlens, int col, int bag,
                                           unsigned char* sel_smem) {
    if (MODE == 1) { const DrawKey key_ = rng_resolve(seed, offset); seed = key_.seed; offset = key_.offset; }
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(sel_smem);
    uint32_t* keys = reinterpret_cast<uint32_t*>(sel_smem + (size_t)Kpad * 8);
    __shared__ uint32_t hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ int s_need, s_bucket, s_count;

    const int tid = threadIdx.x, lane = tid & 31;
    // packed variable-length bags: rows [cu[bag], cu[bag+1]) of one [T, C] score matrix; the indices written are GLOBAL
    // rows of the packed tensor, so every row-wise kernel downstream runs on it as one "bag" of T rows
    int64_t base = (int64_t)bag * N;
    if (cu_seqlens) { base = cu_seqlens[bag]; N = (int)(cu_seqlens[bag + 1] - base); }
    const int64_t idx_add = cu_seqlens ? base : 0;
    const float* sc = scores ? scores + base * C + col : nullptr;
    uint8_t* fl = flags ? flags + base : nullptr;

    auto raw_key = [&](int i) -> uint32_t {
        if (MODE == 0) return float_key(sc[(int64_t)i * C]);
        if (fl[i]) return 0u;
        return (philox_word(seed, offset, (uint32_t)bag, (uint32_t)i) >> 1) | 0x80000000u;
    };
    if (cache_keys) {
        for (int i = tid; i < N; i += SEL_THREADS) keys[i] = raw_key(i);
    }
    if (tid == 0) s_count = 0;
    __syncthreads();
    auto comp_of = [&](int i) -> uint64_t {
        const uint32_t k = cache_keys ? keys[i] : raw_key(i);
        return ((uint64_t)k << 32) | (uint32_t)(~(uint32_t)i);
    };

    uint64_t prefix = 0;
    int need = K;
    const int n_round = (N + 31) & ~31;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n_round; i += SEL_THREADS) {
            bool m = i < N;
            uint64_t comp = 0;
            if (m) {
                comp = comp_of(i);
                if (pass > 0) m = (comp >> (shift + 8)) == (prefix >> (shift + 8));
            }
            const uint32_t digit = (uint32_t)(comp >> shift) & 255u;
            const unsigned active = __ballot_sync(0xffffffffu, m);
            if (m) {
                const unsigned peers = __match_any_sync(active, digit);
                if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; sum += loc[j]; }
            uint32_t incl = sum;   // becomes sum over lanes >= lane
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
                if (lane + o < 32) incl += t;
            }
            const uint32_t above = incl - sum;
            if (above < (uint32_t)need && incl >= (uint32_t)need) {
                uint32_t cum = above;
                bool found = false;
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    if (!found) {
                        if (cum + loc[j] >= (uint32_t)need) {
                            s_prefix = prefix | ((uint64_t)(lane * 8 + j) << shift);
                            s_need = need - (int)cum;
                            s_bucket = (int)loc[j];
                            found = true;
                        } else {
                            cum += loc[j];
                        }
                    }
                }
            }
        }
        __syncthreads();
        prefix = s_prefix;
        need = s_need;
        const int bucket = s_bucket;
        __syncthreads();
        if (bucket == need) break;   // the whole bucket is selected: lower digits cannot matter
    }

    // compaction of the K winners (arbitrary order), then bitonic sort descending
    for (int i = tid; i < n_round; i += SEL_THREADS) {
        uint64_t comp = 0;
        bool m = false;
        if (i < N) { comp = comp_of(i); m = comp >= prefix; }
        const unsigned sel = __ballot_sync(0xffffffffu, m);
        if (sel) {
            int base = 0;
            if (lane == __ffs(sel) - 1) base = atomicAdd(&s_count, __popc(sel));
            base = __shfl_sync(0xffffffffu, base, __ffs(sel) - 1);
            if (m) {
                const int pos = base + __popc(sel & ((1u << lane) - 1u));
                if (pos < Kpad) sortbuf[pos] = comp;
            }
        }
    }
    __syncthreads();
    for (int i = K + tid; i < Kpad; i += SEL_THREADS) sortbuf[i] = 0;
    __syncthreads();
    for (int k = 2; k <= Kpad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Kpad; i += SEL_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = sortbuf[i], b = sortbuf[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { sortbuf[i] = b; sortbuf[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < K; r += SEL_THREADS) {
        const uint32_t idx = ~(uint32_t)sortbuf[r];
        o[r] = (int64_t)idx + idx_add;
        if (MODE == 0 && fl) fl[idx] = 1;
    }
}

template <int MODE>
__global__ void __launch_bounds__(SEL_THREADS, 1)
select_kernel(const float* __restrict__ scores, int N, int C, int K, int Kpad, int cache_keys,
              uint8_t* __restrict__ flags, uint64_t seed, uint64_t offset, int64_t* __restrict__ out,
              const int64_t* __restrict__ cu_seqlens) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    select_cta<MODE>(scores, N, C, K, Kpad, cache_keys, flags, seed, offset,
                     out + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * K, cu_seqlens, (int)blockIdx.x, (int)blockIdx.y, sel_smem);
}

// The selection of the binary path in ONE launch per bag batch (snuffy.py:128-147 + 131, 145-147, 152-155): one CTA per bag
// clears the bag's taken-flags and row map, selects the Ktop highest-scoring rows, draws Krand of the remaining rows (same
// Philox stream as the stand-alone kernel: identical indices), writes S = T ++ R, the row map (row -> slot) and gathers the
// selected rows of x into xs [B, Ksel, d].
__global__ void __launch_bounds__(SEL_THREADS, 1)
select_gather_kernel(const float* __restrict__ scores, const float* __restrict__ x, int N, int d, int Ktop, int Krand,
                     int Kpad_top, int Kpad_rand, int cache_keys, uint8_t* __restrict__ flags, uint64_t seed, uint64_t offset,
                     int64_t* __restrict__ sel, int32_t* __restrict__ row_map, float* __restrict__ xs, int vec_ok) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    const int bag = blockIdx.x, tid = threadIdx.x, Ksel = Ktop + Krand;
    uint8_t* fl = flags + (int64_t)bag * N;
    int32_t* rm = row_map + (int64_t)bag * N;
    for (int i = tid; i < N; i += SEL_THREADS) { fl[i] = 0; rm[i] = -1; }
    __syncthreads();
    int64_t* s_out = sel + (int64_t)bag * Ksel;
    select_cta<0>(scores, N, 1, Ktop, Kpad_top, cache_keys, flags, 0, 0, s_out, nullptr, 0, bag, sel_smem);
    __syncthreads();
    if (Krand > 0) {
        select_cta<1>(nullptr, N, 1, Krand, Kpad_rand, cache_keys, flags, seed, offset, s_out + Ktop, nullptr, 0, bag, sel_smem);
        __syncthreads();
    }
    for (int r = tid; r < Ksel; r += SEL_THREADS) rm[s_out[r]] = (int32_t)(bag * Ksel + r);
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < Ksel; r += SEL_THREADS / 32) {
        const float* src = x + ((int64_t)bag * N + s_out[r]) * d;
        float* dst = xs + ((int64_t)bag * Ksel + r) * d;
        if (vec_ok) {
            for (int e = lane * 4; e < d; e += 128)
                *reinterpret_cast<float4*>(dst + e) = __ldg(reinterpret_cast<const float4*>(src + e));
        } else {
            for (int e = lane; e < d; e += 32) dst[e] = src[e];
        }
    }
}

// ascending distinct indices of the flagged rows of each bag (torch.unique order)
__global__ void __launch_bounds__(1024, 1)
compact_flags_kernel(const uint8_t* __restrict__ flags, int N, int cap, int64_t* __restrict__ out,
                     int32_t* __restrict__ counts) {
    __shared__ int warp_cnt[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bag = blockIdx.x;
    const uint8_t* fl = flags + (int64_t)bag * N;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < N; start += 1024) {
        const int i = start + tid;
        const bool f = i < N && fl[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const int cnt = warp_cnt[w]; if (w < warp) woff += cnt; total += cnt; }
        const int pos = s_base + woff + __popc(bal & ((1u << lane) - 1u));
        if (f && pos < cap) out[(int64_t)bag * cap + pos] = i;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) counts[bag] = s_base;
}

// ------------------------------------------------------------------ gather / row map
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int64_t N, int64_t K, int d,
                   int64_t total, float* __restrict__ out, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= total) return;
    const int64_t bag = r / K;
    const float* src = x + (bag * N + idx[r]) * (int64_t)d;
    float* dst = out + r * (int64_t)d;
    if (vec_ok) {
        for (int e = lane * 4; e < d; e += 128)
            *reinterpret_cast<float4*>(dst + e) = __ldg(reinterpret_cast<const float4*>(src + e));
    } else {
        for (int e = lane; e < d; e += 32) dst[e] = src[e];
    }
}

__global__ void row_map_kernel(const int64_t* __restrict__ idx, int64_t N, int64_t K, int64_t total,
                               int32_t* __restrict__ row_map) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total) return;
    row_map[(r / K) * N + idx[r]] = (int32_t)r;
}

static inline int vec_ok4(const void* p, int64_t d) { return (d % 4 == 0) && (((uintptr_t)p) % 16 == 0); }

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

int snuffy_scores_fwd(const float* x, const float* W, const float* bias, float* c, int64_t rows, int64_t d,
                      int64_t C, cudaStream_t stream) {
    SNUFFY_REQUIRE(rows >= 0 && d > 0 && C > 0, "snuffy_scores_fwd: bad shape rows=%lld d=%lld C=%lld",
                   (long long)rows, (long long)d, (long long)C);
    if (rows == 0) return 0;
    const int vec = (vec_ok4(x, d) && vec_ok4(W, d)) ? (d % 8 == 0 ? 2 : 1) : 0;
    const int warps = 8;
    const unsigned grid = (unsigned)((rows + warps - 1) / warps);
    for (int c0 = 0; c0 < C; c0 += 4) {
        if (C - c0 == 1)
            scores_kernel<1><<<grid, warps * 32, 0, stream>>>(x, W, bias, c, rows, (int)d, (int)C, c0, vec);
        else if (C - c0 == 2)
            scores_kernel<2><<<grid, warps * 32, 0, stream>>>(x, W, bias, c, rows, (int)d, (int)C, c0, vec);
        else
            scores_kernel<4><<<grid, warps * 32, 0, stream>>>(x, W, bias, c, rows, (int)d, (int)C, c0, vec);
    }
    return check_launch("snuffy_scores_fwd", (int)((C + 3) / 4));
}

static int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

static int launch_select(int mode, const float* scores, int64_t B, int64_t N, int64_t C, int64_t K,
                         uint8_t* flags, uint64_t seed, uint64_t offset, int64_t* out, cudaStream_t stream,
                         const int64_t* cu_seqlens = nullptr) {
    SNUFFY_REQUIRE(B >= 1 && N >= 1 && C >= 1, "select: bad shape B=%lld N=%lld C=%lld", (long long)B,
                   (long long)N, (long long)C);
    SNUFFY_REQUIRE(K >= 0 && K <= N, "select: K=%lld must be in [0, N=%lld]", (long long)K, (long long)N);
    SNUFFY_REQUIRE(K <= SEL_MAX_K, "select: K=%lld exceeds the supported maximum %d", (long long)K, SEL_MAX_K);
    SNUFFY_REQUIRE(N < (1ll << 31), "select: N too large");
    if (K == 0) return 0;
    const int Kpad = next_pow2((int)K);
    const int cache = N <= SEL_KEY_CACHE ? 1 : 0;
    const size_t smem = (size_t)Kpad * 8 + (cache ? (size_t)N * 4 : 0);
    dim3 grid((unsigned)C, (unsigned)B);
    if (mode == 0) {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&select_kernel<0>), 220 * 1024));
        select_kernel<0><<<grid, SEL_THREADS, smem, stream>>>(scores, (int)N, (int)C, (int)K, Kpad, cache, flags,
                                                              seed, offset, out, cu_seqlens);
    } else {
        SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&select_kernel<1>), 220 * 1024));
        select_kernel<1><<<grid, SEL_THREADS, smem, stream>>>(nullptr, (int)N, 1, (int)K, Kpad, cache, flags, seed,
                                                              offset, out, cu_seqlens);
    }
    return check_launch("snuffy_select");
}

// idx_out[B, C, K] int64: per (bag, class) the K highest-scoring rows in descending score order,
// ties broken towards the lower index.  flags[B, N] (optional, uint8, caller-zeroed) gets 1 at every winner.
int snuffy_select_topk(const float* scores, int64_t B, int64_t N, int64_t C, int64_t K, int64_t* idx_out,
                       uint8_t* flags, cudaStream_t stream) {
    SNUFFY_REQUIRE(scores && idx_out, "snuffy_select_topk: null pointer");
    return launch_select(0, scores, B, N, C, K, flags, 0, 0, idx_out, stream);
}

// idx_out[B, K] int64: K distinct rows drawn uniformly without replacement from the rows whose flag is 0.
// Counter-based (Philox4x32-10 keyed by seed, counter = (row, bag, offset)): reproducible, no host round trip.
int snuffy_select_random(const uint8_t* flags, int64_t B, int64_t N, int64_t K, uint64_t seed, uint64_t offset,
                         int64_t* idx_out, cudaStream_t stream) {
    SNUFFY_REQUIRE(flags && idx_out, "snuffy_select_random: null pointer");
    return launch_select(1, nullptr, B, N, 1, K, const_cast<uint8_t*>(flags), seed, offset, idx_out, stream);
}

// Fused selection of the binary path: sel[B, Ktop + Krand] int64 (top-k in descending score order, then the random rows),
// flags[B, N] uint8 and row_map[B * N] int32 (both written from scratch), xs[B, Ktop + Krand, d] = the selected rows of x.
int snuffy_select_gather(const float* scores, const float* x, int64_t B, int64_t N, int64_t d, int64_t k_top, int64_t k_rand,
                         uint64_t seed, uint64_t offset, int64_t* sel, uint8_t* flags, int32_t* row_map, float* xs,
                         cudaStream_t stream) {
    SNUFFY_REQUIRE(scores && x && sel && flags && row_map && xs, "snuffy_select_gather: null pointer");
    SNUFFY_REQUIRE(B >= 1 && N >= 1 && d >= 1 && N < (1ll << 31) && B <= 65535, "snuffy_select_gather: bad shape");
    SNUFFY_REQUIRE(k_top >= 1 && k_rand >= 0 && k_top + k_rand <= N && k_top <= SEL_MAX_K && k_rand <= SEL_MAX_K &&
                       B * (k_top + k_rand) < (1ll << 31), "snuffy_select_gather: bad selection sizes");
    const int kp_top = next_pow2((int)k_top), kp_rand = k_rand > 0 ? next_pow2((int)k_rand) : 2;
    const int cache = N <= SEL_KEY_CACHE ? 1 : 0;
    const size_t smem = (size_t)(kp_top > kp_rand ? kp_top : kp_rand) * 8 + (cache ? (size_t)N * 4 : 0);
    SNUFFY_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&select_gather_kernel), 220 * 1024));
    select_gather_kernel<<<(unsigned)B, SEL_THREADS, smem, stream>>>(scores, x, (int)N, (int)d, (int)k_top, (int)k_rand, kp_top,
                                                                   kp_rand, cache, flags, seed, offset, sel, row_map, xs,
                                                                   vec_ok4(x, d) && vec_ok4(xs, d));
    return check_launch("snuffy_select_gather");
}

// Packed variable-length forms (BASELINE configs[3]): bag b owns rows [cu_seqlens[b], cu_seqlens[b+1]) of the packed
// [T, C] scores / [T] flags (cu_seqlens: B+1 int64 on the device; max_n = longest bag; every bag needs >= K rows, which the
// caller checks on the host).  The indices written are GLOBAL rows of the packed tensor.
int snuffy_select_topk_varlen(const float* scores, const int64_t* cu_seqlens, int64_t B, int64_t max_n, int64_t C, int64_t K,
                              int64_t* idx_out, uint8_t* flags, cudaStream_t stream) {
    SNUFFY_REQUIRE(scores && idx_out && cu_seqlens, "snuffy_select_topk_varlen: null pointer");
    return launch_select(0, scores, B, max_n, C, K, flags, 0, 0, idx_out, stream, cu_seqlens);
}
int snuffy_select_random_varlen(const uint8_t* flags, const int64_t* cu_seqlens, int64_t B, int64_t max_n, int64_t K,
                                uint64_t seed, uint64_t offset, int64_t* idx_out, cudaStream_t stream) {
    SNUFFY_REQUIRE(flags && idx_out && cu_seqlens, "snuffy_select_random_varlen: null pointer");
    return launch_select(1, nullptr, B, max_n, 1, K, const_cast<uint8_t*>(flags), seed, offset, idx_out, stream, cu_seqlens);
}

// out[B, cap] = ascending indices of flagged rows (first cap of them), counts[B] = number flagged.
int snuffy_compact_flags(const uint8_t* flags, int64_t B, int64_t N, int64_t cap, int64_t* out, int32_t* counts,
                         cudaStream_t stream) {
    SNUFFY_REQUIRE(flags && out && counts && B >= 1 && N >= 1 && cap >= 1, "snuffy_compact_flags: bad arguments");
    compact_flags_kernel<<<(unsigned)B, 1024, 0, stream>>>(flags, (int)N, (int)cap, out, counts);
    return check_launch("snuffy_compact_flags");
}

// out[B, K, d] = x[b, idx[b, k], :]
int snuffy_gather_rows(const float* x, const int64_t* idx, int64_t B, int64_t N, int64_t K, int64_t d, float* out,
                       cudaStream_t stream) {
    SNUFFY_REQUIRE(x && idx && out, "snuffy_gather_rows: null pointer");
    const int64_t total = B * K;
    if (total == 0) return 0;
    const int vec = vec_ok4(x, d) && vec_ok4(out, d);
    gather_rows_kernel<<<(unsigned)((total + 7) / 8), 256, 0, stream>>>(x, idx, N, K, (int)d, total, out, vec);
    return check_launch("snuffy_gather_rows");
}

// row_map[B*N] int32: -1, or the slot (b*K + k) of the selected row -- lets later kernels read
// "x with the selected rows replaced" without cloning x (snuffy.py:152-155).
int snuffy_build_row_map(const int64_t* idx, int64_t B, int64_t N, int64_t K, int32_t* row_map, cudaStream_t stream) {
    SNUFFY_REQUIRE(idx && row_map, "snuffy_build_row_map: null pointer");
    SNUFFY_CUDA(cudaMemsetAsync(row_map, 0xFF, (size_t)(B * N) * sizeof(int32_t), stream));
    const int64_t total = B * K;
    if (total == 0) return 0;
    row_map_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(idx, N, K, total, row_map);
    return check_launch("snuffy_build_row_map", 2);
}

}  // extern "C"
#pragma GCC visibility pop
