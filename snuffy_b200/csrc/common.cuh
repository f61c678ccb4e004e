// Shared device/host helpers for libsnuffy_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace snuffy {

// ---------------------------------------------------------------- error state
// No exceptions cross the C ABI: every entry point returns 0 / non-zero and the
// message is kept per host thread (SURVEY.md §8b "Errors").
void set_error(const char* fmt, ...);
int  check_launch(const char* what, int launches = 1);   // cudaGetLastError -> status; counts kernel launches
int  sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size): the call costs microseconds on every launch
// otherwise, and the training step is bound by host-side launch overhead.
cudaError_t ensure_dynamic_smem(const void* func, int bytes);

#define SNUFFY_REQUIRE(cond, ...)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            ::snuffy::set_error(__VA_ARGS__);             \
            return 1;                                     \
        }                                                 \
    } while (0)

#define SNUFFY_CUDA(call)                                                             \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            ::snuffy::set_error("%s failed: %s", #call, cudaGetErrorString(e__));     \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------- activations
// ids shared with the host side (snuffy_b200/_lib.py ACT_IDS); snuffy.py:216-221
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_LEAKY = 3, ACT_SELU = 4, ACT_TANH = 5 };

__device__ __forceinline__ float act_apply(int act, float t) {
    switch (act) {
        case ACT_RELU:  return fmaxf(t, 0.f);
        case ACT_GELU:  return 0.5f * t * (1.f + erff(t * 0.70710678118654752440f));
        case ACT_LEAKY: return t >= 0.f ? t : 0.01f * t;
        case ACT_SELU: {
            const float alpha = 1.6732632423543772848170429916717f;
            const float scale = 1.0507009873554804934193349852946f;
            return scale * (t > 0.f ? t : alpha * expm1f(t));
        }
        case ACT_TANH:  return tanhf(t);
        default:        return t;
    }
}

// derivative of the activation w.r.t. its pre-activation input t
__device__ __forceinline__ float act_grad(int act, float t) {
    switch (act) {
        case ACT_RELU:  return t > 0.f ? 1.f : 0.f;
        case ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(t * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * expf(-0.5f * t * t);
            return cdf + t * pdf;
        }
        case ACT_LEAKY: return t >= 0.f ? 1.f : 0.01f;
        case ACT_SELU: {
            const float alpha = 1.6732632423543772848170429916717f;
            const float scale = 1.0507009873554804934193349852946f;
            return t > 0.f ? scale : scale * alpha * expf(t);
        }
        case ACT_TANH: { const float y = tanhf(t); return 1.f - y * y; }
        default:        return 1.f;
    }
}

// ---------------------------------------------------------------- dropout
// Counter-based keep mask: element `idx` of the tensor drawn with (seed, offset) is kept iff u(idx) >= p and
// is then scaled by 1/(1-p) (nn.Dropout semantics).  Forward and backward regenerate the same mask.
// Indirect draws, for CUDA-graph replays (the kernel arguments are frozen at capture): when bit 63 of `seed` is set, `offset` is
// the address of a device uint64 step counter, and the draw actually used is
//   (seed & 0xFFFFFFFF,  *counter + ((seed >> 32) & 0x7FFFFFFF)).
// snuffy_rng_advance bumps the counter once per replay; forward and backward of one replay see the same value.
struct DrawKey { uint64_t seed, offset; };
__device__ __forceinline__ DrawKey rng_resolve(uint64_t seed, uint64_t offset) {
    if (seed >> 63) {
        offset = *reinterpret_cast<const uint64_t*>(offset) + ((seed >> 32) & 0x7FFFFFFFull);
        seed &= 0xFFFFFFFFull;
    }
    return DrawKey{seed, offset};
}
// Dropout draw of element `idx` of a launch keyed by (seed, offset): ONE 32-bit hash per PAIR of consecutive elements, its low
// half deciding the even element and its high half the odd one: keep iff half >= ceil(p * 2^16) (the realised drop
// probability is within 1.6e-5 of p; kept values are scaled by 1 / (1 - p) like torch's dropout).  Producers that walk
// consecutive elements (GEMM epilogues, the attention softmax) hash once per pair; everything else calls per element.
__device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)ceilf(p * 65536.0f); }
__device__ __forceinline__ uint32_t drop_hash(uint64_t seed, uint64_t offset, uint64_t pair) {
    uint32_t x = (uint32_t)pair ^ (uint32_t)seed, y = (uint32_t)(pair >> 32) ^ (uint32_t)(seed >> 32) ^ (uint32_t)offset;
    x *= 0x85EBCA6Bu; x ^= x >> 13; x += y * 0x9E3779B9u + (uint32_t)(offset >> 32);
    x *= 0xC2B2AE35u; x ^= x >> 16; x *= 0x27D4EB2Fu; x ^= x >> 15; x *= 0x165667B1u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float drop_keep_scale(uint64_t seed, uint64_t offset, uint64_t idx, float p) {
    const uint32_t x = drop_hash(seed, offset, idx >> 1);
    const uint32_t half = (idx & 1) ? (x >> 16) : (x & 0xFFFFu);
    return half >= drop_threshold(p) ? 1.f / (1.f - p) : 0.f;
}

// The same draw for element indices below 2^32 with everything that does not depend on the index hoisted (bit-identical to
// drop_keep_scale: the high word of the pair index is 0 there).
struct DropFast { uint32_t s0, k, thr; float scale; };
__device__ __forceinline__ DropFast drop_fast_setup(uint64_t seed, uint64_t offset, float p) {
    DropFast c;
    c.s0 = (uint32_t)seed;
    c.k = ((uint32_t)(seed >> 32) ^ (uint32_t)offset) * 0x9E3779B9u + (uint32_t)(offset >> 32);
    c.thr = drop_threshold(p);
    c.scale = 1.f / (1.f - p);
    return c;
}
__device__ __forceinline__ uint32_t drop_fast_hash(const DropFast& c, uint32_t pair) {
    uint32_t x = pair ^ c.s0;
    x *= 0x85EBCA6Bu; x ^= x >> 13; x += c.k;
    x *= 0xC2B2AE35u; x ^= x >> 16; x *= 0x27D4EB2Fu; x ^= x >> 15; x *= 0x165667B1u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool drop_fast_keep(const DropFast& c, uint32_t idx) {
    const uint32_t x = drop_fast_hash(c, idx >> 1);
    return ((idx & 1u) ? (x >> 16) : (x & 0xFFFFu)) >= c.thr;
}

// ---------------------------------------------------------------- folding many small partials
// out[i] = sum_s part[s * n + i] for FEW outputs and MANY partials (LayerNorm parameter gradients: 148 x 2d, column sums:
// ~300 x d).  One thread per output would walk `splits` dependent-latency loads; here 16 threads share an output (each takes
// every 16th partial, four accumulators) and the 16 sums are added in a fixed order: deterministic, ~10 loads deep.
__device__ __forceinline__ void fold_acc(float& a, float v) { a += v; }
__device__ __forceinline__ void fold_acc(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
template <typename T>
__global__ void __launch_bounds__(256)
fold_wide_kernel(const T* __restrict__ part, int splits, int64_t n, T* __restrict__ out) {
    __shared__ T red[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i = (int64_t)blockIdx.x * 16 + tx;
    T a0 = T(), a1 = T(), a2 = T(), a3 = T();
    if (i < n) {
        int s = ty;
        for (; s + 48 < splits; s += 64) {
            fold_acc(a0, __ldcg(part + (int64_t)s * n + i));
            fold_acc(a1, __ldcg(part + (int64_t)(s + 16) * n + i));
            fold_acc(a2, __ldcg(part + (int64_t)(s + 32) * n + i));
            fold_acc(a3, __ldcg(part + (int64_t)(s + 48) * n + i));
        }
        for (; s < splits; s += 16) fold_acc(a0, __ldcg(part + (int64_t)s * n + i));
    }
    fold_acc(a0, a1); fold_acc(a2, a3); fold_acc(a0, a2);
    red[ty][tx] = a0;
    __syncthreads();
    if (ty == 0 && i < n) {
        T r = red[0][tx];
#pragma unroll
        for (int k = 1; k < 16; ++k) fold_acc(r, red[k][tx]);
        out[i] = r;
    }
}
inline bool fold_wide_pays(int splits, int64_t n) { return splits >= 32 && n <= 16384; }

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load that does not pollute L1 (bag rows are read once per kernel)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ---------------------------------------------------------------- split-bf16 operand planes
// A "plane set" stores an fp32 matrix [rows, K] as two bf16 matrices hi + lo
// (hi = bf16(v), lo = bf16(v - hi)) in the tiled order tcgen05.mma reads with a
// SWIZZLE_NONE / K-major shared-memory descriptor, so that one pipeline stage is a
// single contiguous chunk in HBM and moves with ONE cp.async.bulk (no tensor map):
//
//   chunk(rt, kb)  = rows [rt*RC, rt*RC+RC) x k [kb*32, kb*32+32)      (RC*64 bytes)
//   inside a chunk : [kg 0..3][row 0..RC-1][8 bf16]   -> core matrix = 8 rows x 16 B
//   plane          : chunks ordered [rt][kb]; lo plane follows hi plane at plane_stride
//
// RC (rows per chunk) is 128 for A operands and BLOCK_N (128/256) for B operands.
constexpr int PLANE_KB = 32;   // k elements per chunk

__host__ __device__ __forceinline__ int64_t plane_kblocks(int64_t K) { return (K + PLANE_KB - 1) / PLANE_KB; }
__host__ __device__ __forceinline__ int64_t plane_rtiles(int64_t rows, int rc) { return (rows + rc - 1) / rc; }
// elements (bf16) in ONE plane
__host__ __device__ __forceinline__ int64_t plane_elems(int64_t rows, int64_t K, int rc) {
    return plane_rtiles(rows, rc) * plane_kblocks(K) * (int64_t)rc * PLANE_KB;
}
// element offset of the 8-element (16 B) unit holding (row, k..k+7), k % 8 == 0
__host__ __device__ __forceinline__ int64_t plane_unit_offset(int64_t row, int64_t k, int64_t K, int rc) {
    const int64_t rt = row / rc, rr = row % rc;
    const int64_t kb = k / PLANE_KB, kg = (k % PLANE_KB) / 8;
    return ((rt * plane_kblocks(K) + kb) * 4 + kg) * (int64_t)rc * 8 + rr * 8;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

struct alignas(16) bf16x8 { __nv_bfloat16 v[8]; };

}  // namespace snuffy
