// Patch-level outputs on the device (caller side of the path):
//   patch_probs        sigmoid of the instance scores (train.py:913-916: `torch.sigmoid(ins_prediction.view(-1, 1))`) written
//                      at a row offset of one epoch-wide buffer, so validation needs ONE device->host copy per epoch
//                      instead of `attentions.cpu().numpy()` per bag (train.py:271, 345, 354)
//   froc_detections    per slide: (probability, x*tile + half, y*tile + half) of every patch whose probability is strictly
//                      above the threshold, in patch order (train.py:342-345 list + mp_thresholding train.py:138-141)
// HBM-bound byte work: one streaming pass each; the compaction is a stable ballot/prefix scan, one CTA per slide.
#include "common.cuh"

namespace snuffy {

__device__ __forceinline__ float sigmoid_precise(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void __launch_bounds__(256)
patch_probs_kernel(const float* __restrict__ scores, int64_t n, float* __restrict__ probs) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(scores) | reinterpret_cast<uintptr_t>(probs)) & 15) == 0;
    if (vec) {
        const int64_t n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(scores);
        float4* p4 = reinterpret_cast<float4*>(probs);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 v = __ldg(s4 + i);
            v.x = sigmoid_precise(v.x); v.y = sigmoid_precise(v.y); v.z = sigmoid_precise(v.z); v.w = sigmoid_precise(v.w);
            p4[i] = v;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) probs[i] = sigmoid_precise(scores[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) probs[i] = sigmoid_precise(scores[i]);
    }
}

// One CTA per slide b: rows [start, end) of the packed probability / position arrays.  Kept rows are written at
// det_*[start + rank] (rank = number of kept rows before it in the slide: stable), count[b] = kept rows.
__global__ void __launch_bounds__(256)
froc_detections_kernel(const float* __restrict__ probs, int64_t prob_stride, const int32_t* __restrict__ pos,
                       const int32_t* __restrict__ cu_seqlens, int64_t total, float threshold, int tile, int half,
                       float* __restrict__ det_prob, int32_t* __restrict__ det_xy, int32_t* __restrict__ count) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.x;
    const int64_t start = cu_seqlens ? cu_seqlens[b] : 0;
    const int64_t end = cu_seqlens ? cu_seqlens[b + 1] : total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int64_t r0 = start; r0 < end; r0 += blockDim.x) {
        const int64_t r = r0 + threadIdx.x;
        float p = 0.f;
        bool keep = false;
        if (r < end) { p = probs[r * prob_stride]; keep = p > threshold; }          // strict `>` (train.py:140)
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base, all = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = s_warp[k];
            if (k < warp) before += c;
            all += c;
        }
        if (keep) {
            const int64_t o = start + before + __popc(m & ((1u << lane) - 1u));
            det_prob[o] = p;
            det_xy[o * 2] = pos[r * 2] * tile + half;                                // train.py:343
            det_xy[o * 2 + 1] = pos[r * 2 + 1] * tile + half;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[b] = s_base;
}

}  // namespace snuffy

#pragma GCC visibility push(default)
extern "C" {
using namespace snuffy;

int snuffy_patch_probs(const float* scores, int64_t n, float* probs, cudaStream_t stream) {
    SNUFFY_REQUIRE(n >= 0 && (n == 0 || (scores && probs)), "snuffy_patch_probs: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = (n + 1023) / 1024;
    const int64_t cap = 8 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    patch_probs_kernel<<<(unsigned)blocks, 256, 0, stream>>>(scores, n, probs);
    return check_launch("snuffy_patch_probs");
}

int snuffy_froc_detections(const float* probs, int64_t prob_stride, const int32_t* positions, const int32_t* cu_seqlens,
                           int64_t slides, int64_t total, float threshold, int32_t tile, int32_t half, float* det_prob,
                           int32_t* det_xy, int32_t* count, cudaStream_t stream) {
    SNUFFY_REQUIRE(slides >= 1 && slides <= 2147483647 && total >= 0 && prob_stride >= 1 && count,
                   "snuffy_froc_detections: bad dimensions");
    SNUFFY_REQUIRE(cu_seqlens || slides == 1, "snuffy_froc_detections: several slides need cu_seqlens");
    SNUFFY_REQUIRE(total == 0 || (probs && positions && det_prob && det_xy), "snuffy_froc_detections: null pointer");
    froc_detections_kernel<<<(unsigned)slides, 256, 0, stream>>>(probs, prob_stride, positions, cu_seqlens, total, threshold,
                                                                tile, half, det_prob, det_xy, count);
    return check_launch("snuffy_froc_detections");
}

}  // extern "C"
#pragma GCC visibility pop
