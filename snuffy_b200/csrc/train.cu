// Training-loop glue on the device (reference caller: train.py:828-846, 468-473, 852-854):
//   mil_loss      max over instances + 2 x BCE-with-logits + convex mix, forward and backward in one launch
//   adamw_flat    AdamW over one flat fp32 parameter / gradient buffer (train.py:809-826: optim.AdamW), with the
//                 optional global-norm clip of train.py:469-470 folded in
//   sumsq         deterministic two-stage sum of squares (the clip's gradient norm)
// These remove the per-bag host syncs and the tiny elementwise launches of the reference loop; they are used by
// snuffy_b200/dp.py (the data-parallel trainer), never by the drop-in modules themselves.
#include "common.cuh"

namespace snuffy {

__device__ __forceinline__ float bce_logits(float z, float y) {
    // -[y log s(z) + (1 - y) log(1 - s(z))] = max(z, 0) - z y + log1p(exp(-|z|))
    return fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }
// arg-max order of torch.max: larger value first, NaN above everything (it propagates into the loss), ties to the lower index
__device__ __forceinline__ bool max_takes(float v, int64_t n, float best, int64_t bi) {
    const bool vn = v != v, bn = best != best;
    if (vn || bn) return vn && (!bn || n < bi);
    return v > best || (v == best && n < bi);
}

// grid = B*C CTAs: CTA (b, c) finds max_n classes[b, n, c] (ties: lowest n) and writes its two loss terms; the last CTA
// (ticket) folds them in a fixed order into loss[0..2] = (mixed, bag term, max term) and zeroes the ticket.
__global__ void __launch_bounds__(256)
mil_loss_kernel(const float* __restrict__ classes, const float* __restrict__ bag, const float* __restrict__ label,
                const float* __restrict__ weight, int64_t N, int C, int BC, float w, const float* __restrict__ w_dev,
                float gscale, float* __restrict__ terms, unsigned int* __restrict__ ticket, float* __restrict__ loss,
                float* __restrict__ pred, float* __restrict__ dclasses, float* __restrict__ dbag, float* __restrict__ dw) {
    if (w_dev) w = __ldg(w_dev);                                     // learnable mix weight (train.py:804, --soft_average)
    __shared__ float s_val[8];
    __shared__ int64_t s_idx[8];
    __shared__ int s_last;
    const int bc = blockIdx.x, b = bc / C, c = bc % C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* col = classes + (int64_t)b * N * C + c;
    float best = -INFINITY; int64_t bi = 0x7fffffffffffffffll;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
        const float v = col[n * C];
        if (max_takes(v, n, best, bi)) { best = v; bi = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (max_takes(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k)
            if (max_takes(s_val[k], s_idx[k], best, bi)) { best = s_val[k]; bi = s_idx[k]; }
        const float y = label[bc], zb = bag[bc];
        const float wt = weight ? weight[c] : 1.f;                   // nn.BCEWithLogitsLoss(weight) (train.py:245-246)
        terms[bc * 2] = wt * bce_logits(zb, y);
        terms[bc * 2 + 1] = wt * bce_logits(best, y);
        if (pred) pred[bc] = (1.f - w) * sigmoidf_(best) + w * sigmoidf_(zb);        // train.py:840-844
        if (dbag) dbag[bc] = gscale * w * wt * (sigmoidf_(zb) - y) / (float)BC;
        if (dclasses) dclasses[((int64_t)b * N + bi) * C + c] = gscale * (1.f - w) * wt * (sigmoidf_(best) - y) / (float)BC;
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == (unsigned)BC - 1u);
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    float lb = 0.f, lm = 0.f;
    for (int k = 0; k < BC; ++k) { lb += __ldcg(terms + k * 2); lm += __ldcg(terms + k * 2 + 1); }
    lb /= (float)BC; lm /= (float)BC;                               // reduction = 'mean'
    loss[0] = w * lb + (1.f - w) * lm; loss[1] = lb; loss[2] = lm;
    if (dw) dw[0] = gscale * (lb - lm);                              // d loss / d w
    *ticket = 0;
}

// counter += delta: the last node of a captured training step, so that the next replay draws fresh dropout masks / random
// patches (common.cuh rng_resolve).  Every kernel of the replay itself reads the value from before this node.
__global__ void rng_advance_kernel(unsigned long long* counter, unsigned long long delta) { *counter += delta; }

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partials) {
    __shared__ float red[8];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc = fmaf(x[i], x[i], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k];
        partials[blockIdx.x] = s;
    }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partials, int nparts, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float s = 0.f;
    for (int k = 0; k < nparts; ++k) s += partials[k];
    out[0] = s;
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction), one launch over the flat buffers.
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                  float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                  const float* __restrict__ gnorm_sq, float max_norm, const long long* __restrict__ step_dev,
                  const float* __restrict__ contributors, const float* __restrict__ lr_dev, float clamp_lo, float clamp_hi) {
    if (step_dev) {                                                  // captured step: the count lives on the device
        const double t = (double)*step_dev;
        bc1 = 1.f - (float)pow((double)beta1, t);
        bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
    }
    if (lr_dev) lr = __ldg(lr_dev);
    if (contributors) {                                              // ranks that had a bag this step (sum all-reduced with the gradient)
        const float cnt = __ldg(contributors);
        if (!(cnt > 0.f)) return;
        gscale /= cnt;
    }
    float coef = gscale;
    if (gnorm_sq) coef *= fminf(1.f, max_norm / (sqrtf(gnorm_sq[0]) * fabsf(gscale) + 1e-6f));   // clip_grad_norm_
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
        if (clamp_lo <= clamp_hi) pi = fminf(fmaxf(pi, clamp_lo), clamp_hi);      // train.py:852-854 (mix weight in [0, 1])
        p[i] = pi;
    }
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// loss = w * BCEwL(bag, y) + (1 - w) * BCEwL(max_n classes, y)   (train.py:831-838), mean over the B*C logits.
// classes [B, N, C], bag / label [B, C], weight [C] or null; workspace: 2*B*C floats + one zeroed uint32 ticket.
// Outputs: loss[3] = (mixed, bag term, max term); optional pred [B, C] (train.py:840-844); optional gradients scaled by
// gscale: dbag [B, C] and dclasses [B, N, C] (caller-zeroed; only the arg-max rows are written).
// w_dev (optional): the mix weight is read from the device (a learnable parameter); dw (optional): d loss / d w * gscale.
int snuffy_mil_loss(const float* classes, const float* bag, const float* label, const float* weight, int64_t B,
                    int64_t N, int64_t C, float w, const float* w_dev, float gscale, float* terms, uint32_t* ticket,
                    float* loss, float* pred, float* dclasses, float* dbag, float* dw, cudaStream_t stream) {
    SNUFFY_REQUIRE(classes && bag && label && terms && ticket && loss, "snuffy_mil_loss: null pointer");
    SNUFFY_REQUIRE(B >= 1 && N >= 1 && C >= 1 && B * C <= 65535, "snuffy_mil_loss: bad dimensions");
    mil_loss_kernel<<<(unsigned)(B * C), 256, 0, stream>>>(classes, bag, label, weight, N, (int)C, (int)(B * C), w, w_dev,
                                                          gscale, terms, ticket, loss, pred, dclasses, dbag, dw);
    return check_launch("snuffy_mil_loss");
}

int snuffy_rng_advance(uint64_t* counter, uint64_t delta, cudaStream_t stream) {
    SNUFFY_REQUIRE(counter, "snuffy_rng_advance: null counter");
    rng_advance_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(counter), (unsigned long long)delta);
    return check_launch("snuffy_rng_advance");
}

int64_t snuffy_sumsq_blocks(int64_t n) {
    int64_t b = (n + 1023) / 1024;
    const int64_t cap = 4 * (int64_t)sm_count();
    if (b > cap) b = cap;
    return b < 1 ? 1 : b;
}
// out[0] = sum x^2 (deterministic); partials: snuffy_sumsq_blocks(n) floats
int snuffy_sumsq(const float* x, int64_t n, float* partials, float* out, cudaStream_t stream) {
    SNUFFY_REQUIRE(x && partials && out && n >= 0, "snuffy_sumsq: bad arguments");
    const int64_t blocks = snuffy_sumsq_blocks(n);
    sumsq_partial_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, partials);
    sumsq_final_kernel<<<1, 32, 0, stream>>>(partials, (int)blocks, out);
    return check_launch("snuffy_sumsq", 2);
}

// One AdamW step (torch.optim.AdamW semantics) over flat fp32 buffers; step >= 1.  Gradients are first scaled by gscale
// (1 / world_size after a sum all-reduce) and, when gnorm_sq is given (sum of squares of the UNSCALED gradient),
// clipped to max_norm like torch.nn.utils.clip_grad_norm_ (train.py:469-470).
int snuffy_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int64_t step, float gscale, const float* gnorm_sq, float max_norm,
                      cudaStream_t stream) {
    SNUFFY_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "snuffy_adamw_flat: bad arguments");
    if (n == 0) return 0;
    const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = 8 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    adamw_flat_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s,
                                                           gscale, gnorm_sq, max_norm, nullptr, nullptr, nullptr, 1.f, 0.f);
    return check_launch("snuffy_adamw_flat");
}

// The same step with everything that changes from step to step read from the device, so that it can be a node of a captured
// CUDA graph: step_dev (int64, the 1-based step count; bump it with snuffy_rng_advance(step_dev, 1) before this call),
// contributors (optional float: number of ranks whose gradient is in the sum; the update divides by it and is skipped when
// it is 0), lr_dev (optional float replacing `lr`, for schedulers).  clamp_lo <= clamp_hi clamps the updated parameters
// (train.py:852-854 keeps the learnable mix weight in [0, 1]); pass clamp_lo > clamp_hi for none.
int snuffy_adamw_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, const int64_t* step_dev, const float* contributors,
                          const float* lr_dev, float gscale, const float* gnorm_sq, float max_norm, float clamp_lo,
                          float clamp_hi, cudaStream_t stream) {
    SNUFFY_REQUIRE(p && g && m && v && n >= 0 && step_dev, "snuffy_adamw_flat_dev: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = 8 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    adamw_flat_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, gscale,
                                                           gnorm_sq, max_norm, reinterpret_cast<const long long*>(step_dev),
                                                           contributors, lr_dev, clamp_lo, clamp_hi);
    return check_launch("snuffy_adamw_flat_dev");
}

}  // extern "C"
#pragma GCC visibility pop

// ------------------------------------------------------------------ gradient packing
// Gathers up to 32 tensors per launch into one flat buffer (dst + offsets[i]): ONE launch replaces autograd's per-parameter
// `grad += g` kernels when the gradients are needed as a single all-reduce / optimizer buffer (dp.FlatBuffers.pack).
namespace snuffy {
struct PackArgs { const float* src[32]; long long size[32]; long long off[32]; int n; };
__global__ void __launch_bounds__(256)
pack_f32_kernel(const PackArgs a, float* __restrict__ dst) {
    const int t = blockIdx.y;
    if (t >= a.n) return;
    const float* s = a.src[t];
    float* d = dst + a.off[t];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.size[t]; i += (long long)gridDim.x * blockDim.x)
        d[i] = s[i];
}
}  // namespace snuffy

#pragma GCC visibility push(default)
extern "C" int snuffy_pack_f32(const void* const* srcs, const int64_t* sizes, const int64_t* offsets, int64_t n, float* dst,
                               cudaStream_t stream) {
    using namespace snuffy;
    SNUFFY_REQUIRE(srcs && sizes && offsets && dst && n >= 0, "snuffy_pack_f32: bad arguments");
    int launches = 0;
    for (int64_t base = 0; base < n; base += 32) {
        PackArgs a{};
        a.n = (int)((n - base) < 32 ? (n - base) : 32);
        long long biggest = 1;
        for (int i = 0; i < a.n; ++i) {
            a.src[i] = reinterpret_cast<const float*>(srcs[base + i]);
            a.size[i] = sizes[base + i]; a.off[i] = offsets[base + i];
            SNUFFY_REQUIRE(a.src[i] || a.size[i] == 0, "snuffy_pack_f32: null source");
            if (a.size[i] > biggest) biggest = a.size[i];
        }
        long long bx = (biggest + 255) / 256;
        if (bx > 64) bx = 64;
        dim3 grid((unsigned)bx, (unsigned)a.n);
        pack_f32_kernel<<<grid, 256, 0, stream>>>(a, dst);
        ++launches;
    }
    return launches ? check_launch("snuffy_pack_f32", launches) : 0;
}
#pragma GCC visibility pop
