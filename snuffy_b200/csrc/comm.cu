// Gradient all-reduce of the data-parallel training step over NVLink peer memory (the one exchange step of
// the path; the reference trains one process on one GPU, train.py:249-264, so there is no reference code for it).
//
// Every rank owns one cudaMalloc'd block [gradient buffer | barrier counters | call counter] that all its peers on the node
// have mapped through CUDA IPC.  One kernel per step, capturable in the step's CUDA graph:
//
//   barrier 1   every rank's gradient buffer is complete (its producers precede this kernel on the stream)
//   phase 1     rank r sums slice r of all W buffers, reading its peers' memory directly, in rank order (so the result does not
//               depend on timing, and every rank ends up with the same bits)
//   phase 2     ... and writes the sum into slice r of all W buffers
//   barrier 2   all slices of this rank's buffer have been written
//
// = a two-shot all-reduce: each GPU moves 2 (W - 1) / W of the buffer over NVLink, half in each direction, against NCCL's
// ~0.14 ms latency-bound ring/tree for the 12.6 MB flat gradient of cfg2 at 8 ranks.  The barriers are per CTA: CTA c of every
// rank only exchanges data with the CTAs c of its peers (sub-chunk c of each slice), so no grid-wide synchronisation is needed.
#include <stdlib.h>
#include "common.cuh"

namespace snuffy {

constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_CTAS = 64;
constexpr int COMM_THREADS = 512;

struct PeerComm {
    float* buf[COMM_MAX_WORLD];                  // each rank's gradient buffer (own entry: local memory)
    unsigned int* cnt[COMM_MAX_WORLD];           // each rank's barrier counters [COMM_CTAS][2]
    unsigned int* state;                         // own: [0] calls completed, [1] CTAs done in this call
    int rank, world;
    int64_t n4;                                  // float4 elements
    unsigned long long timeout_ns;               // a peer that has not arrived by then: trap (fail loudly instead of hanging)
};

__device__ __forceinline__ unsigned long long comm_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Signals every peer's counter of this CTA and phase, then waits until all W ranks have signalled this rank's `target` times.
__device__ __forceinline__ void comm_barrier(const PeerComm& c, int phase, unsigned int target) {
    __syncthreads();
    if ((int)threadIdx.x < c.world) {
        __threadfence_system();
        unsigned int* peer = c.cnt[threadIdx.x] + blockIdx.x * 2 + phase;
        asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(peer) : "memory");
    }
    if (threadIdx.x == 0) {
        const unsigned int* mine = c.cnt[c.rank] + blockIdx.x * 2 + phase;
        const unsigned long long t0 = comm_now_ns();
        unsigned int seen;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if ((int)(seen - target) >= 0) break;
            if (comm_now_ns() - t0 > c.timeout_ns) asm volatile("trap;");        // a peer never arrived: fail, do not hang
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(COMM_THREADS)
peer_allreduce_kernel(const __grid_constant__ PeerComm c) {
    const unsigned int call = *reinterpret_cast<volatile unsigned int*>(c.state) + 1u;
    const unsigned int target = call * (unsigned int)c.world;
    const int W = c.world;
    const int64_t slice = (c.n4 + W - 1) / W;
    const int64_t sub = (slice + gridDim.x - 1) / gridDim.x;
    const int64_t s0 = (int64_t)c.rank * slice + (int64_t)blockIdx.x * sub;
    int64_t s1 = s0 + sub;
    const int64_t slice_end = (int64_t)(c.rank + 1) * slice < c.n4 ? (int64_t)(c.rank + 1) * slice : c.n4;
    if (s1 > slice_end) s1 = slice_end;

    comm_barrier(c, 0, target);
    for (int64_t i0 = s0 + threadIdx.x; i0 < s1; i0 += 2 * COMM_THREADS) {
        const int64_t i1 = i0 + COMM_THREADS;
        const bool two = i1 < s1;
        float4 a[COMM_MAX_WORLD], b[COMM_MAX_WORLD];
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q) {
            if (q < W) {
                a[q] = __ldcg(reinterpret_cast<const float4*>(c.buf[q]) + i0);
                if (two) b[q] = __ldcg(reinterpret_cast<const float4*>(c.buf[q]) + i1);
            }
        }
        float4 sa = a[0], sb = b[0];
#pragma unroll
        for (int q = 1; q < COMM_MAX_WORLD; ++q) {
            if (q < W) {
                sa.x += a[q].x; sa.y += a[q].y; sa.z += a[q].z; sa.w += a[q].w;
                if (two) { sb.x += b[q].x; sb.y += b[q].y; sb.z += b[q].z; sb.w += b[q].w; }
            }
        }
#pragma unroll
        for (int q = 0; q < COMM_MAX_WORLD; ++q) {
            if (q < W) {
                __stcg(reinterpret_cast<float4*>(c.buf[q]) + i0, sa);
                if (two) __stcg(reinterpret_cast<float4*>(c.buf[q]) + i1, sb);
            }
        }
    }
    comm_barrier(c, 1, target);
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(c.state + 1, 1u) == gridDim.x - 1) {       // last CTA of this rank: the call is complete
            c.state[1] = 0u;
            __threadfence();
            c.state[0] = call;
        }
    }
}

}  // namespace snuffy

using namespace snuffy;

#pragma GCC visibility push(default)
extern "C" {

// Plain cudaMalloc (IPC handles cannot be taken from a pooled / virtual-memory allocation), zero-filled.
int snuffy_comm_alloc(int64_t bytes, void** ptr) {
    SNUFFY_REQUIRE(ptr && bytes > 0, "snuffy_comm_alloc: bad arguments");
    SNUFFY_CUDA(cudaMalloc(ptr, (size_t)bytes));
    SNUFFY_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    SNUFFY_CUDA(cudaDeviceSynchronize());
    return 0;
}
int snuffy_comm_free(void* ptr) {
    if (ptr) SNUFFY_CUDA(cudaFree(ptr));
    return 0;
}
int snuffy_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }
int snuffy_comm_export(void* ptr, void* handle_out) {
    SNUFFY_REQUIRE(ptr && handle_out, "snuffy_comm_export: null pointer");
    cudaIpcMemHandle_t h;
    SNUFFY_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, sizeof h);
    return 0;
}
int snuffy_comm_import(const void* handle, void** ptr) {
    SNUFFY_REQUIRE(handle && ptr, "snuffy_comm_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    SNUFFY_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int snuffy_comm_close(void* ptr) {
    if (ptr) SNUFFY_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
int snuffy_comm_counter_bytes(void) { return COMM_CTAS * 2 * (int)sizeof(unsigned int); }

// In-place sum over `world` ranks of the fp32 buffers bufs[0..world) (n elements each, n % 4 == 0; bufs[rank] is this rank's,
// the others are its peers' blocks mapped with snuffy_comm_import).  counters[q]: rank q's barrier counters
// (snuffy_comm_counter_bytes, zero at start); state: this rank's two uint32 (zero at start).  Every rank of the group must make
// the same sequence of calls.
int snuffy_peer_allreduce(void* const* bufs, void* const* counters, void* state, int rank, int world, int64_t n,
                          cudaStream_t stream) {
    SNUFFY_REQUIRE(bufs && counters && state, "snuffy_peer_allreduce: null pointer");
    SNUFFY_REQUIRE(world >= 2 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, "snuffy_peer_allreduce: 2..8 ranks");
    SNUFFY_REQUIRE(n >= 4 && n % 4 == 0, "snuffy_peer_allreduce: n must be a positive multiple of 4");
    PeerComm c{};
    for (int q = 0; q < world; ++q) {
        SNUFFY_REQUIRE(bufs[q] && counters[q] && (uintptr_t)bufs[q] % 16 == 0, "snuffy_peer_allreduce: bad peer pointer");
        c.buf[q] = reinterpret_cast<float*>(bufs[q]);
        c.cnt[q] = reinterpret_cast<unsigned int*>(counters[q]);
    }
    c.state = reinterpret_cast<unsigned int*>(state);
    c.rank = rank; c.world = world; c.n4 = n / 4;
    // ranks reach the exchange at different times (data loading, a graph capture on one of them): wait long, like NCCL's
    // watchdog, before declaring the peer lost.  SNUFFY_B200_PEER_TIMEOUT_S overrides (tests use a short one).
    static const unsigned long long timeout_s = [] {
        const char* e = getenv("SNUFFY_B200_PEER_TIMEOUT_S");
        const long long v = e ? atoll(e) : 0;
        return (unsigned long long)(v > 0 ? v : 600);
    }();
    c.timeout_ns = timeout_s * 1000000000ull;
    peer_allreduce_kernel<<<COMM_CTAS, COMM_THREADS, 0, stream>>>(c);
    return check_launch("snuffy_peer_allreduce");
}

}  // extern "C"
#pragma GCC visibility pop
