// Library-wide pieces of the C ABI: error reporting, version, device queries.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>

namespace snuffy {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what, int launches) {
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return 3;
    }
    return 0;
}

cudaError_t ensure_dynamic_smem(const void* func, int bytes) {
    // The attribute is a property of the kernel on the device, shared by all host threads (PyTorch runs backward passes on its
    // own autograd thread): one process-wide table, and the value only ever grows.
    struct Entry { const void* func; int dev; int bytes; };
    static Entry cache[128];
    static int n = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i)
        if (cache[i].func == func && cache[i].dev == dev) {
            if (cache[i].bytes >= bytes) return cudaSuccess;
            const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (e == cudaSuccess) cache[i].bytes = bytes;
            return e;
        }
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && n < 128) cache[n++] = Entry{func, dev, bytes};
    return e;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace snuffy

#pragma GCC visibility push(default)
extern "C" {

int snuffy_version(void) { return 100; }

const char* snuffy_last_error(void) { return snuffy::g_error; }

int snuffy_sm_count(void) { return snuffy::sm_count(); }

// kernels launched by this library since load (bench.py reports the per-step delta as gpu_launches)
long long snuffy_launch_count(void) { return snuffy::g_launches.load(); }

}  // extern "C"
#pragma GCC visibility pop
