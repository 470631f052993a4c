// Shared host/device helpers for libcsb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/csb200.h"

namespace csb {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

// Per-kernel profiling (csb_profile_begin/end): when on, a CUDA event is recorded on the launching stream after every
// launch; consecutive events bracket each kernel (kernels of one stream run back to back).
void profile_mark(const char* what, cudaStream_t st);
extern std::atomic<int> g_profiling;   // 0 off, 1 per kernel name, 2 detailed (conv launches keyed by layer shape; CSB_PROFILE_DETAIL=1)
const char* profile_intern(const char* text);   // stable copy of a transient label

// Call after every kernel launch: counts it and converts a launch error into a status.
inline int launched(const char* what, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CSB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    if (g_profiling.load(std::memory_order_relaxed)) profile_mark(what, st);
    return CSB_OK;
}

// memsets issued by the library are not counted as launches but must be visible to the profiler
inline void memset_done(cudaStream_t st) {
    if (g_profiling.load(std::memory_order_relaxed)) profile_mark("memset", st);
}

inline int cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) return fail(CSB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return CSB_OK;
}

#define CSB_REQUIRE(cond, msg)                                        \
    do {                                                              \
        if (!(cond)) return csb::fail(CSB_ERR_INVALID, "%s: %s", __func__, msg); \
    } while (0)

#define CSB_TRY(expr)                  \
    do {                               \
        int _st = (expr);              \
        if (_st != CSB_OK) return _st; \
    } while (0)

// tc_halo.cu: launch of the halo-tile conv kernel (also reached from csb_conv2d_nhwc for the dense shapes it suits)
int conv_halo_launch(const csb_conv_desc* d, const void* x, const void* w, const float* bias, const float* act_param, const void* residual, void* y, float* y_f32,
                     float* stats, void* stream);

// SM count of the CURRENT device (cached per device: one process may drive several GPUs, e.g. depth_est_device != device)
inline int num_sms() {
    static int cache[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cache[dev];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev] = n;
    }
    return n;
}

// true exactly once per device: guards per-device one-time setup such as cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which applies only to
// the device that is current when it runs.  `flags` is a zero-initialised static array of 64 entries owned by the call site.
inline bool first_use_on_device(unsigned char* flags) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (flags[dev]) return false;
    flags[dev] = 1;
    return true;
}

// Grid for a grid-stride kernel: a whole number of waves (multiple of the SM count), capped by the work.
inline int wave_grid(long long work_items, int threads, int ctas_per_sm) {
    long long need = (work_items + threads - 1) / threads;
    long long wave = (long long) num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int) (need < wave ? need : wave);
}

}  // namespace csb
