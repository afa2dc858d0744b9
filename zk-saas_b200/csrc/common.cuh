// common.cuh -- contexts, workspaces and error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/zksaas_gpu.h"

namespace zkg {

void set_error(const char* fmt, ...);

#define ZKG_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            zkg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? ZKG_ERR_OOM : ZKG_ERR_CUDA;                    \
        }                                                                                           \
    } while (0)

#define ZKG_TRY(call)                   \
    do {                                \
        int32_t _r = (call);            \
        if (_r != ZKG_OK) return _r;    \
    } while (0)

#define ZKG_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            zkg::set_error(__VA_ARGS__); \
            return ZKG_ERR_BAD_ARG;     \
        }                               \
    } while (0)

// Grow-only device buffer (one cudaMalloc per high-water mark, never per call)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int32_t reserve(size_t need) {
        if (need <= bytes) return ZKG_OK;
        if (p) { ZKG_CUDA(cudaFree(p)); p = nullptr; bytes = 0; }
        size_t want = need + need / 8;           // slack so slowly growing sizes do not realloc
        ZKG_CUDA(cudaMalloc(&p, want));
        bytes = want;
        return ZKG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace zkg

struct zkg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    zkg::DevBuf ws;        // kernel workspace (digits, sorted indices, buckets, NTT scratch ...)
    zkg::DevBuf io;        // staging for host-pointer entry points
    int sm_count = 0;
    // persistent device-side parameter cache (twiddle/power tables, PSS matrices), keyed by content
    struct CacheEnt { uint64_t h1, h2; size_t bytes; void* p; };
    std::vector<CacheEnt> cache;
    size_t cache_bytes = 0;            // sum of the entries' sizes (bounded: see ctx_cache_get)
    std::vector<void*> cache_dead;     // entries dropped by the last flush; freed by the next one (pointers handed out
                                       // earlier in the call that triggered the flush stay valid until it returns)
    // second stream + events for overlapping H2D copies with compute in host-pointer entry points
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev[16] = {};
    int copy_ev_count = 0;
    // high-priority side stream + events of the MSM sort pipeline (msm_impl.cuh: the counting sort of chunk k+1 runs
    // under the bucket accumulation of chunk k)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_ev[5] = {};        // 0: inputs ready (main -> side), 1-2: sorted[set], 3-4: accumulated[set]
    // instrumentation (bench.py): kernels launched so far, and optional per-phase CUDA events
    uint64_t launches = 0;
    bool profile = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // phase boundaries of the last profiled call
    int ev_count = 0;
};

namespace zkg {

// RAII: borrow a pooled context for a blocking host-pointer call
struct PooledCtx {
    zkg_ctx* ctx = nullptr;
    int32_t acquire(int device);
    ~PooledCtx();
};

void auto_register_clear();             // msm_api.cu: drops the transparent registrations of ZKG_AUTO_REGISTER (zkg_shutdown)
int32_t ctx_copy_stream(zkg_ctx* ctx, int n_events);
int32_t ctx_aux_stream(zkg_ctx* ctx);
// host <-> device copies that take pageable host memory at PCIe speed (staging.cu); stream semantics of
// cudaMemcpyAsync on pageable memory
int32_t copy_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
int32_t copy_d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st);
// Look up (or reserve) a persistent device block for the parameter identified by `key`.
// *fresh = true means the caller must fill it (on ctx->stream) before use.
int32_t ctx_cache_get(zkg_ctx* ctx, const void* key, size_t key_bytes, size_t bytes, void** out, bool* fresh);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// mark a phase boundary on the context's stream when profiling is enabled (no-op otherwise)
static inline void phase_mark(zkg_ctx* ctx, int idx) {
    if (!ctx->profile || idx >= 4) return;
    if (!ctx->ev[idx]) cudaEventCreate(&ctx->ev[idx]);
    cudaEventRecord(ctx->ev[idx], ctx->stream);
    if (idx + 1 > ctx->ev_count) ctx->ev_count = idx + 1;
}

}  // namespace zkg
