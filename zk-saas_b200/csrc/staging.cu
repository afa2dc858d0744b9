// staging.cu -- host <-> device copies for the host-pointer entry points.
//
// The reference hands us plain Rust `Vec<F>` / `Vec<G::Affine>` buffers, i.e. PAGEABLE memory.  The
// driver's own pageable path copies through one internal bounce buffer on the calling thread
// (measured here: 8 GB/s, a 2^22-point host-pointer MSM spends 52 ms in it against 15 ms with pinned
// buffers).  copy_h2d / copy_d2h detect pageable pointers and run their own pipeline instead:
// a few pinned slots, a small pool of threads that memcpy user memory <-> slot in parallel, and
// the DMA of slot k overlapping the memcpy of slot k+1.  Pinned (or registered) pointers and small
// copies go straight to cudaMemcpyAsync.  Semantics are those of cudaMemcpyAsync on pageable
// memory: on return from copy_h2d the source may be reused (DMAs may still be in flight on the
// stream); copy_d2h returns when the destination holds the data.
#include <condition_variable>
#include <cstring>
#include <emmintrin.h>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace zkg {

namespace {

static size_t SLOT_BYTES = 8u << 20;    // ZKG_STAGING_SLOT_MB
static int g_nt = 1;                     // ZKG_STAGING_NT: streaming stores into the pinned slot (no read-for-ownership of the destination)

// user memory -> pinned slot with streaming stores: the slot is written once and read only by the DMA engine, so the stores
// bypass the cache and skip the read-for-ownership of the destination.
// Pageable 2^22-point MSM: 18.3 -> 13.5 ms (pinned: 12.2).  Falls back to memcpy for a destination that is not 16-byte aligned.
static void copy_in(uint8_t* d, const uint8_t* s, size_t n) {
    size_t i = 0;
    if (g_nt && ((uintptr_t)d & 15) == 0) {
        for (; i + 64 <= n; i += 64) {
            __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 16));
            __m128i c = _mm_loadu_si128((const __m128i*)(s + i + 32)), e = _mm_loadu_si128((const __m128i*)(s + i + 48));
            _mm_stream_si128((__m128i*)(d + i), a);
            _mm_stream_si128((__m128i*)(d + i + 16), b);
            _mm_stream_si128((__m128i*)(d + i + 32), c);
            _mm_stream_si128((__m128i*)(d + i + 48), e);
        }
        _mm_sfence();
    }
    if (i < n) memcpy(d + i, s + i, n - i);
}
constexpr int N_SLOTS = 4;
constexpr int MAX_DEV = 16;
constexpr size_t STAGE_MIN = 1u << 20;       // below this the driver's path is as good

// persistent helpers that split one memcpy; heap-allocated and never destroyed (threads are detached
// and sleep on the condition variable until the process exits)
struct CopyPool {
    struct Job { uint8_t* d; const uint8_t* s; size_t n; bool in; };
    std::mutex m;
    std::condition_variable cv, done_cv;
    std::vector<Job> jobs;
    size_t next = 0, pending = 0;
    int n_threads = 0;

    explicit CopyPool(int threads) : n_threads(threads) {
        for (int i = 0; i < threads; ++i) std::thread([this] { worker(); }).detach();
    }
    bool pop(Job* j) {                       // caller holds m
        if (next >= jobs.size()) return false;
        *j = jobs[next++];
        return true;
    }
    void finish_one() {
        std::lock_guard<std::mutex> lk(m);
        if (--pending == 0) done_cv.notify_all();
    }
    void worker() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return next < jobs.size(); });
                pop(&j);
            }
            if (j.in) copy_in(j.d, j.s, j.n); else memcpy(j.d, j.s, j.n);
            finish_one();
        }
    }
    // one caller at a time (the stager's mutex serialises callers)
    void run(uint8_t* d, const uint8_t* s, size_t n, bool in = false) {
        const size_t part_min = 256u << 10;
        size_t parts = n / part_min;
        if (parts > (size_t)n_threads + 1) parts = (size_t)n_threads + 1;
        if (parts <= 1) { if (in) copy_in(d, s, n); else memcpy(d, s, n); return; }
        const size_t per = (n / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(m);
            jobs.clear();
            next = 0;
            for (size_t off = 0; off < n; off += per) jobs.push_back({d + off, s + off, off + per <= n ? per : n - off, in});
            pending = jobs.size();
        }
        cv.notify_all();
        for (;;) {                           // the caller works too
            Job j;
            {
                std::lock_guard<std::mutex> lk(m);
                if (!pop(&j)) break;
            }
            if (j.in) copy_in(j.d, j.s, j.n); else memcpy(j.d, j.s, j.n);
            finish_one();
        }
        std::unique_lock<std::mutex> lk(m);
        done_cv.wait(lk, [&] { return pending == 0; });
    }
};

struct Stager {
    std::mutex mu;                           // one staged transfer at a time per process (they share PCIe and DRAM anyway)
    uint8_t* slot[N_SLOTS] = {};
    cudaEvent_t ev[MAX_DEV][N_SLOTS] = {};
    int slot_dev[N_SLOTS];                   // device whose event guards the slot's last use, -1 = free
    CopyPool* pool = nullptr;
    bool ready = false;
    bool enabled = true;

    int32_t init() {
        if (ready) return ZKG_OK;
        const char* e = getenv("ZKG_STAGING");
        enabled = !(e && e[0] == '0');
        if (const char* mb = getenv("ZKG_STAGING_SLOT_MB")) { int v = atoi(mb); if (v >= 1 && v <= 256) SLOT_BYTES = (size_t)v << 20; }
        if (const char* nt = getenv("ZKG_STAGING_NT")) g_nt = nt[0] != '0';
        for (int s = 0; s < N_SLOTS; ++s) {
            ZKG_CUDA(cudaHostAlloc((void**)&slot[s], SLOT_BYTES, cudaHostAllocPortable));
            slot_dev[s] = -1;
        }
        unsigned hc = std::thread::hardware_concurrency();
        int threads = hc >= 16 ? 11 : hc >= 8 ? 5 : hc >= 4 ? 3 : 1;   // 16 cores, 2^22-point pageable MSM: 7 helpers 13.8-14.9 ms, 11 13.2-13.9, 15 13.3-13.4
        const char* t = getenv("ZKG_STAGING_THREADS");
        if (t && atoi(t) >= 0 && atoi(t) <= 32) threads = atoi(t);
        pool = new CopyPool(threads);
        ready = true;
        return ZKG_OK;
    }
    int32_t event_for(int dev, int s, cudaEvent_t* out) {
        ZKG_REQUIRE(dev >= 0 && dev < MAX_DEV, "staging: device %d out of range", dev);
        if (!ev[dev][s]) ZKG_CUDA(cudaEventCreateWithFlags(&ev[dev][s], cudaEventDisableTiming));
        *out = ev[dev][s];
        return ZKG_OK;
    }
    int32_t wait_slot(int s) {
        if (slot_dev[s] >= 0) {
            ZKG_CUDA(cudaEventSynchronize(ev[slot_dev[s]][s]));
            slot_dev[s] = -1;
        }
        return ZKG_OK;
    }
};

Stager* stager() {
    static Stager* g = new Stager();         // leaked on purpose: outlives every caller
    return g;
}

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace

int32_t copy_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return ZKG_OK;
    Stager* S = stager();
    if (bytes < STAGE_MIN || !is_pageable(h_src)) {
        ZKG_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return ZKG_OK;
    }
    std::lock_guard<std::mutex> lk(S->mu);
    ZKG_TRY(S->init());
    if (!S->enabled) {
        ZKG_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return ZKG_OK;
    }
    int dev = 0;
    ZKG_CUDA(cudaGetDevice(&dev));
    int k = 0;
    for (size_t off = 0; off < bytes; off += SLOT_BYTES, ++k) {
        const int s = k % N_SLOTS;
        const size_t len = off + SLOT_BYTES <= bytes ? SLOT_BYTES : bytes - off;
        ZKG_TRY(S->wait_slot(s));
        S->pool->run(S->slot[s], (const uint8_t*)h_src + off, len, true);
        ZKG_CUDA(cudaMemcpyAsync((uint8_t*)d_dst + off, S->slot[s], len, cudaMemcpyHostToDevice, st));
        cudaEvent_t e;
        ZKG_TRY(S->event_for(dev, s, &e));
        ZKG_CUDA(cudaEventRecord(e, st));
        S->slot_dev[s] = dev;
    }
    return ZKG_OK;
}

int32_t copy_d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return ZKG_OK;
    Stager* S = stager();
    if (bytes < STAGE_MIN || !is_pageable(h_dst)) {
        ZKG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        return ZKG_OK;
    }
    std::lock_guard<std::mutex> lk(S->mu);
    ZKG_TRY(S->init());
    if (!S->enabled) {
        ZKG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        return ZKG_OK;
    }
    int dev = 0;
    ZKG_CUDA(cudaGetDevice(&dev));
    const size_t n_chunks = (bytes + SLOT_BYTES - 1) / SLOT_BYTES;
    // DMA of chunk k runs while chunk k - (N_SLOTS - 1) is copied out of its slot
    for (size_t k = 0; k < n_chunks + (N_SLOTS - 1); ++k) {
        if (k < n_chunks) {
            const int s = (int)(k % N_SLOTS);
            const size_t off = k * SLOT_BYTES, len = off + SLOT_BYTES <= bytes ? SLOT_BYTES : bytes - off;
            ZKG_TRY(S->wait_slot(s));        // an earlier h2d may still be reading the slot
            ZKG_CUDA(cudaMemcpyAsync(S->slot[s], (const uint8_t*)d_src + off, len, cudaMemcpyDeviceToHost, st));
            cudaEvent_t e;
            ZKG_TRY(S->event_for(dev, s, &e));
            ZKG_CUDA(cudaEventRecord(e, st));
            S->slot_dev[s] = dev;
        }
        if (k >= (size_t)(N_SLOTS - 1)) {
            const size_t j = k - (N_SLOTS - 1);
            const int s = (int)(j % N_SLOTS);
            const size_t off = j * SLOT_BYTES, len = off + SLOT_BYTES <= bytes ? SLOT_BYTES : bytes - off;
            ZKG_TRY(S->wait_slot(s));
            S->pool->run((uint8_t*)h_dst + off, S->slot[s], len);           // (streaming stores measured no faster in this direction)
        }
    }
    return ZKG_OK;
}

}  // namespace zkg
