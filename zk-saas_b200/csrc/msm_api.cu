// msm_api.cu -- C-ABI entry points of the MSM path (dispatch over G1/G2 + device-resident CRS registry).
// Declarations and reference citations: include/zksaas_gpu.h.
#include "common.cuh"
#include "host_fr.hpp"
#include <condition_variable>
#include <string>
#include <thread>
#include <utility>

namespace zkg {
#define ZKG_MSM_DECLARE(G)                                                                                          \
    int32_t msm_run_##G(zkg_ctx* ctx, const void* d_bases, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t pack_bases_##G(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed);               \
    int32_t msm_host_##G(int device, const void* bases, size_t stride, size_t n_bases, const uint64_t* scalars,     \
                         size_t n_scalars, uint64_t* out_xyz);                                                      \
    int32_t msm_host_sharded_##G(const int32_t* devices, int32_t n_dev, const void* bases, size_t stride, size_t n_bases, \
                                 const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz);                     \
    int32_t combine_##G(zkg_ctx* ctx, const uint64_t* d_parts, size_t n, uint64_t* d_out);                          \
    int32_t fixed_base_##G(zkg_ctx* ctx, const uint64_t* d_scalars, size_t n, void* d_packed);                      \
    int32_t prepare_##G(zkg_ctx* ctx, const void* d_bases, size_t n, int c, void* d_table);                         \
    int32_t msm_run_prepared_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t msm_run_prepared_host_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* h_scalars, size_t n, void* d_out, int mode); \
    int32_t crs_det_pack_##G(int device, const void* bases, size_t stride, size_t n, int l, int n_parties,          \
                             const uint32_t* h_scal, void* const* out_by_party, size_t out_stride);
ZKG_MSM_DECLARE(g1)
ZKG_MSM_DECLARE(g2)
#define ZKG_GROUP_DECLARE(G)                                                                                              \
    int32_t group_unpack_##G(int device, const uint64_t* shares_xyz, uint32_t n_recv, const uint32_t* h_scal, uint32_t rows, \
                             uint64_t* out_rows, uint64_t* out_sum);                                                     \
    int32_t point_wire_##G(int device, int dir, const void* in, void* out, size_t n);
ZKG_GROUP_DECLARE(g1)
ZKG_GROUP_DECLARE(g2)

// device-resident CRS share: table[w*n + i] = 2^(c*w) * P_i (packed affine), W = 254/c + 1 window shifts.
// Lifetime: a handle is (generation << 32) | (slot + 1); slots are reused, stale handles are rejected by the
// generation.  A call that uses the table holds a reference from lookup until its work is ordered behind an
// event (`_dev` entry points: recorded on the context's stream right after the enqueue) or finished (blocking
// entry points); zkg_bases_release waits for the references to drain and for every recorded event before it
// frees the table, so a release that races an MSM on another thread can never free memory a kernel still reads.
struct BaseSet {
    int device = 0, group = 0;
    size_t n = 0;
    int c = 0;
    void* d_table = nullptr;
    uint32_t generation = 0;
    int refs = 0;                    // calls between lookup and "ordered behind an event / finished"
    bool released = false;
    std::vector<std::pair<zkg_ctx*, cudaEvent_t>> last_use;     // one event per context that enqueued work on the table
    // a set registered over several GPUs (zkg_bases_register_sharded): per-device handles of consecutive point ranges
    std::vector<uint64_t> shard_handles;
    std::vector<size_t> shard_lo;                               // shard k owns points [shard_lo[k], shard_lo[k+1])
};
int msm_pick_c_merged_host(size_t n);
static std::mutex g_bases_mu;
static std::condition_variable g_bases_cv;
static std::vector<BaseSet*> g_bases;        // slot -> live set or nullptr
static std::vector<uint32_t> g_bases_gen;    // slot -> generation of the last set stored there
static inline size_t packed_bytes(int group) { return group == 1 ? 64 : 128; }

static BaseSet* base_set_lookup_locked(uint64_t handle) {
    const uint64_t slot = (handle & 0xffffffffu), gen = handle >> 32;
    if (slot < 1 || slot > g_bases.size()) return nullptr;
    BaseSet* bs = g_bases[slot - 1];
    if (!bs || bs->released || bs->generation != (uint32_t)gen) return nullptr;
    return bs;
}
// RAII reference on a registered base set
struct BaseRef {
    BaseSet* bs = nullptr;
    int32_t acquire(uint64_t handle) {
        std::lock_guard<std::mutex> lk(g_bases_mu);
        bs = base_set_lookup_locked(handle);
        ZKG_REQUIRE(bs, "bad bases handle %llu (never registered, or already released)", (unsigned long long)handle);
        bs->refs += 1;
        return ZKG_OK;
    }
    // order a later release behind everything `ctx` has enqueued so far (asynchronous entry points)
    int32_t mark_use(zkg_ctx* ctx) {
        std::lock_guard<std::mutex> lk(g_bases_mu);
        cudaEvent_t ev = nullptr;
        for (auto& e : bs->last_use) if (e.first == ctx) ev = e.second;
        if (!ev) {
            ZKG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            bs->last_use.emplace_back(ctx, ev);
        }
        ZKG_CUDA(cudaEventRecord(ev, ctx->stream));
        return ZKG_OK;
    }
    ~BaseRef() {
        if (!bs) return;
        std::lock_guard<std::mutex> lk(g_bases_mu);
        bs->refs -= 1;
        if (bs->refs == 0) g_bases_cv.notify_all();
    }
};

// MSM against a set registered over several GPUs: every shard runs the prepared pipeline on its own scalar range (one
// host thread per device for the chunked scalar upload), leaves an XYZZ partial, and the first shard's device adds them.
static int32_t msm_registered_sharded(const BaseSet& bs, const uint64_t* scalars, uint64_t* out_xyz) {
    const int K = (int)bs.shard_handles.size();
    const size_t part_b = bs.group == 1 ? 128 : 256;
    std::vector<BaseRef> refs(K);
    std::vector<PooledCtx> pcs(K);
    for (int k = 0; k < K; ++k) {
        ZKG_TRY(refs[k].acquire(bs.shard_handles[k]));
        ZKG_TRY(pcs[k].acquire(refs[k].bs->device));
    }
    std::vector<int32_t> rc(K, ZKG_OK);
    std::vector<std::string> msg(K);
    std::vector<void*> d_part(K, nullptr);
    std::vector<cudaEvent_t> done(K, nullptr);
    auto work = [&](int k) {
        zkg_ctx* ctx = pcs[k].ctx;
        const BaseSet& sh = *refs[k].bs;
        DeviceGuard dg(ctx->device);
        const size_t cnt = sh.n, sc_bytes = align_up(cnt * 32, 256);
        rc[k] = ctx->io.reserve(sc_bytes + 512 + 16 * 256 + 256);
        if (rc[k] == ZKG_OK) {
            d_part[k] = (uint8_t*)ctx->io.p + sc_bytes;
            rc[k] = bs.group == 1 ? msm_run_prepared_host_g1(ctx, sh.d_table, sh.c, scalars + bs.shard_lo[k] * 4, cnt, d_part[k], 1)
                                  : msm_run_prepared_host_g2(ctx, sh.d_table, sh.c, scalars + bs.shard_lo[k] * 4, cnt, d_part[k], 1);
        }
        if (rc[k] == ZKG_OK && (cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming) != cudaSuccess ||
                                cudaEventRecord(done[k], ctx->stream) != cudaSuccess)) {
            set_error("msm_registered (sharded): event on device %d failed", ctx->device);
            rc[k] = ZKG_ERR_CUDA;
        }
        if (rc[k] != ZKG_OK) msg[k] = zkg_last_error();
    };
    std::vector<std::thread> th;
    for (int k = 1; k < K; ++k) th.emplace_back(work, k);
    work(0);
    for (auto& t : th) t.join();
    int32_t first = ZKG_OK;
    for (int k = 0; k < K && first == ZKG_OK; ++k)
        if (rc[k] != ZKG_OK) { first = rc[k]; set_error("%s", msg[k].c_str()); }
    zkg_ctx* c0 = pcs[0].ctx;
    {
        DeviceGuard dg(c0->device);
        if (first == ZKG_OK) {
            uint8_t* gather = (uint8_t*)d_part[0] + 512;                 // 16 partial slots + the result, inside device 0's staging area
            uint8_t* d_out = gather + 16 * 256;
            for (int k = 0; k < K && first == ZKG_OK; ++k) {
                cudaError_t e = cudaStreamWaitEvent(c0->stream, done[k], 0);
                if (e == cudaSuccess)
                    e = k == 0 ? cudaMemcpyAsync(gather, d_part[0], part_b, cudaMemcpyDeviceToDevice, c0->stream)
                               : cudaMemcpyPeerAsync(gather + k * part_b, c0->device, d_part[k], pcs[k].ctx->device, part_b, c0->stream);
                if (e != cudaSuccess) { set_error("msm_registered (sharded): gathering shard %d failed: %s", k, cudaGetErrorString(e)); first = ZKG_ERR_CUDA; }
            }
            if (first == ZKG_OK) first = bs.group == 1 ? combine_g1(c0, (const uint64_t*)gather, K, (uint64_t*)d_out)
                                                       : combine_g2(c0, (const uint64_t*)gather, K, (uint64_t*)d_out);
            if (first == ZKG_OK) first = copy_d2h(out_xyz, d_out, bs.group == 1 ? 96 : 192, c0->stream);
        }
    }
    for (int k = 0; k < K; ++k) {
        DeviceGuard dgk(pcs[k].ctx->device);
        cudaStreamSynchronize(pcs[k].ctx->stream);
        if (done[k]) cudaEventDestroy(done[k]);
    }
    return first;
}

static int32_t base_set_create_fwd(zkg_ctx* ctx, int32_t group, const void* d_packed, size_t n, uint64_t* handle);

// ------------------------------------------------------------------------------------------------------------
// Opt-in (ZKG_AUTO_REGISTER=1): transparent registration for the UNCHANGED caller.  `d_msm(bases, scalars)`
// (dist-primitives/src/dmsm/mod.rs:59-73) hands over the same static CRS share on every proof
// (groth16/src/proving_key.rs:15-45) but has no place to keep a handle.  With the switch on, the second call that shows the
// same (device, pointer, stride, length) registers the bases (one-time table preparation) and keeps a packed device copy;
// later calls still ship the bases -- every byte the caller passed crosses PCIe, on a second stream -- but only to be
// COMPARED on the device with that copy, while the MSM itself runs against the prepared table with the caller's scalars.
// A single differing byte sets a flag; the call then drops the entry and recomputes through the ordinary path, so the result is
// the reference's for any input (tests/test_gpu_round2.py::test_auto_register_*).
// ------------------------------------------------------------------------------------------------------------
template <int WORDS16>      // 16-byte words per packed base: 4 (G1) or 8 (G2)
__global__ void k_verify_bases(const uint8_t* __restrict__ ark, size_t stride, size_t n, const uint4* __restrict__ packed, int* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = ark + i * stride;
    const bool inf = p[WORDS16 * 16] != 0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < WORDS16; ++k) {
        uint4 want = packed[i * WORDS16 + k];
        uint4 got = make_uint4(0, 0, 0, 0);
        if (!inf) {
            if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
                const uint2* q = reinterpret_cast<const uint2*>(p + 16 * k);
                uint2 a = q[0], b = q[1];
                got = make_uint4(a.x, a.y, b.x, b.y);
            } else {
                uint32_t w[4];
                for (int j = 0; j < 4; ++j) {
                    const uint8_t* b = p + 16 * k + 4 * j;
                    w[j] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
                }
                got = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        bad |= got.x != want.x || got.y != want.y || got.z != want.z || got.w != want.w;
    }
    if (bad) atomicOr(flag, 1);
}

struct AutoEntry {
    int device = 0, group = 0;
    const void* ptr = nullptr;
    size_t stride = 0, n = 0;
    int seen = 0;
    bool busy = false;
    uint64_t handle = 0;
    size_t bytes = 0;                // device memory held (table + packed copy + landing area)
    uint64_t last_use = 0;
    uint8_t* d_packed = nullptr;     // packed copy of the registered bases (what the table was built from)
    uint8_t* d_ark = nullptr;        // landing area of the bases shipped by each call + the flag
    cudaStream_t vstream = nullptr;
    int* h_flag = nullptr;           // pinned
};
static std::mutex g_auto_mu;
static std::vector<AutoEntry*> g_auto;
static uint64_t g_auto_clock = 0;
static size_t auto_budget_bytes() {                  // ZKG_AUTO_REGISTER_MAX_MB (default 32 GiB of the 180 GB): least recently used sets go first
    const char* e = getenv("ZKG_AUTO_REGISTER_MAX_MB");
    long mb = e ? atol(e) : 32768;
    return (size_t)(mb < 0 ? 0 : mb) << 20;
}

static bool auto_register_enabled() {
    const char* e = getenv("ZKG_AUTO_REGISTER");
    return e && e[0] == '1';
}
static void auto_entry_free(AutoEntry* e) {          // caller holds no lock; the entry is already unlinked
    DeviceGuard dg(e->device);
    if (e->handle) zkg_bases_release(e->handle);
    if (e->d_packed) cudaFree(e->d_packed);
    if (e->d_ark) cudaFree(e->d_ark);
    if (e->vstream) cudaStreamDestroy(e->vstream);
    if (e->h_flag) cudaFreeHost(e->h_flag);
    delete e;
}
// Registers the bases of `e` (table + packed copy + landing area).  Called without the lock, with e->busy set.
static int32_t auto_entry_prepare(AutoEntry* e, const void* bases) {
    DeviceGuard dg(e->device);
    const size_t pk = e->n * packed_bytes(e->group), ark = align_up(e->n * e->stride, 256);
    PooledCtx pc;
    ZKG_TRY(pc.acquire(e->device));
    zkg_ctx* ctx = pc.ctx;
    ZKG_TRY(ctx->io.reserve(ark + pk + 256));
    uint8_t* d_ark = (uint8_t*)ctx->io.p;
    uint8_t* d_pk = d_ark + ark;
    ZKG_TRY(copy_h2d(d_ark, bases, e->n * e->stride, ctx->stream));
    ZKG_TRY(e->group == 1 ? pack_bases_g1(ctx, d_ark, e->stride, e->n, d_pk) : pack_bases_g2(ctx, d_ark, e->stride, e->n, d_pk));
    ZKG_CUDA(cudaMalloc(&e->d_packed, pk));
    ZKG_CUDA(cudaMemcpyAsync(e->d_packed, d_pk, pk, cudaMemcpyDeviceToDevice, ctx->stream));
    ZKG_TRY(base_set_create_fwd(ctx, e->group, d_pk, e->n, &e->handle));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    ZKG_CUDA(cudaMalloc(&e->d_ark, ark + 256));
    int prio_lo = 0, prio_hi = 0;                        // the comparison kernel takes the first slot an accumulation block frees
    ZKG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    ZKG_CUDA(cudaStreamCreateWithPriority(&e->vstream, cudaStreamNonBlocking, prio_hi));
    ZKG_CUDA(cudaHostAlloc((void**)&e->h_flag, sizeof(int), cudaHostAllocPortable));
    return ZKG_OK;
}
void auto_register_clear() {                         // zkg_shutdown: drop every transparent registration that is not in use
    std::vector<AutoEntry*> gone;
    {
        std::lock_guard<std::mutex> lk(g_auto_mu);
        for (size_t i = 0; i < g_auto.size();) {
            if (g_auto[i]->busy) { ++i; continue; }
            gone.push_back(g_auto[i]);
            g_auto.erase(g_auto.begin() + i);
        }
    }
    for (AutoEntry* e : gone) auto_entry_free(e);
}
// Returns ZKG_OK with *handled = true when the call was served from a verified registration; *handled = false means
// "run the ordinary path" (first sightings, a busy entry, a failed verification, any set-up error).
static int32_t msm_host_auto(int group, int device, const void* bases, size_t stride, size_t n, const uint64_t* scalars, uint64_t* out_xyz,
                             bool* handled) {
    *handled = false;
    AutoEntry* e = nullptr;
    bool prepare = false, give_up = false;
    std::vector<AutoEntry*> evicted;
    {
        std::lock_guard<std::mutex> lk(g_auto_mu);
        for (AutoEntry* x : g_auto)
            if (x->device == device && x->group == group && x->ptr == bases && x->stride == stride && x->n == n) e = x;
        if (!e) {
            if (g_auto.size() >= 16) {
                // a handful of CRS shares per prover: never grow without bound; the oldest pointer that was seen only once goes
                size_t victim = g_auto.size();
                for (size_t i = 0; i < g_auto.size(); ++i)
                    if (!g_auto[i]->handle && !g_auto[i]->busy && (victim == g_auto.size() || g_auto[i]->last_use < g_auto[victim]->last_use)) victim = i;
                if (victim == g_auto.size()) return ZKG_OK;
                delete g_auto[victim];
                g_auto.erase(g_auto.begin() + victim);
            }
            e = new AutoEntry();
            e->device = device; e->group = group; e->ptr = bases; e->stride = stride; e->n = n; e->seen = 1;
            e->last_use = ++g_auto_clock;
            g_auto.push_back(e);
            return ZKG_OK;
        }
        if (e->busy) return ZKG_OK;
        if (!e->handle) {
            e->seen += 1;
            e->last_use = ++g_auto_clock;
            if (e->seen < 2) return ZKG_OK;
            // device memory the registration will hold; make room by dropping the least recently used idle sets, or give up
            int c = msm_pick_c_merged_host(n);
            while (c < 23 && n * (size_t)(254 / c + 1) >= ((size_t)1 << 31)) ++c;
            e->bytes = n * (size_t)(254 / c + 1) * packed_bytes(group) + n * packed_bytes(group) + align_up(n * stride, 256) + 256;
            const size_t budget = auto_budget_bytes();
            if (e->bytes > budget) return ZKG_OK;
            for (;;) {
                size_t held = 0;
                AutoEntry* lru = nullptr;
                for (AutoEntry* x : g_auto) {
                    if (!x->handle) continue;
                    held += x->bytes;
                    if (!x->busy && (!lru || x->last_use < lru->last_use)) lru = x;
                }
                if (held + e->bytes <= budget) break;
                if (!lru) { give_up = true; break; }                      // everything that could go is in use
                for (size_t i = 0; i < g_auto.size(); ++i) if (g_auto[i] == lru) { g_auto.erase(g_auto.begin() + i); break; }
                evicted.push_back(lru);
            }
            prepare = !give_up;
        }
        if (!give_up) {
            e->busy = true;
            e->last_use = ++g_auto_clock;
        }
    }
    for (AutoEntry* x : evicted) auto_entry_free(x);
    if (give_up) return ZKG_OK;
    auto unbusy = [&] { std::lock_guard<std::mutex> lk(g_auto_mu); e->busy = false; };
    auto drop = [&] {
        { std::lock_guard<std::mutex> lk(g_auto_mu); for (size_t i = 0; i < g_auto.size(); ++i) if (g_auto[i] == e) { g_auto.erase(g_auto.begin() + i); break; } }
        auto_entry_free(e);
    };
    if (prepare) {
        // this call pays the one-time preparation and then runs the ordinary path; the next one is served from the table
        if (auto_entry_prepare(e, bases) != ZKG_OK) {
            cudaGetLastError();                              // e.g. out of memory: not this call's error -- the ordinary path answers
            drop();
            return ZKG_OK;
        }
        unbusy();
        return ZKG_OK;
    }
    int32_t rc = ZKG_OK;
    bool mismatch = false;
    {
        DeviceGuard dg(device);
        // (1) the MSM against the prepared table, with the caller's scalars (enqueued first: with pageable scalars the
        //     staging of this thread is then interleaved with the accumulations, not with the verification copies)
        BaseRef ref;
        rc = ref.acquire(e->handle);
        PooledCtx pc;
        if (rc == ZKG_OK) rc = pc.acquire(device);
        zkg_ctx* ctx = pc.ctx;
        void* d_out = nullptr;
        if (rc == ZKG_OK) {
            const size_t sc_bytes = align_up(n * 32, 256);
            rc = ctx->io.reserve(sc_bytes + 512);
            d_out = (uint8_t*)ctx->io.p + sc_bytes;
        }
        if (rc == ZKG_OK)
            rc = group == 1 ? msm_run_prepared_host_g1(ctx, ref.bs->d_table, ref.bs->c, scalars, n, d_out, 0)
                            : msm_run_prepared_host_g2(ctx, ref.bs->d_table, ref.bs->c, scalars, n, d_out, 0);
        // (2) the bases the caller passed: shipped in full on the entry's own stream and compared with the registered copy
        if (rc == ZKG_OK) {
            const size_t ark = align_up(n * stride, 256);
            int* d_flag = (int*)(e->d_ark + ark);
            cudaError_t ce = cudaMemsetAsync(d_flag, 0, sizeof(int), e->vstream);
            // every copy first, ONE comparison kernel after them: a kernel queued between two copies would wait for a free
            // slot under the resident accumulation blocks and hold back the copies behind it (measured: 12.9 ms instead of 9.7)
            const int K = 8;
            for (int j = 0; j < K && rc == ZKG_OK; ++j) {
                const size_t lo = n * (size_t)j / K, hi = n * (size_t)(j + 1) / K;
                if (lo < hi) rc = copy_h2d(e->d_ark + lo * stride, (const uint8_t*)bases + lo * stride, (hi - lo) * stride, e->vstream);
            }
            if (ce == cudaSuccess && rc == ZKG_OK) {
                const unsigned blocks = (unsigned)((n + 255) / 256);
                if (group == 1) k_verify_bases<4><<<blocks, 256, 0, e->vstream>>>(e->d_ark, stride, n, (const uint4*)e->d_packed, d_flag);
                else k_verify_bases<8><<<blocks, 256, 0, e->vstream>>>(e->d_ark, stride, n, (const uint4*)e->d_packed, d_flag);
                ce = cudaGetLastError();
            }
            if (ce == cudaSuccess && rc == ZKG_OK) ce = cudaMemcpyAsync(e->h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->vstream);
            // the result comes back only now: a copy into PAGEABLE caller memory blocks this thread until the MSM has finished,
            // which must not hold back the verification copies above
            if (ce == cudaSuccess && rc == ZKG_OK) rc = copy_d2h(out_xyz, d_out, group == 1 ? 96 : 192, ctx->stream);
            if (ce == cudaSuccess && rc == ZKG_OK) ce = cudaStreamSynchronize(e->vstream);
            if (ce != cudaSuccess && rc == ZKG_OK) { set_error("auto-register verification failed: %s", cudaGetErrorString(ce)); rc = ZKG_ERR_CUDA; }
            if (rc == ZKG_OK) mismatch = *e->h_flag != 0;
        }
        if (ctx) {
            cudaError_t se = cudaStreamSynchronize(ctx->stream);
            if (rc == ZKG_OK && se != cudaSuccess) { set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(se)); rc = ZKG_ERR_CUDA; }
        }
    }
    if (rc != ZKG_OK || mismatch) {
        // the memory behind the pointer changed (or something failed): forget the entry and let the ordinary path answer
        cudaGetLastError();
        drop();
        return ZKG_OK;
    }
    unbusy();
    *handled = true;
    return ZKG_OK;
}

}  // namespace zkg

using namespace zkg;

extern "C" {

int32_t zkg_msm_bn254_g1(int32_t device, const void* bases, size_t base_stride, size_t n_bases, const uint64_t* scalars,
                         size_t n_scalars, uint64_t out_xyz[12]) {
    if (n_bases == n_scalars && n_bases >= ((size_t)1 << 16) && bases && scalars && out_xyz && auto_register_enabled()) {
        bool handled = false;
        msm_host_auto(1, device, bases, base_stride, n_bases, scalars, out_xyz, &handled);
        if (handled) return ZKG_OK;
    }
    return msm_host_g1(device, bases, base_stride, n_bases, scalars, n_scalars, out_xyz);
}
int32_t zkg_msm_bn254_g2(int32_t device, const void* bases, size_t base_stride, size_t n_bases, const uint64_t* scalars,
                         size_t n_scalars, uint64_t out_xyz[24]) {
    if (n_bases == n_scalars && n_bases >= ((size_t)1 << 16) && bases && scalars && out_xyz && auto_register_enabled()) {
        bool handled = false;
        msm_host_auto(2, device, bases, base_stride, n_bases, scalars, out_xyz, &handled);
        if (handled) return ZKG_OK;
    }
    return msm_host_g2(device, bases, base_stride, n_bases, scalars, n_scalars, out_xyz);
}

int32_t zkg_msm_bn254_g1_sharded(const int32_t* devices, int32_t n_devices, const void* bases, size_t base_stride, size_t n_bases,
                                 const uint64_t* scalars, size_t n_scalars, uint64_t out_xyz[12]) {
    return msm_host_sharded_g1(devices, n_devices, bases, base_stride, n_bases, scalars, n_scalars, out_xyz);
}
int32_t zkg_msm_bn254_g2_sharded(const int32_t* devices, int32_t n_devices, const void* bases, size_t base_stride, size_t n_bases,
                                 const uint64_t* scalars, size_t n_scalars, uint64_t out_xyz[24]) {
    return msm_host_sharded_g2(devices, n_devices, bases, base_stride, n_bases, scalars, n_scalars, out_xyz);
}

int32_t zkg_pack_bases_dev(zkg_ctx* ctx, int32_t group, const void* d_bases_ark, size_t base_stride, size_t n,
                           void* d_bases_packed) {
    ZKG_REQUIRE(ctx && (group == 1 || group == 2), "pack_bases: bad ctx/group");
    DeviceGuard dg(ctx->device);
    return group == 1 ? pack_bases_g1(ctx, d_bases_ark, base_stride, n, d_bases_packed)
                      : pack_bases_g2(ctx, d_bases_ark, base_stride, n, d_bases_packed);
}
int32_t zkg_msm_bn254_g1_dev(zkg_ctx* ctx, const void* d_bases_packed, const uint64_t* d_scalars, size_t n,
                             uint64_t* d_out_xyz) {
    ZKG_REQUIRE(ctx, "msm: ctx is NULL");
    DeviceGuard dg(ctx->device);
    return msm_run_g1(ctx, d_bases_packed, d_scalars, n, d_out_xyz, 0);
}
int32_t zkg_msm_bn254_g2_dev(zkg_ctx* ctx, const void* d_bases_packed, const uint64_t* d_scalars, size_t n,
                             uint64_t* d_out_xyz) {
    ZKG_REQUIRE(ctx, "msm: ctx is NULL");
    DeviceGuard dg(ctx->device);
    return msm_run_g2(ctx, d_bases_packed, d_scalars, n, d_out_xyz, 0);
}
int32_t zkg_msm_bn254_partial_dev(zkg_ctx* ctx, int32_t group, const void* d_bases_packed, const uint64_t* d_scalars,
                                  size_t n, uint64_t* d_out_xyzz) {
    ZKG_REQUIRE(ctx && (group == 1 || group == 2), "msm_partial: bad ctx/group");
    DeviceGuard dg(ctx->device);
    return group == 1 ? msm_run_g1(ctx, d_bases_packed, d_scalars, n, d_out_xyzz, 1)
                      : msm_run_g2(ctx, d_bases_packed, d_scalars, n, d_out_xyzz, 1);
}
int32_t zkg_msm_combine_dev(zkg_ctx* ctx, int32_t group, const uint64_t* d_partials_xyzz, size_t n_partials,
                            uint64_t* d_out_xyz) {
    ZKG_REQUIRE(ctx && (group == 1 || group == 2), "msm_combine: bad ctx/group");
    DeviceGuard dg(ctx->device);
    return group == 1 ? combine_g1(ctx, d_partials_xyzz, n_partials, d_out_xyz)
                      : combine_g2(ctx, d_partials_xyzz, n_partials, d_out_xyz);
}
int32_t zkg_fixed_base_dev(zkg_ctx* ctx, int32_t group, const uint64_t* d_scalars, size_t n, void* d_bases_packed) {
    ZKG_REQUIRE(ctx && (group == 1 || group == 2), "fixed_base: bad ctx/group");
    DeviceGuard dg(ctx->device);
    return group == 1 ? fixed_base_g1(ctx, d_scalars, n, d_bases_packed) : fixed_base_g2(ctx, d_scalars, n, d_bases_packed);
}

static int32_t base_set_create(zkg_ctx* ctx, int32_t group, const void* d_packed, size_t n, uint64_t* handle) {
    BaseSet* bs = new BaseSet();
    bs->device = ctx->device; bs->group = group; bs->n = n;
    if (n) {
        bs->c = msm_pick_c_merged_host(n);
        int env_c = getenv("ZKG_MSM_PREP_C") ? atoi(getenv("ZKG_MSM_PREP_C")) : 0;
        if (env_c >= 4 && env_c <= 23) bs->c = env_c;
        // a sorted entry is (window * n + point) with the sign in bit 31 (msm_plan): shrink the table (larger windows)
        // until it is addressable, and refuse what cannot be made so, BEFORE allocating gigabytes
        while (bs->c < 23 && n * (size_t)(254 / bs->c + 1) >= ((size_t)1 << 31)) bs->c += 1;
        const int W = 254 / bs->c + 1;
        if (n * (size_t)W >= ((size_t)1 << 31)) {
            delete bs;
            set_error("bases_register: %zu points x %d window shifts exceeds the 2^31-1 table entries an MSM can address", n, W);
            return ZKG_ERR_BAD_ARG;
        }
        size_t bytes = n * (size_t)W * packed_bytes(group);
        cudaError_t e = cudaMalloc(&bs->d_table, bytes);
        if (e != cudaSuccess) {
            delete bs;
            set_error("bases_register: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            return ZKG_ERR_OOM;
        }
        int32_t rc = group == 1 ? prepare_g1(ctx, d_packed, n, bs->c, bs->d_table) : prepare_g2(ctx, d_packed, n, bs->c, bs->d_table);
        if (rc != ZKG_OK) { cudaFree(bs->d_table); delete bs; return rc; }
    }
    std::lock_guard<std::mutex> lk(g_bases_mu);
    size_t slot = g_bases.size();
    for (size_t i = 0; i < g_bases.size(); ++i) if (!g_bases[i]) { slot = i; break; }       // reuse a released slot
    if (slot == g_bases.size()) { g_bases.push_back(nullptr); g_bases_gen.push_back(0); }
    bs->generation = ++g_bases_gen[slot];
    g_bases[slot] = bs;
    *handle = ((uint64_t)bs->generation << 32) | (uint64_t)(slot + 1);
    return ZKG_OK;
}

}  // extern "C"
namespace zkg {
static int32_t base_set_create_fwd(zkg_ctx* ctx, int32_t group, const void* d_packed, size_t n, uint64_t* handle) {
    return base_set_create(ctx, group, d_packed, n, handle);
}
}  // namespace zkg
extern "C" {

int32_t zkg_crs_det_pack_bn254(int32_t device, int32_t group, const void* bases, size_t base_stride, size_t n, uint32_t l,
                               void* const* out_by_party, size_t out_stride) {
    ZKG_REQUIRE((group == 1 || group == 2) && out_by_party && (n == 0 || bases), "crs_det_pack: bad argument");
    ZKG_REQUIRE(l == 2 || l == 4 || l == 8, "packing factor l = %u unsupported (2, 4, 8)", l);
    host::PssMatrices pm;
    host::pss_matrices(l, &pm);
    // det_pack pads with zeros, so only the first l columns of the n x (l+t) pack matrix matter; the
    // kernel wants canonical (non-Montgomery) scalars
    std::vector<uint32_t> scal((size_t)pm.n * l * 8);
    host::HFr raw_one = host::h_zero();
    raw_one.v[0] = 1;
    for (uint32_t i = 0; i < pm.n; ++i)
        for (uint32_t k = 0; k < l; ++k) {
            host::HFr c = host::h_mul(pm.pack[(size_t)i * (pm.l + pm.t) + k], raw_one);
            memcpy(&scal[((size_t)i * l + k) * 8], c.v, 32);
        }
    return group == 1 ? crs_det_pack_g1(device, bases, base_stride, n, (int)l, (int)pm.n, scal.data(), out_by_party, out_stride)
                      : crs_det_pack_g2(device, bases, base_stride, n, (int)l, (int)pm.n, scal.data(), out_by_party, out_stride);
}

int32_t zkg_bases_register(int32_t device, int32_t group, const void* bases, size_t base_stride, size_t n,
                           uint64_t* handle) {
    ZKG_REQUIRE(handle && (group == 1 || group == 2) && (n == 0 || bases), "bases_register: bad argument");
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t ark_bytes = align_up(n * base_stride, 256);
    ZKG_TRY(ctx->io.reserve(ark_bytes + n * packed_bytes(group) + 256));
    uint8_t* d_ark = (uint8_t*)ctx->io.p;
    uint8_t* d_pk = d_ark + ark_bytes;
    if (n) {
        ZKG_TRY(copy_h2d(d_ark, bases, n * base_stride, ctx->stream));
        ZKG_TRY(group == 1 ? pack_bases_g1(ctx, d_ark, base_stride, n, d_pk) : pack_bases_g2(ctx, d_ark, base_stride, n, d_pk));
    }
    ZKG_TRY(base_set_create(ctx, group, d_pk, n, handle));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_bases_register_sharded(const int32_t* devices, int32_t n_devices, int32_t group, const void* bases, size_t base_stride,
                                   size_t n, uint64_t* handle) {
    ZKG_REQUIRE(devices && n_devices >= 1 && n_devices <= 16 && handle && (group == 1 || group == 2) && (n == 0 || bases),
                "bases_register_sharded: bad argument");
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < a; ++b) ZKG_REQUIRE(devices[a] != devices[b], "bases_register_sharded: device %d listed twice", devices[a]);
    if (n_devices == 1 || n < (size_t)n_devices) return zkg_bases_register(devices[0], group, bases, base_stride, n, handle);
    BaseSet* top = new BaseSet();
    top->device = devices[0]; top->group = group; top->n = n;
    int32_t rc = ZKG_OK;
    for (int d = 0; d < n_devices && rc == ZKG_OK; ++d) {
        const size_t lo = n * (size_t)d / (size_t)n_devices, hi = n * (size_t)(d + 1) / (size_t)n_devices;
        uint64_t h = 0;
        rc = zkg_bases_register(devices[d], group, (const uint8_t*)bases + lo * base_stride, base_stride, hi - lo, &h);
        if (rc == ZKG_OK) { top->shard_handles.push_back(h); top->shard_lo.push_back(lo); }
    }
    if (rc != ZKG_OK) {
        std::string keep = zkg_last_error();
        for (uint64_t h : top->shard_handles) zkg_bases_release(h);
        delete top;
        set_error("%s", keep.c_str());
        return rc;
    }
    top->shard_lo.push_back(n);
    std::lock_guard<std::mutex> lk(g_bases_mu);
    size_t slot = g_bases.size();
    for (size_t i = 0; i < g_bases.size(); ++i) if (!g_bases[i]) { slot = i; break; }
    if (slot == g_bases.size()) { g_bases.push_back(nullptr); g_bases_gen.push_back(0); }
    top->generation = ++g_bases_gen[slot];
    g_bases[slot] = top;
    *handle = ((uint64_t)top->generation << 32) | (uint64_t)(slot + 1);
    return ZKG_OK;
}

int32_t zkg_bases_register_dev(zkg_ctx* ctx, int32_t group, const void* d_bases_packed, size_t n, uint64_t* handle) {
    ZKG_REQUIRE(ctx && handle && (group == 1 || group == 2) && (n == 0 || d_bases_packed), "bases_register_dev: bad argument");
    DeviceGuard dg(ctx->device);
    return base_set_create(ctx, group, d_bases_packed, n, handle);
}

int32_t zkg_bases_release(uint64_t handle) {
    BaseSet* bs = nullptr;
    {
        std::unique_lock<std::mutex> lk(g_bases_mu);
        bs = base_set_lookup_locked(handle);
        ZKG_REQUIRE(bs, "bases_release: bad handle %llu (never registered, or already released)", (unsigned long long)handle);
        bs->released = true;                                  // no new references from here on
        g_bases_cv.wait(lk, [&] { return bs->refs == 0; });   // calls that already hold the table finish (or record their event) first
        g_bases[(handle & 0xffffffffu) - 1] = nullptr;        // the slot may be reused under a new generation
    }
    DeviceGuard dg(bs->device);
    for (auto& e : bs->last_use) {                            // asynchronous users: wait for the work they enqueued
        cudaEventSynchronize(e.second);
        cudaEventDestroy(e.second);
    }
    if (bs->d_table) cudaFree(bs->d_table);
    int32_t rc = ZKG_OK;
    for (uint64_t h : bs->shard_handles) { int32_t r = zkg_bases_release(h); if (r != ZKG_OK) rc = r; }
    delete bs;
    return rc;
}

int32_t zkg_msm_bn254_registered_dev(zkg_ctx* ctx, uint64_t handle, const uint64_t* d_scalars, size_t n_scalars,
                                     uint64_t* d_out, int32_t partial) {
    ZKG_REQUIRE(ctx && d_out, "msm_registered_dev: NULL argument");
    BaseRef ref;
    ZKG_TRY(ref.acquire(handle));
    const BaseSet& bs = *ref.bs;
    ZKG_REQUIRE(bs.shard_handles.empty(), "msm_registered_dev: the handle is sharded over several GPUs; use zkg_msm_bn254_registered");
    ZKG_REQUIRE(bs.device == ctx->device, "msm_registered_dev: bases live on device %d, context on %d", bs.device, ctx->device);
    if (bs.n != n_scalars) {
        set_error("msm: bases.len() = %zu, scalars.len() = %zu", bs.n, n_scalars);
        return ZKG_ERR_LEN_MISMATCH;
    }
    DeviceGuard dg(ctx->device);
    int32_t rc = bs.group == 1 ? msm_run_prepared_g1(ctx, bs.d_table, bs.c, d_scalars, n_scalars, d_out, partial ? 1 : 0)
                               : msm_run_prepared_g2(ctx, bs.d_table, bs.c, d_scalars, n_scalars, d_out, partial ? 1 : 0);
    // whatever was enqueued (even by a call that failed half way) must finish before the table may be freed
    int32_t rc2 = ref.mark_use(ctx);
    return rc != ZKG_OK ? rc : rc2;
}

int32_t zkg_msm_bn254_registered(uint64_t handle, const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz) {
    BaseRef ref;
    ZKG_TRY(ref.acquire(handle));
    const BaseSet& bs = *ref.bs;
    if (bs.n != n_scalars) {
        set_error("msm: bases.len() = %zu, scalars.len() = %zu", bs.n, n_scalars);
        return ZKG_ERR_LEN_MISMATCH;
    }
    ZKG_REQUIRE(out_xyz && (n_scalars == 0 || scalars), "msm_registered: NULL argument");
    if (!bs.shard_handles.empty()) return msm_registered_sharded(bs, scalars, out_xyz);
    PooledCtx pc;
    ZKG_TRY(pc.acquire(bs.device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t sc_bytes = align_up(n_scalars * 32, 256);
    ZKG_TRY(ctx->io.reserve(sc_bytes + 512));
    void* d_out = (uint8_t*)ctx->io.p + sc_bytes;
    int32_t rc = bs.group == 1 ? msm_run_prepared_host_g1(ctx, bs.d_table, bs.c, scalars, n_scalars, d_out, 0)
                               : msm_run_prepared_host_g2(ctx, bs.d_table, bs.c, scalars, n_scalars, d_out, 0);
    if (rc == ZKG_OK) rc = copy_d2h(out_xyz, d_out, bs.group == 1 ? 96 : 192, ctx->stream);
    // blocking call: the reference is held until the stream has drained, also on the error paths
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    if (rc == ZKG_OK && se != cudaSuccess) { set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(se)); rc = ZKG_ERR_CUDA; }
    return rc;
}


// King side of d_msm (dmsm/mod.rs:85-87): unpack_missing_shares over n_recv group elements, then the sum.
static int32_t pss_unpack2_group(int32_t device, int group, uint32_t l, const uint64_t* shares_xyz, const uint32_t* parties,
                                 uint32_t n_recv, uint64_t* out_unpacked, uint64_t* out_sum) {
    ZKG_REQUIRE(shares_xyz && (out_unpacked || out_sum), "pss_unpack2 over group elements: NULL argument");
    ZKG_REQUIRE(l == 2 || l == 4 || l == 8, "packing factor l = %u unsupported (2, 4, 8)", l);
    host::PssMatrices pm;
    host::pss_matrices(l, &pm);
    std::vector<host::HFr> lag;
    const std::vector<host::HFr>* M = &pm.unpack2;                   // l x n_recv, row-major
    if (n_recv != pm.n) {
        ZKG_REQUIRE(parties, "parties list required when shares are missing");
        if (!host::pss_lagrange_matrix(l, parties, n_recv, &lag)) {
            set_error("not enough shares to reconstruct: got %u of n = %u (need > %u distinct parties)", n_recv, pm.n,
                      2 * (pm.t + pm.l - 1));
            return ZKG_ERR_BAD_ARG;
        }
        M = &lag;
    }
    host::HFr raw_one = host::h_zero();
    raw_one.v[0] = 1;
    // the sum alone needs one row: the column sums of the matrix (sum_i sum_j M_ij S_j = sum_j (sum_i M_ij) S_j)
    const uint32_t rows = out_unpacked ? l : 1;
    std::vector<uint32_t> scal((size_t)rows * n_recv * 8);
    for (uint32_t j = 0; j < n_recv; ++j) {
        host::HFr col = host::h_zero();
        for (uint32_t i = 0; i < l; ++i) {
            const host::HFr& e = (*M)[(size_t)i * n_recv + j];
            col = host::h_add(col, e);
            if (out_unpacked) { host::HFr c = host::h_mul(e, raw_one); memcpy(&scal[((size_t)i * n_recv + j) * 8], c.v, 32); }
        }
        if (!out_unpacked) { host::HFr c = host::h_mul(col, raw_one); memcpy(&scal[(size_t)j * 8], c.v, 32); }
    }
    return group == 1 ? group_unpack_g1(device, shares_xyz, n_recv, scal.data(), rows, out_unpacked, out_sum)
                      : group_unpack_g2(device, shares_xyz, n_recv, scal.data(), rows, out_unpacked, out_sum);
}

int32_t zkg_pss_unpack2_bn254_g1(int32_t device, uint32_t l, const uint64_t* shares_xyz, const uint32_t* parties, uint32_t n_recv,
                                 uint64_t* out_unpacked_xyz, uint64_t* out_sum_xyz) {
    return pss_unpack2_group(device, 1, l, shares_xyz, parties, n_recv, out_unpacked_xyz, out_sum_xyz);
}
int32_t zkg_pss_unpack2_bn254_g2(int32_t device, uint32_t l, const uint64_t* shares_xyz, const uint32_t* parties, uint32_t n_recv,
                                 uint64_t* out_unpacked_xyz, uint64_t* out_sum_xyz) {
    return pss_unpack2_group(device, 2, l, shares_xyz, parties, n_recv, out_unpacked_xyz, out_sum_xyz);
}

int32_t zkg_g1_from_wire_bn254(int32_t device, const void* wire, uint64_t* points_xyz, size_t n) { return point_wire_g1(device, 0, wire, points_xyz, n); }
int32_t zkg_g1_to_wire_bn254(int32_t device, const uint64_t* points_xyz, void* wire, size_t n) { return point_wire_g1(device, 1, points_xyz, wire, n); }
int32_t zkg_g2_from_wire_bn254(int32_t device, const void* wire, uint64_t* points_xyz, size_t n) { return point_wire_g2(device, 0, wire, points_xyz, n); }
int32_t zkg_g2_to_wire_bn254(int32_t device, const uint64_t* points_xyz, void* wire, size_t n) { return point_wire_g2(device, 1, points_xyz, wire, n); }

}  // extern "C"
