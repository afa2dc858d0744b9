// api.cu -- context management, error reporting and small element-wise entry points of the C ABI.
#include "common.cuh"
#include "ec.cuh"

namespace zkg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static std::mutex g_pool_mu;
static std::vector<zkg_ctx*> g_pool;     // idle pooled contexts (any device)

static int32_t ctx_new(int device, void* stream, zkg_ctx** out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libzksaas_gpu has no CPU fallback", cudaGetErrorString(e));
        return ZKG_ERR_CUDA;
    }
    ZKG_REQUIRE(device >= 0 && device < count, "device ordinal %d out of range (have %d)", device, count);
    DeviceGuard dg(device);
    zkg_ctx* c = new zkg_ctx();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { delete c; set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e2)); return ZKG_ERR_CUDA; }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        // experiment (round 2): cudaLimitMaxL2FetchGranularity 32 / 64 / 128 changes neither the 2^22 MSM (9.76 ms in all
        // three) nor its gather traffic pattern enough to matter; left as a switch
        const char* g = getenv("ZKG_L2_FETCH");
        if (g && *g) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    }
    *out = c;
    return ZKG_OK;
}

static void ctx_free(zkg_ctx* c) {
    if (!c) return;
    DeviceGuard dg(c->device);
    cudaStreamSynchronize(c->stream);
    c->ws.release(); c->io.release();
    for (auto& e : c->cache) cudaFree(e.p);
    for (void* d : c->cache_dead) cudaFree(d);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < c->copy_ev_count; ++i) cudaEventDestroy(c->copy_ev[i]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 5; ++i) if (c->aux_ev[i]) cudaEventDestroy(c->aux_ev[i]);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int32_t PooledCtx::acquire(int device) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i]->device == device) { ctx = g_pool[i]; g_pool.erase(g_pool.begin() + i); return ZKG_OK; }
    }
    return ctx_new(device, nullptr, &ctx);
}
PooledCtx::~PooledCtx() {
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(ctx);
}

static void fnv2(const void* key, size_t n, uint64_t* h1, uint64_t* h2) {
    const uint8_t* b = (const uint8_t*)key;
    uint64_t a = 0xcbf29ce484222325ULL, c = 0x9e3779b97f4a7c15ULL;
    for (size_t i = 0; i < n; ++i) {
        a = (a ^ b[i]) * 0x100000001b3ULL;
        c = (c + b[i] + 0x632be59bd9b4e019ULL) * 0xff51afd7ed558ccdULL;
        c ^= c >> 29;
    }
    *h1 = a; *h2 = c;
}

int32_t ctx_cache_get(zkg_ctx* ctx, const void* key, size_t key_bytes, size_t bytes, void** out, bool* fresh) {
    uint64_t h1, h2;
    fnv2(key, key_bytes, &h1, &h2);
    for (auto& e : ctx->cache)
        if (e.h1 == h1 && e.h2 == h2 && e.bytes == bytes) { *out = e.p; *fresh = false; return ZKG_OK; }
    // bounded in entries and in bytes (the per-pass NTT twiddle tables of a 2^23-point transform are 256 MiB each; a
    // prover cycling through many domains must not pin them all): drop everything once nothing in flight can still
    // read it.  ZKG_CACHE_MAX_MB overrides the 16 GiB default.
    const char* mv = getenv("ZKG_CACHE_MAX_MB");               // read on misses only
    const size_t max_bytes = (mv && *mv ? (size_t)atoll(mv) : (size_t)16384) << 20;
    if (ctx->cache.size() >= 256 || (ctx->cache_bytes + bytes > max_bytes && !ctx->cache.empty())) {
        ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
        for (void* d : ctx->cache_dead) cudaFree(d);
        ctx->cache_dead.clear();
        for (auto& e : ctx->cache) ctx->cache_dead.push_back(e.p);     // freed one flush later
        ctx->cache.clear();
        ctx->cache_bytes = 0;
    }
    void* p = nullptr;
    ZKG_CUDA(cudaMalloc(&p, bytes ? bytes : 256));
    ctx->cache.push_back({h1, h2, bytes, p});
    ctx->cache_bytes += bytes;
    *out = p; *fresh = true;
    return ZKG_OK;
}

// ark-serialize (compressed) form of Fr: 32 little-endian bytes of the CANONICAL value.
// dir 0: wire -> Montgomery image (values >= r are invalid: arkworks' deserializer rejects them); dir 1: back.
__global__ void k_fr_wire(int dir, const Fr* __restrict__ in, Fr* __restrict__ out, size_t n, int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = in[i];
    if (dir == 0) {
        // x < r ?  (compare from the top limb)
        bool lt = false, eq = true;
#pragma unroll
        for (int k = 7; k >= 0; --k) {
            uint32_t m = FrParams::mod(k);
            if (eq && x.v[k] < m) { lt = true; eq = false; }
            else if (eq && x.v[k] > m) { eq = false; }
        }
        if (!lt) { atomicExch(err, 1); return; }
        out[i] = fp_to_mont(x);
    } else {
        out[i] = fp_from_mont(x);
    }
}

int32_t ctx_copy_stream(zkg_ctx* ctx, int n_events) {
    if (!ctx->copy_stream) ZKG_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (n_events > 16) n_events = 16;
    while (ctx->copy_ev_count < n_events) {
        ZKG_CUDA(cudaEventCreateWithFlags(&ctx->copy_ev[ctx->copy_ev_count], cudaEventDisableTiming));
        ctx->copy_ev_count += 1;
    }
    return ZKG_OK;
}

int32_t ctx_aux_stream(zkg_ctx* ctx) {
    if (ctx->aux_stream) return ZKG_OK;
    int lo = 0, hi = 0;                                    // "greatest" priority is the numerically lowest value
    ZKG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ZKG_CUDA(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 5; ++i) ZKG_CUDA(cudaEventCreateWithFlags(&ctx->aux_ev[i], cudaEventDisableTiming));
    return ZKG_OK;
}

template <class F>
__global__ void k_field_op(int op, const F* __restrict__ a, const F* __restrict__ b, F* __restrict__ o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i];
    // 0 mul, 1 add, 2 sub; unit-test views of the other field routines the kernels are built from:
    // 3 dedicated squaring a^2, 4 two-term inner product a*b - b*a' with a' = a + 1 (fp_dot via f_mulsub's pattern:
    // a*b - (a+1)*b = -b), 5 inverse of a (0 -> 0)
    if (op == 3) { o[i] = fp_sqr(x); return; }
    if (op == 4) { F xs[2] = {x, fp_neg(fp_add(x, F::one()))}, ys[2] = {y, y}; o[i] = fp_dot<typename F::Params, 2>(xs, ys); return; }
    if (op == 5) { o[i] = fp_inv(x); return; }
    o[i] = op == 0 ? fp_mul(x, y) : op == 1 ? fp_add(x, y) : fp_sub(x, y);
}


// Share-wise h = a*b - c of the QAP (groth16/src/ext_wit.rs:173-177; :82-86 with the 1/Z(g) factor), fused with the
// out-mask additions that end the three preceding d_fft calls (dist-primitives/src/dfft/mod.rs:313-317): one pass over
// the share vectors instead of three mask adds, a product and a subtraction.  (a+ma)(b+mb) - (c+mc) is formed as a
// two-term inner product with ONE Montgomery reduction; canonical, so identical to the separate operations.
struct FrArgH { uint32_t v[8]; };
__global__ void __launch_bounds__(256) k_qap_h(const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c,
                                               const Fr* __restrict__ ma, const Fr* __restrict__ mb, const Fr* __restrict__ mc,
                                               int has_factor, FrArgH factor_, Fr* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = a[i], y = b[i], z = c[i];
    if (ma) x = fp_add(x, ma[i]);
    if (mb) y = fp_add(y, mb[i]);
    if (mc) z = fp_add(z, mc[i]);
    Fr xs[2] = {x, fp_neg(z)}, ys[2] = {y, Fr::one()};
    Fr h = fp_dot<FrParams, 2>(xs, ys);
    if (has_factor) {
        Fr f;
#pragma unroll
        for (int k = 0; k < 8; ++k) f.v[k] = factor_.v[k];
        h = fp_mul(h, f);
    }
    out[i] = h;
}

}  // namespace zkg

using namespace zkg;

extern "C" {

int32_t zkg_version(void) { return 100; }

int32_t zkg_device_count(int32_t* count) {
    ZKG_REQUIRE(count, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return ZKG_ERR_CUDA; }
    *count = c;
    return ZKG_OK;
}

const char* zkg_last_error(void) { return g_err; }

int32_t zkg_ctx_create(int32_t device, void* stream, zkg_ctx** out) {
    ZKG_REQUIRE(out, "out is NULL");
    return ctx_new(device, stream, out);
}
int32_t zkg_ctx_destroy(zkg_ctx* ctx) {
    ZKG_REQUIRE(ctx, "ctx is NULL");
    ctx_free(ctx);
    return ZKG_OK;
}
int32_t zkg_ctx_sync(zkg_ctx* ctx) {
    ZKG_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard dg(ctx->device);
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}
void* zkg_ctx_stream(zkg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int32_t zkg_ctx_launch_count(zkg_ctx* ctx, uint64_t* count) {
    ZKG_REQUIRE(ctx && count, "launch_count: NULL argument");
    *count = ctx->launches;
    return ZKG_OK;
}
int32_t zkg_ctx_set_profiling(zkg_ctx* ctx, int32_t enable) {
    ZKG_REQUIRE(ctx, "ctx is NULL");
    ctx->profile = enable != 0;
    ctx->ev_count = 0;
    return ZKG_OK;
}
int32_t zkg_ctx_phase_ms(zkg_ctx* ctx, int32_t phase, float* ms) {
    ZKG_REQUIRE(ctx && ms, "phase_ms: NULL argument");
    ZKG_REQUIRE(phase >= 0 && phase + 1 < ctx->ev_count, "phase %d was not recorded (profiling off, or no call yet)", phase);
    DeviceGuard dg(ctx->device);
    ZKG_CUDA(cudaEventSynchronize(ctx->ev[phase + 1]));
    ZKG_CUDA(cudaEventElapsedTime(ms, ctx->ev[phase], ctx->ev[phase + 1]));
    return ZKG_OK;
}

int32_t zkg_shutdown(void) {
    zkg::auto_register_clear();             // tables kept by ZKG_AUTO_REGISTER (msm_api.cu)
    std::vector<zkg_ctx*> all;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        all.swap(g_pool);
    }
    for (zkg_ctx* c : all) ctx_free(c);
    return ZKG_OK;
}

static int32_t fr_wire(int32_t device, int dir, const void* in, void* out, size_t n) {
    ZKG_REQUIRE(n == 0 || (in && out), "fr wire conversion: NULL argument");
    if (n == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t bytes = align_up(n * 32, 256);
    ZKG_TRY(ctx->io.reserve(2 * bytes + 256));
    uint8_t* d = (uint8_t*)ctx->io.p;
    int* d_err = (int*)(d + 2 * bytes);
    ZKG_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    ZKG_TRY(copy_h2d(d, in, n * 32, ctx->stream));
    k_fr_wire<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dir, (const Fr*)d, (Fr*)(d + bytes), n, d_err);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    int h_err = 0;
    ZKG_TRY(copy_d2h(&h_err, d_err, sizeof(int), ctx->stream));
    ZKG_TRY(copy_d2h(out, d + bytes, n * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    ZKG_REQUIRE(h_err == 0, "fr_from_wire: an element is not below the field modulus");
    return ZKG_OK;
}

int32_t zkg_fr_from_wire_bn254(int32_t device, const void* wire, uint64_t* out_mont, size_t n) { return fr_wire(device, 0, wire, out_mont, n); }
int32_t zkg_fr_to_wire_bn254(int32_t device, const uint64_t* in_mont, void* wire, size_t n) { return fr_wire(device, 1, in_mont, wire, n); }

int32_t zkg_field_op_dev(zkg_ctx* ctx, int32_t field, int32_t op, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, size_t n) {
    ZKG_REQUIRE(ctx && (field == 0 || field == 1) && op >= 0 && op <= 5 && (n == 0 || (d_a && d_b && d_out)), "field_op_dev: bad argument");
    if (n == 0) return ZKG_OK;
    DeviceGuard dg(ctx->device);
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (field == 0) k_field_op<Fr><<<blocks, 256, 0, ctx->stream>>>(op, (const Fr*)d_a, (const Fr*)d_b, (Fr*)d_out, n);
    else k_field_op<Fq><<<blocks, 256, 0, ctx->stream>>>(op, (const Fq*)d_a, (const Fq*)d_b, (Fq*)d_out, n);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

int32_t zkg_field_op(int32_t device, int32_t field, int32_t op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    ZKG_REQUIRE((field == 0 || field == 1) && op >= 0 && op <= 5 && (n == 0 || (a && b && out)), "field_op: bad argument");
    if (n == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t bytes = align_up(n * 32, 256);
    ZKG_TRY(ctx->io.reserve(3 * bytes));
    uint8_t* d = (uint8_t*)ctx->io.p;
    ZKG_TRY(copy_h2d(d, a, n * 32, ctx->stream));
    ZKG_TRY(copy_h2d(d + bytes, b, n * 32, ctx->stream));
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (field == 0) k_field_op<Fr><<<blocks, 256, 0, ctx->stream>>>(op, (const Fr*)d, (const Fr*)(d + bytes), (Fr*)(d + 2 * bytes), n);
    else k_field_op<Fq><<<blocks, 256, 0, ctx->stream>>>(op, (const Fq*)d, (const Fq*)(d + bytes), (Fq*)(d + 2 * bytes), n);
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(copy_d2h(out, d + 2 * bytes, n * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}


// ---- peer-visible buffers for one-process-per-GPU sharding (CUDA IPC) ------------------------------------------------
int32_t zkg_shared_alloc(zkg_ctx* ctx, size_t bytes, void** d_ptr, uint8_t ipc_handle[64]) {
    ZKG_REQUIRE(ctx && d_ptr && ipc_handle && bytes, "shared_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    DeviceGuard dg(ctx->device);
    void* p = nullptr;
    ZKG_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return ZKG_ERR_NCCL; }
    memcpy(ipc_handle, &h, 64);
    *d_ptr = p;
    return ZKG_OK;
}
int32_t zkg_shared_open(zkg_ctx* ctx, const uint8_t ipc_handle[64], void** d_ptr) {
    ZKG_REQUIRE(ctx && d_ptr && ipc_handle, "shared_open: bad argument");
    DeviceGuard dg(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle failed: %s (no peer access between the two GPUs?)", cudaGetErrorString(e)); return ZKG_ERR_NCCL; }
    return ZKG_OK;
}
int32_t zkg_shared_close(zkg_ctx* ctx, void* d_ptr) {
    ZKG_REQUIRE(ctx && d_ptr, "shared_close: bad argument");
    DeviceGuard dg(ctx->device);
    ZKG_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return ZKG_OK;
}
int32_t zkg_shared_free(zkg_ctx* ctx, void* d_ptr) {
    ZKG_REQUIRE(ctx && d_ptr, "shared_free: bad argument");
    DeviceGuard dg(ctx->device);
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    ZKG_CUDA(cudaFree(d_ptr));
    return ZKG_OK;
}

int32_t zkg_qap_h_bn254_dev(zkg_ctx* ctx, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_c, const uint64_t* d_mask_a,
                            const uint64_t* d_mask_b, const uint64_t* d_mask_c, const uint64_t* factor, uint64_t* d_out, size_t n) {
    ZKG_REQUIRE(ctx && (n == 0 || (d_a && d_b && d_c && d_out)), "qap_h_dev: NULL argument");
    if (n == 0) return ZKG_OK;
    DeviceGuard dg(ctx->device);
    FrArgH f;
    memset(&f, 0, sizeof f);
    if (factor) memcpy(f.v, factor, 32);
    k_qap_h<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const Fr*)d_a, (const Fr*)d_b, (const Fr*)d_c, (const Fr*)d_mask_a,
                                                                 (const Fr*)d_mask_b, (const Fr*)d_mask_c, factor ? 1 : 0, f, (Fr*)d_out, n);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

int32_t zkg_qap_h_bn254(int32_t device, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* mask_a,
                        const uint64_t* mask_b, const uint64_t* mask_c, const uint64_t* factor, uint64_t* out, size_t n) {
    ZKG_REQUIRE(n == 0 || (a && b && c && out), "qap_h: NULL argument");
    if (n == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t bytes = align_up(n * 32, 256);
    ZKG_TRY(ctx->io.reserve(7 * bytes));
    uint8_t* d = (uint8_t*)ctx->io.p;
    const uint64_t* src[6] = {a, b, c, mask_a, mask_b, mask_c};
    const uint64_t* dv[6];
    for (int k = 0; k < 6; ++k) {
        dv[k] = nullptr;
        if (!src[k]) continue;
        ZKG_TRY(copy_h2d(d + k * bytes, src[k], n * 32, ctx->stream));
        dv[k] = (const uint64_t*)(d + k * bytes);
    }
    ZKG_TRY(zkg_qap_h_bn254_dev(ctx, dv[0], dv[1], dv[2], dv[3], dv[4], dv[5], factor, (uint64_t*)(d + 6 * bytes), n));
    ZKG_TRY(copy_d2h(out, d + 6 * bytes, n * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

}  // extern "C"
