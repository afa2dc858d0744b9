// msm_api.cu -- C-ABI entry points of the MSM path (dispatch over G1/G2 + device-resident CRS registry).
// Declarations and reference citations: include/zksaas_gpu.h.
#include "common.cuh"

namespace zkg {
#define ZKG_MSM_DECLARE(G)                                                                                          \
    int32_t msm_run_##G(zkg_ctx* ctx, const void* d_bases, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t pack_bases_##G(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed);               \
    int32_t msm_host_##G(int device, const void* bases, size_t stride, size_t n_bases, const uint64_t* scalars,     \
                         size_t n_scalars, uint64_t* out_xyz);                                                      \
    int32_t combine_##G(zkg_ctx* ctx, const uint64_t* d_parts, size_t n, uint64_t* d_out);                          \
    int32_t fixed_base_##G(zkg_ctx* ctx, const uint64_t* d_scalars, size_t n, void* d_packed);
ZKG_MSM_DECLARE(g1)
ZKG_MSM_DECLARE(g2)

struct BaseSet { int device; int group; size_t n; void* d_packed; };
static std::mutex g_bases_mu;
static std::vector<BaseSet*> g_bases;   // handle = index + 1
static inline size_t packed_bytes(int group) { return group == 1 ? 64 : 128; }
}  // namespace zkg

using namespace zkg;

extern "C" {

int32_t zkg_msm_bn254_g1(int32_t device, const void* bases, size_t base_stride, size_t n_bases, const uint64_t* scalars,
                         size_t n_scalars, uint64_t out_xyz[12]) {
    return msm_host_g1(device, bases, base_stride, n_bases, scalars, n_scalars, out_xyz);
}
int32_t zkg_msm_bn254_g2(int32_t device, const void* bases, size_t base_stride, size_t n_bases, const uint64_t* scalars,
                         size_t n_scalars, uint64_t out_xyz[24]) {
    return msm_host_g2(device, bases, base_stride, n_bases, scalars, n_scalars, out_xyz);
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

int32_t zkg_bases_register(int32_t device, int32_t group, const void* bases, size_t base_stride, size_t n,
                           uint64_t* handle) {
    ZKG_REQUIRE(handle && (group == 1 || group == 2) && (n == 0 || bases), "bases_register: bad argument");
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    BaseSet* bs = new BaseSet{ctx->device, group, n, nullptr};
    if (n) {
        cudaError_t e = cudaMalloc(&bs->d_packed, n * packed_bytes(group));
        if (e != cudaSuccess) {
            delete bs;
            set_error("bases_register: cudaMalloc(%zu) failed: %s", n * packed_bytes(group), cudaGetErrorString(e));
            return ZKG_ERR_OOM;
        }
        int32_t rc = ctx->io.reserve(n * base_stride);
        if (rc == ZKG_OK) {
            cudaError_t e2 = cudaMemcpyAsync(ctx->io.p, bases, n * base_stride, cudaMemcpyHostToDevice, ctx->stream);
            if (e2 != cudaSuccess) { set_error("bases_register: H2D failed: %s", cudaGetErrorString(e2)); rc = ZKG_ERR_CUDA; }
        }
        if (rc == ZKG_OK)
            rc = group == 1 ? pack_bases_g1(ctx, ctx->io.p, base_stride, n, bs->d_packed)
                            : pack_bases_g2(ctx, ctx->io.p, base_stride, n, bs->d_packed);
        if (rc == ZKG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("bases_register: sync failed"); rc = ZKG_ERR_CUDA; }
        if (rc != ZKG_OK) { cudaFree(bs->d_packed); delete bs; return rc; }
    }
    std::lock_guard<std::mutex> lk(g_bases_mu);
    g_bases.push_back(bs);
    *handle = g_bases.size();
    return ZKG_OK;
}

int32_t zkg_bases_release(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    ZKG_REQUIRE(handle >= 1 && handle <= g_bases.size() && g_bases[handle - 1], "bases_release: bad handle");
    BaseSet* bs = g_bases[handle - 1];
    g_bases[handle - 1] = nullptr;
    DeviceGuard dg(bs->device);
    if (bs->d_packed) cudaFree(bs->d_packed);
    delete bs;
    return ZKG_OK;
}

int32_t zkg_msm_bn254_registered(uint64_t handle, const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz) {
    BaseSet bs;
    {
        std::lock_guard<std::mutex> lk(g_bases_mu);
        ZKG_REQUIRE(handle >= 1 && handle <= g_bases.size() && g_bases[handle - 1], "msm_registered: bad handle");
        bs = *g_bases[handle - 1];
    }
    if (bs.n != n_scalars) {
        set_error("msm: bases.len() = %zu, scalars.len() = %zu", bs.n, n_scalars);
        return ZKG_ERR_LEN_MISMATCH;
    }
    ZKG_REQUIRE(out_xyz && (n_scalars == 0 || scalars), "msm_registered: NULL argument");
    PooledCtx pc;
    ZKG_TRY(pc.acquire(bs.device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t sc_bytes = align_up(n_scalars * 32, 256);
    ZKG_TRY(ctx->io.reserve(sc_bytes + 256));
    uint8_t* d_sc = (uint8_t*)ctx->io.p;
    void* d_out = d_sc + sc_bytes;
    if (n_scalars) ZKG_CUDA(cudaMemcpyAsync(d_sc, scalars, n_scalars * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZKG_TRY(bs.group == 1 ? msm_run_g1(ctx, bs.d_packed, (const uint64_t*)d_sc, n_scalars, d_out, 0)
                          : msm_run_g2(ctx, bs.d_packed, (const uint64_t*)d_sc, n_scalars, d_out, 0));
    ZKG_CUDA(cudaMemcpyAsync(out_xyz, d_out, bs.group == 1 ? 96 : 192, cudaMemcpyDeviceToHost, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

}  // extern "C"
