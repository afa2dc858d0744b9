// msm_api.cu -- C-ABI entry points of the MSM path (dispatch over G1/G2 + device-resident CRS registry).
// Declarations and reference citations: include/zksaas_gpu.h.
#include "common.cuh"
#include "host_fr.hpp"

namespace zkg {
#define ZKG_MSM_DECLARE(G)                                                                                          \
    int32_t msm_run_##G(zkg_ctx* ctx, const void* d_bases, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t pack_bases_##G(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed);               \
    int32_t msm_host_##G(int device, const void* bases, size_t stride, size_t n_bases, const uint64_t* scalars,     \
                         size_t n_scalars, uint64_t* out_xyz);                                                      \
    int32_t combine_##G(zkg_ctx* ctx, const uint64_t* d_parts, size_t n, uint64_t* d_out);                          \
    int32_t fixed_base_##G(zkg_ctx* ctx, const uint64_t* d_scalars, size_t n, void* d_packed);                      \
    int32_t prepare_##G(zkg_ctx* ctx, const void* d_bases, size_t n, int c, void* d_table);                         \
    int32_t msm_run_prepared_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t msm_run_prepared_host_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* h_scalars, size_t n, void* d_out); \
    int32_t crs_det_pack_##G(int device, const void* bases, size_t stride, size_t n, int l, int n_parties,          \
                             const uint32_t* h_scal, void* const* out_by_party, size_t out_stride);
ZKG_MSM_DECLARE(g1)
ZKG_MSM_DECLARE(g2)

// device-resident CRS share: table[w*n + i] = 2^(c*w) * P_i (packed affine), W = 254/c + 1 window shifts
struct BaseSet { int device; int group; size_t n; int c; void* d_table; };
int msm_pick_c_merged_host(size_t n);
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

static int32_t base_set_create(zkg_ctx* ctx, int32_t group, const void* d_packed, size_t n, uint64_t* handle) {
    BaseSet* bs = new BaseSet{ctx->device, group, n, 0, nullptr};
    if (n) {
        bs->c = msm_pick_c_merged_host(n);
        int env_c = getenv("ZKG_MSM_PREP_C") ? atoi(getenv("ZKG_MSM_PREP_C")) : 0;
        if (env_c >= 4 && env_c <= 23) bs->c = env_c;
        const int W = 254 / bs->c + 1;
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
    g_bases.push_back(bs);
    *handle = g_bases.size();
    return ZKG_OK;
}

static int32_t base_set_get(uint64_t handle, BaseSet* out) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    ZKG_REQUIRE(handle >= 1 && handle <= g_bases.size() && g_bases[handle - 1], "bad bases handle %llu", (unsigned long long)handle);
    *out = *g_bases[handle - 1];
    return ZKG_OK;
}

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

int32_t zkg_bases_register_dev(zkg_ctx* ctx, int32_t group, const void* d_bases_packed, size_t n, uint64_t* handle) {
    ZKG_REQUIRE(ctx && handle && (group == 1 || group == 2) && (n == 0 || d_bases_packed), "bases_register_dev: bad argument");
    DeviceGuard dg(ctx->device);
    return base_set_create(ctx, group, d_bases_packed, n, handle);
}

int32_t zkg_bases_release(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_bases_mu);
    ZKG_REQUIRE(handle >= 1 && handle <= g_bases.size() && g_bases[handle - 1], "bases_release: bad handle");
    BaseSet* bs = g_bases[handle - 1];
    g_bases[handle - 1] = nullptr;
    DeviceGuard dg(bs->device);
    cudaDeviceSynchronize();
    if (bs->d_table) cudaFree(bs->d_table);
    delete bs;
    return ZKG_OK;
}

int32_t zkg_msm_bn254_registered_dev(zkg_ctx* ctx, uint64_t handle, const uint64_t* d_scalars, size_t n_scalars,
                                     uint64_t* d_out, int32_t partial) {
    ZKG_REQUIRE(ctx && d_out, "msm_registered_dev: NULL argument");
    BaseSet bs;
    ZKG_TRY(base_set_get(handle, &bs));
    ZKG_REQUIRE(bs.device == ctx->device, "msm_registered_dev: bases live on device %d, context on %d", bs.device, ctx->device);
    if (bs.n != n_scalars) {
        set_error("msm: bases.len() = %zu, scalars.len() = %zu", bs.n, n_scalars);
        return ZKG_ERR_LEN_MISMATCH;
    }
    DeviceGuard dg(ctx->device);
    return bs.group == 1 ? msm_run_prepared_g1(ctx, bs.d_table, bs.c, d_scalars, n_scalars, d_out, partial ? 1 : 0)
                         : msm_run_prepared_g2(ctx, bs.d_table, bs.c, d_scalars, n_scalars, d_out, partial ? 1 : 0);
}

int32_t zkg_msm_bn254_registered(uint64_t handle, const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz) {
    BaseSet bs;
    ZKG_TRY(base_set_get(handle, &bs));
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
    ZKG_TRY(ctx->io.reserve(sc_bytes + 512));
    void* d_out = (uint8_t*)ctx->io.p + sc_bytes;
    ZKG_TRY(bs.group == 1 ? msm_run_prepared_host_g1(ctx, bs.d_table, bs.c, scalars, n_scalars, d_out)
                          : msm_run_prepared_host_g2(ctx, bs.d_table, bs.c, scalars, n_scalars, d_out));
    ZKG_TRY(copy_d2h(out_xyz, d_out, bs.group == 1 ? 96 : 192, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

}  // extern "C"
