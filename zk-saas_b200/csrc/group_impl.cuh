// group_impl.cuh -- small group-element operations around the MSM (SURVEY.md 8f rows 1-2 and the king side of d_msm):
//
//   * `pp.unpack_missing_shares(&shares, &parties)` over GROUP elements followed by the sum of the l results --
//     the king closure of d_msm, dist-primitives/src/dmsm/mod.rs:85-87 (secret-sharing/src/pss.rs:141-166 when all
//     n shares arrived, :170-221 Lagrange otherwise); also the A/B/C recombination of groth16/examples/sha256.rs:375-377.
//     unpack2 / lagrange_unpack are linear maps with a fixed l x n_recv matrix over Fr, so row i is the n_recv-term
//     linear combination  sum_j M[i][j] * share_j : one 254-bit double-and-add per (row, share), all in parallel,
//     then a row sum.  n_recv <= 32 points: latency, not throughput, is what matters here.
//   * ark-serialize 0.4 COMPRESSED points, the wire form of the element d_msm ships (mpc-net/src/ser_net.rs:25
//     serialize_compressed, :40,:119 deserialize_compressed = Compress::Yes + Validate::Yes).  Layout and flag
//     semantics are restated (include/zksaas_gpu.h, and the test oracle) from the public ark-ec / ark-serialize
//     0.4.2 behaviour; the crates are un-vendored, so this format is NOT pinned by a reference-held vector.
//
// Instantiated for Fq (G1) in group_g1.cu and Fq2 (G2) in group_g2.cu.
#pragma once
#include "common.cuh"
#include "ec.cuh"
#include "msm_common.cuh"

namespace zkg {

template <class T>
__device__ __forceinline__ T g_load(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
    return r;
}
template <class T>
__device__ __forceinline__ void g_store(T* p, const T& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
}

// arkworks Projective (Jacobian X, Y, Z; identity Z = 0) -> XYZZ: same X, Y with ZZ = Z^2, ZZZ = Z^3
template <class F>
ZKG_NI XYZZ<F> jac_image_to_xyzz(const F* p) {
    F Z = g_load(p + 2);
    if (Z.is_zero()) return XYZZ<F>::inf();
    XYZZ<F> r;
    r.x = g_load(p);
    r.y = g_load(p + 1);
    r.zz = f_sqr(Z);
    r.zzz = f_mul(r.zz, Z);
    return r;
}
// normalised Jacobian image: (x, y, 1) or (1, 1, 0)
template <class F>
ZKG_NI void store_normalised(F* out, const XYZZ<F>& a) {
    if (a.is_inf()) { g_store(out, F::one()); g_store(out + 1, F::one()); g_store(out + 2, F::zero()); return; }
    Affine<F> p = xyzz_to_affine(a);
    g_store(out, p.x); g_store(out + 1, p.y); g_store(out + 2, F::one());
}

// terms[row * n + j] = scal[row][j] * points[j]; one block per term (the 254 dependent doublings are pure latency, so
// every term gets its own SM sub-partition).  scal: rows x n canonical scalars, 8 x u32 each.
template <class F>
__global__ void __launch_bounds__(32) k_group_scale(const F* __restrict__ pts_xyz, uint32_t n, const uint32_t* __restrict__ scal,
                                                    XYZZ<F>* __restrict__ terms) {
    if (threadIdx.x != 0) return;
    const uint32_t t = blockIdx.x, j = t % n;
    const uint32_t* s = scal + (size_t)t * 8;
    XYZZ<F> p = jac_image_to_xyzz(pts_xyz + (size_t)j * 3);
    XYZZ<F> acc = XYZZ<F>::inf();
    int top = 253;
    while (top >= 0 && !((s[top >> 5] >> (top & 31)) & 1)) --top;
    for (int b = top; b >= 0; --b) {
        xyzz_dbl(acc);
        if ((s[b >> 5] >> (b & 31)) & 1) xyzz_add(acc, p);
    }
    g_store(terms + t, acc);
}
// out_rows[row] = sum_j terms[row][j] (normalised), out_sum = sum of the rows (normalised); either may be NULL
template <class F>
__global__ void __launch_bounds__(32) k_group_rowsum(const XYZZ<F>* __restrict__ terms, uint32_t rows, uint32_t n,
                                                     F* __restrict__ out_rows, F* __restrict__ out_sum) {
    __shared__ XYZZ<F> sh[32];
    const uint32_t lane = threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::inf();
    if (lane < rows)
        for (uint32_t j = 0; j < n; ++j) xyzz_add(acc, g_load(terms + (size_t)lane * n + j));
    sh[lane] = acc;
    __syncthreads();
    if (lane < rows && out_rows) store_normalised(out_rows + (size_t)lane * 3, acc);
    if (lane == 0 && out_sum) {
        XYZZ<F> tot = sh[0];
        for (uint32_t i = 1; i < rows; ++i) xyzz_add(tot, sh[i]);
        store_normalised(out_sum, tot);
    }
}

// host-pointer entry: shares (n_recv Jacobian images) -> l unpacked points and / or their sum
template <class F>
static int32_t group_unpack_host(int device, const uint64_t* shares_xyz, uint32_t n_recv, const uint32_t* h_scal, uint32_t rows,
                                 uint64_t* out_rows, uint64_t* out_sum) {
    ZKG_REQUIRE(n_recv >= 1 && n_recv <= 32 && rows >= 1 && rows <= 32, "group unpack: %u shares x %u rows unsupported", n_recv, rows);
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t pts_b = align_up((size_t)n_recv * 3 * sizeof(F), 256), sc_b = align_up((size_t)rows * n_recv * 32, 256);
    const size_t tm_b = align_up((size_t)rows * n_recv * sizeof(XYZZ<F>), 256), or_b = align_up((size_t)rows * 3 * sizeof(F), 256);
    ZKG_TRY(ctx->io.reserve(pts_b + sc_b + tm_b + or_b + 3 * sizeof(F) + 256));
    uint8_t* d = (uint8_t*)ctx->io.p;
    uint8_t *d_pts = d, *d_sc = d + pts_b, *d_tm = d_sc + sc_b, *d_or = d_tm + tm_b, *d_os = d_or + or_b;
    ZKG_TRY(copy_h2d(d_pts, shares_xyz, (size_t)n_recv * 3 * sizeof(F), ctx->stream));
    ZKG_TRY(copy_h2d(d_sc, h_scal, (size_t)rows * n_recv * 32, ctx->stream));
    k_group_scale<F><<<rows * n_recv, 32, 0, ctx->stream>>>((const F*)d_pts, n_recv, (const uint32_t*)d_sc, (XYZZ<F>*)d_tm);
    k_group_rowsum<F><<<1, 32, 0, ctx->stream>>>((const XYZZ<F>*)d_tm, rows, n_recv, out_rows ? (F*)d_or : nullptr,
                                                 out_sum ? (F*)d_os : nullptr);
    ctx->launches += 2;
    ZKG_CUDA(cudaGetLastError());
    if (out_rows) ZKG_TRY(copy_d2h(out_rows, d_or, (size_t)rows * 3 * sizeof(F), ctx->stream));
    if (out_sum) ZKG_TRY(copy_d2h(out_sum, d_os, 3 * sizeof(F), ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// compressed wire form
// ------------------------------------------------------------------------------------------------------------------
static constexpr int WIRE_ERR_FLAGS = 1, WIRE_ERR_RANGE = 2, WIRE_ERR_CURVE = 3, WIRE_ERR_SUBGROUP = 4;

// canonical(a) > canonical(-a), i.e. a > (q-1)/2
ZKG_NI bool fq_gt_neg(const Fq& a) {
    Fq c = fp_from_mont(a), d = fp_from_mont(fp_neg(a));
    return big_lt(d.v, c.v);
}
// square root for q = 3 mod 4: a^((q+1)/4); false for a non-residue
ZKG_NI bool fq_sqrt(const Fq& a, Fq* out) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = FqParams::mod(i);
    e[0] += 1;                                     // q = ...47 (hex): no carry
#pragma unroll
    for (int i = 0; i < 7; ++i) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
    e[7] >>= 2;
    Fq y = fp_pow(a, e, 8);
    *out = y;
    return fp_sqr(y) == a;
}
// 32 little-endian bytes (flag bits already cleared) -> Montgomery image; false if the value is >= q
ZKG_NI bool fq_from_canonical_words(const uint32_t* w, Fq* out) {
    Fq x, m;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x.v[i] = w[i]; m.v[i] = FqParams::mod(i); }
    if (!big_lt(x.v, m.v)) return false;
    *out = fp_to_mont(x);
    return true;
}

template <class F> struct WireOps;
template <> struct WireOps<Fq> {
    static constexpr int WORDS = 8;                                          // 32 bytes
    __device__ static Fq curve_b() { Fq b; for (int i = 0; i < 8; ++i) b.v[i] = BN254_G1_B_MONT_L(i); return b; }
    __device__ static bool y_is_negative(const Fq& y) { return fq_gt_neg(y); }
    __device__ static void x_to_words(const Fq& x, uint32_t* w) { Fq c = fp_from_mont(x); for (int i = 0; i < 8; ++i) w[i] = c.v[i]; }
    __device__ static bool x_from_words(const uint32_t* w, Fq* x) { return fq_from_canonical_words(w, x); }
    __device__ static bool sqrt(const Fq& a, Fq* y) { return fq_sqrt(a, y); }
    __device__ static bool in_subgroup(const Affine<Fq>&) { return true; }   // cofactor 1
};
template <> struct WireOps<Fq2> {
    static constexpr int WORDS = 16;                                         // c0 then c1
    __device__ static Fq2 curve_b() {
        Fq2 b;
        for (int i = 0; i < 8; ++i) { b.c0.v[i] = BN254_G2_B_C0_MONT_L(i); b.c1.v[i] = BN254_G2_B_C1_MONT_L(i); }
        return b;
    }
    // QuadExtField ordering: c1 is the most significant coordinate
    __device__ static bool y_is_negative(const Fq2& y) { return y.c1.is_zero() ? fq_gt_neg(y.c0) : fq_gt_neg(y.c1); }
    __device__ static void x_to_words(const Fq2& x, uint32_t* w) {
        Fq a = fp_from_mont(x.c0), b = fp_from_mont(x.c1);
        for (int i = 0; i < 8; ++i) { w[i] = a.v[i]; w[8 + i] = b.v[i]; }
    }
    __device__ static bool x_from_words(const uint32_t* w, Fq2* x) {
        return fq_from_canonical_words(w, &x->c0) && fq_from_canonical_words(w + 8, &x->c1);
    }
    // norm method: a = (c0 + c1 u)^2 with c0^2 = (a0 +- sqrt(a0^2 + a1^2)) / 2, c1 = a1 / (2 c0)
    __device__ static bool sqrt(const Fq2& a, Fq2* y) {
        if (a.is_zero()) { *y = Fq2::zero(); return true; }
        Fq s;
        if (a.c1.is_zero()) {
            if (fq_sqrt(a.c0, &s)) { y->c0 = s; y->c1 = Fq::zero(); return true; }
            fq_sqrt(fp_neg(a.c0), &s);             // -1 is a non-residue, so -a0 is a square
            y->c0 = Fq::zero(); y->c1 = s;
            return true;
        }
        Fq alpha;
        if (!fq_sqrt(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)), &alpha)) return false;
        Fq two = fp_dbl(Fq::one()), half = fp_inv(two);
        Fq delta = fp_mul(fp_add(a.c0, alpha), half);
        if (!fq_sqrt(delta, &s)) {
            delta = fp_sub(delta, alpha);
            if (!fq_sqrt(delta, &s)) return false;
        }
        y->c0 = s;
        y->c1 = fp_mul(a.c1, fp_inv(fp_dbl(s)));
        return f_sqr(*y) == a;
    }
    // Validate::Yes for G2: r * P == identity
    __device__ static bool in_subgroup(const Affine<Fq2>& p) {
        XYZZ<Fq2> acc = XYZZ<Fq2>::inf();
        for (int b = 253; b >= 0; --b) {
            xyzz_dbl(acc);
            if ((FrParams::mod(b >> 5) >> (b & 31)) & 1) xyzz_madd(acc, p, false);
        }
        return acc.is_inf();
    }
};

template <class F>
__global__ void __launch_bounds__(64) k_point_to_wire(const F* __restrict__ pts_xyz, size_t n, uint32_t* __restrict__ wire) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = WireOps<F>::WORDS;
    uint32_t w[W];
    XYZZ<F> p = jac_image_to_xyzz(pts_xyz + i * 3);
    if (p.is_inf()) {
#pragma unroll
        for (int k = 0; k < W; ++k) w[k] = 0;
        w[W - 1] = 0x40000000u;                                    // PointAtInfinity: bit 6 of the last byte
    } else {
        Affine<F> a = xyzz_to_affine(p);
        WireOps<F>::x_to_words(a.x, w);
        if (WireOps<F>::y_is_negative(a.y)) w[W - 1] |= 0x80000000u;   // YIsNegative: bit 7 of the last byte
    }
#pragma unroll
    for (int k = 0; k < W; ++k) wire[i * W + k] = w[k];
}

template <class F>
__global__ void __launch_bounds__(64) k_point_from_wire(const uint32_t* __restrict__ wire, size_t n, F* __restrict__ pts_xyz,
                                                        int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = WireOps<F>::WORDS;
    uint32_t w[W];
#pragma unroll
    for (int k = 0; k < W; ++k) w[k] = wire[i * W + k];
    const uint32_t flags = w[W - 1] >> 30;                         // bit 1: negative, bit 0: infinity
    w[W - 1] &= 0x3fffffffu;
    F* out = pts_xyz + i * 3;
    auto fail = [&](int code) {
        atomicCAS(err, 0, code);
        g_store(out, F::one()); g_store(out + 1, F::one()); g_store(out + 2, F::zero());
    };
    if (flags == 3) { fail(WIRE_ERR_FLAGS); return; }
    F x;
    if (!WireOps<F>::x_from_words(w, &x)) { fail(WIRE_ERR_RANGE); return; }
    if (flags == 1) { g_store(out, F::one()); g_store(out + 1, F::one()); g_store(out + 2, F::zero()); return; }
    F y;
    if (!WireOps<F>::sqrt(f_add(f_mul(f_sqr(x), x), WireOps<F>::curve_b()), &y)) { fail(WIRE_ERR_CURVE); return; }
    if (WireOps<F>::y_is_negative(y) != (flags == 2)) y = f_neg(y);
    Affine<F> a;
    a.x = x; a.y = y;
    if (!WireOps<F>::in_subgroup(a)) { fail(WIRE_ERR_SUBGROUP); return; }
    g_store(out, x); g_store(out + 1, y); g_store(out + 2, F::one());
}

template <class F>
static int32_t point_wire_host(int device, int dir, const void* in, void* out, size_t n) {
    ZKG_REQUIRE(n == 0 || (in && out), "point wire conversion: NULL argument");
    if (n == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t wire_rec = WireOps<F>::WORDS * 4, img_rec = 3 * sizeof(F);
    const size_t in_b = align_up(n * (dir == 0 ? wire_rec : img_rec), 256), out_b = align_up(n * (dir == 0 ? img_rec : wire_rec), 256);
    ZKG_TRY(ctx->io.reserve(in_b + out_b + 256));
    uint8_t* d = (uint8_t*)ctx->io.p;
    int* d_err = (int*)(d + in_b + out_b);
    ZKG_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    ZKG_TRY(copy_h2d(d, in, n * (dir == 0 ? wire_rec : img_rec), ctx->stream));
    const unsigned blocks = (unsigned)((n + 63) / 64);
    if (dir == 0) k_point_from_wire<F><<<blocks, 64, 0, ctx->stream>>>((const uint32_t*)d, n, (F*)(d + in_b), d_err);
    else k_point_to_wire<F><<<blocks, 64, 0, ctx->stream>>>((const F*)d, n, (uint32_t*)(d + in_b));
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    int h_err = 0;
    ZKG_TRY(copy_d2h(&h_err, d_err, sizeof(int), ctx->stream));
    ZKG_TRY(copy_d2h(out, d + in_b, n * (dir == 0 ? img_rec : wire_rec), ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    static const char* const why[] = {"", "both flag bits set", "x is not below the field modulus", "x is not the abscissa of a curve point",
                                      "point is not in the prime-order subgroup"};
    ZKG_REQUIRE(h_err == 0, "point_from_wire: invalid encoding (%s)", why[h_err & 7]);
    return ZKG_OK;
}

#define ZKG_GROUP_DECLARE(G)                                                                                              \
    int32_t group_unpack_##G(int device, const uint64_t* shares_xyz, uint32_t n_recv, const uint32_t* h_scal, uint32_t rows, \
                             uint64_t* out_rows, uint64_t* out_sum);                                                     \
    int32_t point_wire_##G(int device, int dir, const void* in, void* out, size_t n);
ZKG_GROUP_DECLARE(g1)
ZKG_GROUP_DECLARE(g2)

#define ZKG_GROUP_DEFINE(G, F)                                                                                            \
    int32_t group_unpack_##G(int device, const uint64_t* shares_xyz, uint32_t n_recv, const uint32_t* h_scal, uint32_t rows, \
                             uint64_t* out_rows, uint64_t* out_sum) {                                                    \
        return group_unpack_host<F>(device, shares_xyz, n_recv, h_scal, rows, out_rows, out_sum);                        \
    }                                                                                                                     \
    int32_t point_wire_##G(int device, int dir, const void* in, void* out, size_t n) {                                   \
        return point_wire_host<F>(device, dir, in, out, n);                                                              \
    }

}  // namespace zkg
