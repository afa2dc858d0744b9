// tests/host_emu/emu.cpp -- compiles the product's device arithmetic headers for the HOST
// (ZKG_HOST_EMU: PTX carry-chain primitives emulated in C++), so the exact limb schedules used
// by the CUDA kernels can be checked against the oracle on a machine without a GPU.
// Test infrastructure only.
#define ZKG_HOST_EMU 1
#include <cstring>
#include "../../zk-saas_b200/csrc/fp.cuh"

using namespace zkg;

template <class F>
static F ld(const uint64_t* p) { F r; memcpy(r.v, p, 32); return r; }
template <class F>
static void st(uint64_t* p, const F& a) { memcpy(p, a.v, 32); }

#define FIELD_API(NAME, F)                                                                         \
    extern "C" void emu_##NAME##_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_mul(ld<F>(a + 4 * i), ld<F>(b + 4 * i)));  \
    }                                                                                              \
    extern "C" void emu_##NAME##_sqr(const uint64_t* a, uint64_t* o, size_t n) {                   \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_sqr(ld<F>(a + 4 * i)));                    \
    }                                                                                              \
    extern "C" void emu_##NAME##_add(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_add(ld<F>(a + 4 * i), ld<F>(b + 4 * i)));  \
    }                                                                                              \
    extern "C" void emu_##NAME##_sub(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_sub(ld<F>(a + 4 * i), ld<F>(b + 4 * i)));  \
    }                                                                                              \
    extern "C" void emu_##NAME##_neg(const uint64_t* a, uint64_t* o, size_t n) {                   \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_neg(ld<F>(a + 4 * i)));                    \
    }                                                                                              \
    extern "C" void emu_##NAME##_inv(const uint64_t* a, uint64_t* o, size_t n) {                   \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_inv(ld<F>(a + 4 * i)));                    \
    }                                                                                              \
    extern "C" void emu_##NAME##_dot(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n, int k) { \
        for (size_t i = 0; i < n; ++i) {                                                           \
            F x[4], y[4];                                                                          \
            for (int j = 0; j < k; ++j) { x[j] = ld<F>(a + 4 * (i * k + j)); y[j] = ld<F>(b + 4 * (i * k + j)); } \
            F r = k == 1 ? fp_dot<F::Params, 1>(x, y) : k == 2 ? fp_dot<F::Params, 2>(x, y)        \
                : k == 3 ? fp_dot<F::Params, 3>(x, y) : fp_dot<F::Params, 4>(x, y);                \
            st(o + 4 * i, r);                                                                      \
        }                                                                                          \
    }                                                                                              \
    extern "C" void emu_##NAME##_inv_fermat(const uint64_t* a, uint64_t* o, size_t n) {            \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_inv_fermat(ld<F>(a + 4 * i)));             \
    }                                                                                              \
    extern "C" void emu_##NAME##_from_mont(const uint64_t* a, uint64_t* o, size_t n) {             \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_from_mont(ld<F>(a + 4 * i)));              \
    }                                                                                              \
    extern "C" void emu_##NAME##_to_mont(const uint64_t* a, uint64_t* o, size_t n) {               \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_to_mont(ld<F>(a + 4 * i)));                \
    }

FIELD_API(fr, Fr)
FIELD_API(fq, Fq)

// ---- group arithmetic + MSM digit logic of the product, host-emulated -------------------------
#include <vector>
#include "../../zk-saas_b200/csrc/ec.cuh"
#include "../../zk-saas_b200/csrc/msm_common.cuh"

template <class F> static F ldf(const uint64_t* p) { F r; memcpy(&r, p, sizeof(F)); return r; }
template <class F> static void stf(uint64_t* p, const F& a) { memcpy(p, &a, sizeof(F)); }

// Bucket-method MSM written with exactly the device building blocks (msm_signed_digits, xyzz_madd,
// the multi-level running-sum reduction of k_reduce_lvl, Horner + xyzz_to_affine), single-threaded.
template <class F>
static void emu_msm(const uint64_t* bases_packed, const uint64_t* scalars_mont, size_t n, int c, uint64_t* out_xyz) {
    const int W = msm_num_windows(c);
    const uint32_t nb = 1u << (c - 1);
    std::vector<XYZZ<F>> buckets((size_t)W * nb, XYZZ<F>::inf());
    std::vector<uint32_t> dig(W);
    for (size_t i = 0; i < n; ++i) {
        Fr s = fp_from_mont(ld<Fr>(scalars_mont + 4 * i));
        msm_signed_digits(s.v, c, W, dig.data());
        Affine<F> p = ldf<Affine<F>>(bases_packed + i * (sizeof(Affine<F>) / 8));
        for (int w = 0; w < W; ++w)
            if (dig[w] != MSM_DIGIT_NONE) xyzz_madd(buckets[(size_t)w * nb + (dig[w] & 0x7fffffffu)], p, (dig[w] >> 31) != 0);
    }
    // multi-level reduction, L = 8 (same invariant as k_reduce_lvl)
    const uint32_t L = 8;
    std::vector<XYZZ<F>> Rin = buckets, Cin;
    uint32_t n_in = nb;
    int log2_M = 0;
    while (true) {
        uint32_t n_out = (n_in + L - 1) / L;
        std::vector<XYZZ<F>> Rout((size_t)W * n_out), Cout((size_t)W * n_out);
        for (int w = 0; w < W; ++w)
            for (uint32_t u = 0; u < n_out; ++u) {
                uint32_t lo = u * L, hi = lo + L < n_in ? lo + L : n_in;
                XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf(), cs = XYZZ<F>::inf();
                for (uint32_t k = hi; k-- > lo;) {
                    xyzz_add(run, Rin[(size_t)w * n_in + k]);
                    if (k != lo) xyzz_add(acc, run);
                    if (!Cin.empty()) xyzz_add(cs, Cin[(size_t)w * n_in + k]);
                }
                xyzz_dbl_k(acc, log2_M);
                xyzz_add(cs, acc);
                Rout[(size_t)w * n_out + u] = run;
                Cout[(size_t)w * n_out + u] = cs;
            }
        Rin.swap(Rout); Cin.swap(Cout);
        n_in = n_out;
        log2_M += 3;
        if (n_out == 1) break;
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = W - 1; w >= 0; --w) {
        xyzz_dbl_k(acc, c);
        XYZZ<F> s = Rin[w];
        xyzz_add(s, Cin[w]);
        xyzz_add(acc, s);
    }
    F o[3];
    if (acc.is_inf()) { o[0] = F::one(); o[1] = F::one(); o[2] = F::zero(); }
    else { Affine<F> a = xyzz_to_affine(acc); o[0] = a.x; o[1] = a.y; o[2] = F::one(); }
    memcpy(out_xyz, o, sizeof o);
}

extern "C" void emu_msm_g1(const uint64_t* bases_packed, const uint64_t* scalars, size_t n, int c, uint64_t* out_xyz) {
    emu_msm<Fq>(bases_packed, scalars, n, c, out_xyz);
}
extern "C" void emu_msm_g2(const uint64_t* bases_packed, const uint64_t* scalars, size_t n, int c, uint64_t* out_xyz) {
    emu_msm<Fq2>(bases_packed, scalars, n, c, out_xyz);
}
extern "C" int emu_pick_c(size_t n) { return msm_pick_c(n); }
extern "C" int emu_num_windows(int c) { return msm_num_windows(c); }
// signed digits of a canonical scalar -> int32 values (0 for none), for reconstruction checks
extern "C" void emu_digits(const uint64_t* scalar_canonical, int c, int32_t* out) {
    uint32_t s[8]; memcpy(s, scalar_canonical, 32);
    int W = msm_num_windows(c);
    std::vector<uint32_t> d(W);
    msm_signed_digits(s, c, W, d.data());
    for (int w = 0; w < W; ++w)
        out[w] = d[w] == MSM_DIGIT_NONE ? 0 : ((d[w] >> 31) ? -(int32_t)((d[w] & 0x7fffffffu) + 1) : (int32_t)((d[w] & 0x7fffffffu) + 1));
}
extern "C" void emu_fq2_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) stf(o + 8 * i, f_mul(ldf<Fq2>(a + 8 * i), ldf<Fq2>(b + 8 * i)));
}
extern "C" void emu_fq2_sqr(const uint64_t* a, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) stf(o + 8 * i, f_sqr(ldf<Fq2>(a + 8 * i)));
}
extern "C" void emu_fq2_inv(const uint64_t* a, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) stf(o + 8 * i, f_inv(ldf<Fq2>(a + 8 * i)));
}

// the carry-free radix-2^29 schedule (experiment kept in fp.cuh): must agree bit for bit
extern "C" void emu_fr_mul_r29(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_mul_r29(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
extern "C" void emu_fq_mul_r29(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_mul_r29(ld<Fq>(a + 4 * i), ld<Fq>(b + 4 * i)));
}
extern "C" int emu_pick_c_g2(size_t n) { return msm_pick_c(n, true); }
