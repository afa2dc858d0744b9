// fp.cuh -- 254-bit prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form
// (R = 2^256), i.e. bit-identical to ark-ff 0.4 `Fp<MontBackend<_,4>,4>` memory images
// (4 x u64 little-endian == 8 x u32 little-endian).  Replaces the ark-ff arithmetic under
// every hot function of the reference.
//
// The multiplier runs on the integer-MAD pipe: products are formed as 32x32->64 multiply-adds
// with hardware carry chains (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into
// IMAD.WIDE.U32[.X]).  Partial products of even and odd limbs are kept in two accumulators that
// are 32 bits out of phase, so every wide MAD lands on a register pair and the Montgomery
// reduction is interleaved row by row (CIOS).  Both BN254 moduli are < 2^254, which is what lets
// the per-row top carries be folded with a single addc (see the bounds in mont_mul).
//
// The same source compiles for the host when ZKG_HOST_EMU is defined: the PTX carry-flag
// primitives are then emulated in portable C++ so that the exact limb schedule can be checked
// on a machine without a GPU (tests/test_host_emulation.py).
#pragma once
#include <stdint.h>
#include "bn254_consts.cuh"

#ifdef ZKG_HOST_EMU
#define ZKG_HD inline
#define ZKG_D inline
#define ZKG_NI inline
#else
#define ZKG_HD __host__ __device__ __forceinline__
#define ZKG_D __device__ __forceinline__
// out-of-line device function: used for everything that is big but not on the innermost hot
// path, so that ptxas sees kernels of a few thousand instructions instead of a few hundred thousand
#define ZKG_NI static __device__ __noinline__
#endif

namespace zkg {

// --------------------------------------------------------------------------------------------
// carry-chain primitives
// --------------------------------------------------------------------------------------------
#ifdef ZKG_HOST_EMU
namespace emu { static thread_local uint32_t cf = 0; }
// d = a + b (sets CF)
ZKG_D uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
ZKG_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + emu::cf; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
ZKG_D uint32_t addc(uint32_t a, uint32_t b) { return a + b + emu::cf; }
ZKG_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; emu::cf = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
ZKG_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - emu::cf; emu::cf = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
ZKG_D uint32_t subc(uint32_t a, uint32_t b) { return a - b - emu::cf; }
ZKG_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ZKG_D uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
ZKG_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
ZKG_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
ZKG_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
ZKG_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
ZKG_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + emu::cf; }
// 32x32->64 multiply-accumulate pairs (one IMAD.WIDE each on the device)
ZKG_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
ZKG_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = mad_lo_cc(a, b, lo); hi = madc_hi_cc(a, b, hi); }
ZKG_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = madc_lo_cc(a, b, lo); hi = madc_hi_cc(a, b, hi); }
ZKG_D void madc_wide_cc_from(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) { lo = madc_lo_cc(a, b, c0); hi = madc_hi_cc(a, b, c1); }
#else
// NOTE on sub.cc: PTX defines CC.CF after sub.cc as the *borrow* (1 = borrow occurred), and
// subc consumes it as a borrow; the emulation above follows the same convention.
ZKG_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZKG_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZKG_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZKG_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZKG_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZKG_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZKG_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ZKG_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
ZKG_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZKG_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZKG_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZKG_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZKG_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// 32x32->64 multiply-accumulate pairs.  The lo/hi halves MUST sit in one asm statement: that is
// the pattern ptxas fuses into a single full-rate IMAD.WIDE.U32[.X]; issued as separate statements
// it emits IMAD + IMAD.HI (half rate on sm_100, measured) + 2 IADD3.X instead.
ZKG_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
ZKG_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
ZKG_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
ZKG_D void madc_wide_cc_from(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c0), "r"(c1));
}
#endif

// --------------------------------------------------------------------------------------------
// Field parameter packs (limbs from the generated bn254_consts.cuh)
// --------------------------------------------------------------------------------------------
struct FrParams {
    static constexpr uint32_t INV = BN254_FR_INV;
    static constexpr uint32_t INV29 = BN254_FR_INV29;
    ZKG_HD static constexpr uint32_t mod29(int i) { return BN254_FR_MOD29_L(i); }
    ZKG_HD static constexpr uint32_t mod(int i) { return BN254_FR_MOD_L(i); }
    ZKG_HD static constexpr uint32_t one(int i) { return BN254_FR_R_L(i); }
    ZKG_HD static constexpr uint32_t r2(int i) { return BN254_FR_R2_L(i); }
};
struct FqParams {
    static constexpr uint32_t INV = BN254_FQ_INV;
    static constexpr uint32_t INV29 = BN254_FQ_INV29;
    ZKG_HD static constexpr uint32_t mod29(int i) { return BN254_FQ_MOD29_L(i); }
    ZKG_HD static constexpr uint32_t mod(int i) { return BN254_FQ_MOD_L(i); }
    ZKG_HD static constexpr uint32_t one(int i) { return BN254_FQ_R_L(i); }
    ZKG_HD static constexpr uint32_t r2(int i) { return BN254_FQ_R2_L(i); }
};

template <class P>
struct Fp {
    typedef P Params;
    static constexpr int N = 8;
    uint32_t v[N];

    ZKG_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = 0;
        return r;
    }
    ZKG_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = P::one(i);
        return r;
    }
    ZKG_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = P::r2(i);
        return r;
    }
    ZKG_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    ZKG_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    ZKG_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - p if a >= p else a   (a < 2p)
template <class P>
ZKG_D void final_sub(uint32_t* a) {
    uint32_t t[8];
    t[0] = sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) t[i] = subc_cc(a[i], P::mod(i));
    uint32_t borrow = subc(0, 0);   // 0 - 0 - CF : 0 if no borrow, 0xffffffff if borrow
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = borrow ? a[i] : t[i];
}

template <class P>
ZKG_D Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.v[i] = addc_cc(a.v[i], b.v[i]);
    r.v[7] = addc(a.v[7], b.v[7]);     // a + b < 2^255: no carry out
    final_sub<P>(r.v);
    return r;
}

template <class P>
ZKG_D Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) r.v[i] = subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = subc(0, 0);
    // add p back under the borrow mask
    r.v[0] = add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.v[i] = addc_cc(r.v[i], P::mod(i) & borrow);
    r.v[7] = addc(r.v[7], P::mod(7) & borrow);
    return r;
}

template <class P>
ZKG_D Fp<P> fp_neg(const Fp<P>& a) {
    // p - a, with 0 -> 0
    Fp<P> r;
    r.v[0] = sub_cc(P::mod(0), a.v[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.v[i] = subc_cc(P::mod(i), a.v[i]);
    r.v[7] = subc(P::mod(7), a.v[7]);
    uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] &= nz;
    return r;
}

template <class P>
ZKG_D Fp<P> fp_dbl(const Fp<P>& a) { return fp_add(a, a); }

// ---- wide-MAD row helpers ------------------------------------------------------------------
// acc[j], acc[j+1] = a[j] * b   for j = 0,2,4,6          (4 x 32x32->64 multiplies)
ZKG_D void mul_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) mul_wide(acc[j], acc[j + 1], a[j], b);
}
// acc += (a[0],a[2],a[4],a[6]) * b as one carry chain; leaves the carry-out in CF
ZKG_D void cmad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
    mad_wide_cc(acc[0], acc[1], a[0], b);
#pragma unroll
    for (int j = 2; j < 8; j += 2) madc_wide_cc(acc[j], acc[j + 1], a[j], b);
}
// acc = (acc >> 64) + (a[0],a[2],a[4],a[6]) * b, consuming CF as carry-in; carry-out is 0
// (the top pair only receives a[6]*b, and a[6] < 2^30 for both BN254 moduli)
ZKG_D void madc_row_rshift(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) madc_wide_cc_from(acc[j], acc[j + 1], a[j], b, acc[j + 2], acc[j + 3]);
    madc_wide_cc_from(acc[6], acc[7], a[6], b, 0, 0);
}

// One CIOS row.  `x` is the accumulator whose limb k sits on column k, `y` the accumulator that is
// one limb out of phase.  Adds a*bi, then m*p with m chosen so that column 0 cancels; the caller
// swaps the roles of x and y for the next row (that swap is the division by 2^32).
template <class P, bool FIRST>
ZKG_D void mont_row(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t bi) {
    uint32_t pm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pm[i] = P::mod(i);
    if (FIRST) {
        mul_row(y, a + 1, bi);
        mul_row(x, a, bi);
    } else {
        x[0] = add_cc(x[0], y[1]);          // the limb that falls off y when it is shifted by 64 bits
        madc_row_rshift(y, a + 1, bi);
        cmad_row(x, a, bi);
        y[7] = addc(y[7], 0);
    }
    uint32_t m = mul_lo(x[0], P::INV);
    cmad_row(y, pm + 1, m);                  // top pair: a7*bi + p7*m < 2^63, carry-out is 0
    cmad_row(x, pm, m);
    y[7] = addc(y[7], 0);
}

// r = a * b * 2^-256 mod p, fully reduced -- carry-chain (CIOS) schedule: 128 IMAD.WIDE.U32[.X].
// This is the production multiplier: 67 G products/s on B200 (tools/microbench/intpipe.cu), 93 % of the
// 9.27e12/s at which the 32x32->64 multiplier issues in any form (tools/microbench/widemad.cu).
template <class P>
ZKG_D Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
    uint32_t even[8], odd[8];
    mont_row<P, true>(even, odd, a.v, b.v[0]);
    mont_row<P, false>(odd, even, a.v, b.v[1]);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        mont_row<P, false>(even, odd, a.v, b.v[i]);
        mont_row<P, false>(odd, even, a.v, b.v[i + 1]);
    }
    // merge the two out-of-phase accumulators: result limb k = even[k] + odd[k+1]
    Fp<P> r;
    r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.v[i] = addc_cc(even[i], odd[i + 1]);
    r.v[7] = addc(even[7], 0);
    final_sub<P>(r.v);
    return r;
}


// --------------------------------------------------------------------------------------------
// Montgomery dot product  r = (a_0*b_0 + ... + a_(K-1)*b_(K-1)) * 2^-256 mod p,  K <= 4.
// Same two-accumulator CIOS rows as fp_mul, but every row adds the K partial products before the
// ONE reduction row, so a K-term inner product costs 8*(8K + 8) wide MADs instead of 8*16K: 0.75x
// at K = 2, 0.625x at K = 4 -- the constant-matrix maps of the PSS transforms (unpack2, pack,
// Lagrange) are made of these.  Bounds: with a_k < p and the running value t < (K+1) p, a row
// leaves t' = (t + sum_k a_k b_k[i] + m p) / 2^32 < (K+1) p again, and (K+1) p < 2^256 for K <= 4
// (5 p = 0.947 * 2^256 for both BN254 moduli), so no column beyond the ninth is ever needed: chains
// into the out-of-phase accumulator cannot carry out, chains into the in-phase one hand their
// carry to its top limb exactly as in fp_mul.  K final conditional subtractions give the
// canonical residue -- bit-identical to summing K fp_mul results.
// --------------------------------------------------------------------------------------------
template <class P, int K, bool FIRST>
ZKG_D void mont_row_dot(uint32_t* x, uint32_t* y, const Fp<P>* a, const Fp<P>* b, int i) {
    uint32_t pm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pm[j] = P::mod(j);
    if (FIRST) {
        mul_row(y, a[0].v + 1, b[0].v[i]);
        mul_row(x, a[0].v, b[0].v[i]);
    } else {
        x[0] = add_cc(x[0], y[1]);
        madc_row_rshift(y, a[0].v + 1, b[0].v[i]);
        cmad_row(x, a[0].v, b[0].v[i]);
        y[7] = addc(y[7], 0);
    }
#pragma unroll
    for (int k = 1; k < K; ++k) {
        cmad_row(y, a[k].v + 1, b[k].v[i]);      // cannot carry out (value bound above)
        cmad_row(x, a[k].v, b[k].v[i]);
        y[7] = addc(y[7], 0);
    }
    uint32_t m = mul_lo(x[0], P::INV);
    cmad_row(y, pm + 1, m);
    cmad_row(x, pm, m);
    y[7] = addc(y[7], 0);
}

template <class P, int K>
ZKG_D Fp<P> fp_dot(const Fp<P>* a, const Fp<P>* b) {
    static_assert(K >= 1 && K <= 4, "fp_dot: (K+1) p must stay below 2^256");
    uint32_t even[8], odd[8];
    mont_row_dot<P, K, true>(even, odd, a, b, 0);
    mont_row_dot<P, K, false>(odd, even, a, b, 1);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        mont_row_dot<P, K, false>(even, odd, a, b, i);
        mont_row_dot<P, K, false>(odd, even, a, b, i + 1);
    }
    Fp<P> r;
    r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.v[i] = addc_cc(even[i], odd[i + 1]);
    r.v[7] = addc(even[7], 0);
#pragma unroll
    for (int k = 0; k < K; ++k) final_sub<P>(r.v);
    return r;
}

// --------------------------------------------------------------------------------------------
// EXPERIMENT (not used by the kernels; kept with its measurement because it decides the design):
// carry-free multiplier.  The idea was that wide MADs without carry might issue faster than the carry-propagating
// form, so that an unsaturated radix-2^29 product (162 carry-free MADs) would beat the CIOS schedule (128).  It does
// not, and cannot: IMAD.WIDE.U32 issues at 9.3e12/s in EVERY form on B200 (tools/microbench/widemad.cu; the 18.4e12/s
// first quoted here was a loop ptxas had hoisted), ptxas splits every `mad.wide` accumulate into IMAD.WIDE(+RZ) +
// IADD3/IADD3.X, and the re-slicing adds ~150 ALU instructions.  Measured: 40 G products/s vs 67 G/s for CIOS
// (profiles/r01_intpipe_microbench_v2.json).  The product is formed on
// an UNSATURATED radix-2^29 view of the operands: 9 limbs of 29 bits, 58-bit partial products, and
// every column of the product (<= 18 terms + carry < 2^63) accumulates in one 64-bit register pair
// with plain wide MADs -- no carry flags anywhere.  Montgomery reduction is interleaved column by
// column (product scanning).  Eight columns retire 29 bits each and the ninth retires 24 bits, 256
// in total, so the result is a*b*2^-256 mod p: bit-identical to the CIOS schedule above and to
// arkworks.  The 8x32 <-> 9x29 re-slicing is funnel shifts on the (otherwise idle) ALU pipe.
// --------------------------------------------------------------------------------------------
static constexpr uint32_t MASK29 = (1u << 29) - 1;

ZKG_D uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
#ifdef ZKG_HOST_EMU
    return c + (uint64_t)a * b;
#else
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
#endif
}
// bits [lo, lo+32) of the 256-bit value v (8 x u32), zero above bit 255
ZKG_D uint32_t bits32_at(const uint32_t* v, int lo) {
    int w = lo >> 5, s = lo & 31;
    uint32_t a = w < 8 ? v[w] : 0u, b = w + 1 < 8 ? v[w + 1] : 0u;
#ifdef ZKG_HOST_EMU
    return s ? (a >> s) | (b << (32 - s)) : a;
#else
    return __funnelshift_r(a, b, s);
#endif
}
// 8 x 32-bit limbs -> 9 x 29-bit limbs (value < 2^256, so the top limb has 24 bits)
ZKG_D void split29(const uint32_t* v, uint32_t* x) {
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = bits32_at(v, 29 * i) & MASK29;
}

template <class P>
ZKG_D Fp<P> fp_mul_r29(const Fp<P>& a, const Fp<P>& b) {
    uint32_t x[9], y[9], m[9], u[10];
    split29(a.v, x);
    split29(b.v, y);
    uint64_t carry = 0;
    // columns 0..8: accumulate a*b and m*p, pick m[k] so that the column's low bits cancel
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i <= k; ++i) c = mad_wide(x[i], y[k - i], c);
        uint64_t d = carry;
#pragma unroll
        for (int i = 0; i < k; ++i) d = mad_wide(m[i], P::mod29(k - i), d);
        d += c;
        if (k < 8) {
            m[k] = ((uint32_t)d * P::INV29) & MASK29;
            d = mad_wide(m[k], P::mod29(0), d);          // low 29 bits are now zero
            carry = d >> 29;
        } else {
            m[8] = ((uint32_t)d * P::INV29) & ((1u << 24) - 1);
            d = mad_wide(m[8], P::mod29(0), d);          // low 24 bits are now zero: 8*29 + 24 = 256 retired
            u[0] = (uint32_t)d & MASK29;
            carry = d >> 29;
        }
    }
    // columns 9..16
#pragma unroll
    for (int k = 9; k < 17; ++k) {
        uint64_t c = 0;
#pragma unroll
        for (int i = k - 8; i <= 8; ++i) c = mad_wide(x[i], y[k - i], c);
        uint64_t d = carry;
#pragma unroll
        for (int i = k - 8; i <= 8; ++i) d = mad_wide(m[i], P::mod29(k - i), d);
        d += c;
        u[k - 8] = (uint32_t)d & MASK29;
        carry = d >> 29;
    }
    u[9] = (uint32_t)carry;
    // result = (sum_j u[j] 2^(29 j)) >> 24, re-sliced into 8 x 32 bits; it is < 2p < 2^255
    Fp<P> r;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        // bits [32w + 24, 32w + 56) of U
        int lo = 32 * w + 24;
        int j = lo / 29, s = lo % 29;                    // starts inside limb j at bit s
        uint64_t t = (uint64_t)u[j] >> s;
        int have = 29 - s;
        if (j + 1 < 10) { t |= (uint64_t)u[j + 1] << have; have += 29; }
        if (have < 32 && j + 2 < 10) t |= (uint64_t)u[j + 2] << have;
        r.v[w] = (uint32_t)t;
    }
    final_sub<P>(r.v);
    return r;
}

// --------------------------------------------------------------------------------------------
// Dedicated squaring: 36 + 64 wide MADs instead of 128.
//   a^2 = sum_i a_i * 2^(32 i) * ( a_i * 2^(32 i) + 2 * A_i ),   A_i = sum_{j > i} a_j 2^(32 j).
// The doubling is folded into the multiplicand: limb j of 2 A_i is d[j] = limb j of 2a for j > i+1 and
// e[i+1] = a[i+1] << 1 for j = i+1 (the bit that 2a carries in from a[i] belongs to no A_i); 2a < 2^256
// because a is canonical (< 2^254).  The triangular product is formed first (two accumulators, E on even
// limb positions and O on odd ones, so that every wide MAD lands on an aligned register pair; row i is one
// carry chain in each), then reduced word by word: the rows of the reduction are the m*p halves of the
// CIOS rows above, and the upper limbs T[8..15] enter through the addend of the top MAD of each row.
// Bit-identical to fp_mul(a, a).
// --------------------------------------------------------------------------------------------
template <class P>
ZKG_D Fp<P> fp_sqr(const Fp<P>& a_) {
    const uint32_t* a = a_.v;
    uint32_t d[8], e[8];
#pragma unroll
    for (int j = 1; j < 8; ++j) { d[j] = (a[j] << 1) | (a[j - 1] >> 31); e[j] = a[j] << 1; }
    uint32_t E[16], O[16];        // E[k] <-> limb k;  O[k] <-> limb k + 1
    // row 0: fresh pairs
    mul_wide(E[0], E[1], a[0], a[0]); mul_wide(E[2], E[3], a[0], d[2]); mul_wide(E[4], E[5], a[0], d[4]); mul_wide(E[6], E[7], a[0], d[6]);
    mul_wide(O[0], O[1], a[0], e[1]); mul_wide(O[2], O[3], a[0], d[3]); mul_wide(O[4], O[5], a[0], d[5]); mul_wide(O[6], O[7], a[0], d[7]);
    // row 1
    mad_wide_cc(E[2], E[3], a[1], a[1]); madc_wide_cc(E[4], E[5], a[1], d[3]); madc_wide_cc(E[6], E[7], a[1], d[5]);
    madc_wide_cc_from(E[8], E[9], a[1], d[7], 0, 0);
    mad_wide_cc(O[2], O[3], a[1], e[2]); madc_wide_cc(O[4], O[5], a[1], d[4]); madc_wide_cc(O[6], O[7], a[1], d[6]);
    O[8] = addc(0, 0); O[9] = 0;
    // row 2
    mad_wide_cc(E[4], E[5], a[2], a[2]); madc_wide_cc(E[6], E[7], a[2], d[4]); madc_wide_cc(E[8], E[9], a[2], d[6]);
    E[10] = addc(0, 0); E[11] = 0;
    mad_wide_cc(O[4], O[5], a[2], e[3]); madc_wide_cc(O[6], O[7], a[2], d[5]); madc_wide_cc(O[8], O[9], a[2], d[7]);
    O[10] = addc(0, 0); O[11] = 0;
    // row 3
    mad_wide_cc(E[6], E[7], a[3], a[3]); madc_wide_cc(E[8], E[9], a[3], d[5]); madc_wide_cc(E[10], E[11], a[3], d[7]);
    E[12] = addc(0, 0); E[13] = 0;
    mad_wide_cc(O[6], O[7], a[3], e[4]); madc_wide_cc(O[8], O[9], a[3], d[6]);
    O[10] = addc_cc(O[10], 0); O[11] = addc(O[11], 0);
    // row 4
    mad_wide_cc(E[8], E[9], a[4], a[4]); madc_wide_cc(E[10], E[11], a[4], d[6]);
    E[12] = addc_cc(E[12], 0); E[13] = addc(E[13], 0);
    mad_wide_cc(O[8], O[9], a[4], e[5]); madc_wide_cc(O[10], O[11], a[4], d[7]);
    O[12] = addc(0, 0); O[13] = 0;
    // row 5
    mad_wide_cc(E[10], E[11], a[5], a[5]); madc_wide_cc(E[12], E[13], a[5], d[7]);
    E[14] = addc(0, 0); E[15] = 0;
    mad_wide_cc(O[10], O[11], a[5], e[6]);
    O[12] = addc_cc(O[12], 0); O[13] = addc(O[13], 0);
    // row 6
    mad_wide_cc(E[12], E[13], a[6], a[6]);
    E[14] = addc_cc(E[14], 0); E[15] = addc(E[15], 0);
    mad_wide_cc(O[12], O[13], a[6], e[7]);
    O[14] = addc(0, 0);
    // row 7 (a^2 < 2^508: no carry out)
    mad_wide_cc(E[14], E[15], a[7], a[7]);
    // T = E + O * 2^32
    uint32_t T[16];
    T[0] = E[0];
    T[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 15; ++k) T[k] = addc_cc(E[k], O[k - 1]);
    T[15] = addc(E[15], O[14]);

    // word-by-word Montgomery reduction of T (8 rows; x is the accumulator aligned on the limb being cleared)
    uint32_t pm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pm[i] = P::mod(i);
    uint32_t even[8], odd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) even[k] = T[k];
    {
        uint32_t m = mul_lo(even[0], P::INV);
        mul_row(odd, pm + 1, m);
        cmad_row(even, pm, m);
        odd[7] = addc(odd[7], 0);
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        uint32_t* x = (i & 1) ? odd : even;      // aligned on the limb this row clears
        uint32_t* y = (i & 1) ? even : odd;      // last row's x: y[0] == 0, y[1] falls onto x[0]
        x[0] = add_cc(x[0], y[1]);
        uint32_t m = mul_lo(x[0], P::INV);
#pragma unroll
        for (int j = 0; j < 6; j += 2) madc_wide_cc_from(y[j], y[j + 1], pm[j + 1], m, y[j + 2], y[j + 3]);
        madc_wide_cc_from(y[6], y[7], pm[7], m, T[7 + i], i == 7 ? T[15] : 0u);
        cmad_row(x, pm, m);
        y[7] = addc(y[7], 0);
    }
    // after row 7 (x = odd): result limb k = even[k] + odd[k + 1]
    Fp<P> r;
    r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.v[i] = addc_cc(even[i], odd[i + 1]);
    r.v[7] = addc(even[7], 0);
    final_sub<P>(r.v);
    return r;
}

// Montgomery -> canonical (ark-ff into_bigint): multiply by 1
template <class P>
ZKG_D Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> o = Fp<P>::zero();
    o.v[0] = 1;
    return fp_mul(a, o);
}
template <class P>
ZKG_D Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul(a, Fp<P>::r2()); }

// a^e for a canonical 256-bit exponent given as 8 limbs (used for inversion / roots)
template <class P>
ZKG_NI Fp<P> fp_pow(const Fp<P>& a, const uint32_t* e, int nlimbs) {
    Fp<P> acc = Fp<P>::one();
    for (int i = nlimbs - 1; i >= 0; --i)
        for (int b = 31; b >= 0; --b) {
            acc = fp_sqr(acc);
            if ((e[i] >> b) & 1) acc = fp_mul(acc, a);
        }
    return acc;
}

// Fermat inversion a^(p-2); 0 -> 0.  Uniform control flow (384 multiplications): kept as the
// cross-check of fp_inv below in the host-emulation tests.
template <class P>
ZKG_NI Fp<P> fp_inv_fermat(const Fp<P>& a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = P::mod(i);
    e[0] -= 2;   // both moduli are odd and > 2: low limb does not borrow
    return fp_pow(a, e, 8);
}

// 256-bit helpers for the binary inversion (plain C++: add/sub/shift chains, no multiplier)
ZKG_HD bool big_lt(const uint32_t* a, const uint32_t* b) {
    bool lt = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) lt = a[i] < b[i] || (a[i] == b[i] && lt);     // most significant limb decides last
    return lt;
}
ZKG_HD void big_sub(uint32_t* a, const uint32_t* b) {          // a -= b   (a >= b)
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        a[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
}
template <class P>
ZKG_HD void big_add_p(uint32_t* a) {                            // a += p   (a < p < 2^254: no carry out)
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a[i] + P::mod(i);
        a[i] = (uint32_t)c;
        c >>= 32;
    }
}
ZKG_HD void big_shr1(uint32_t* a) {
#pragma unroll
    for (int i = 0; i < 7; ++i) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] >>= 1;
}
ZKG_HD bool big_is_one(const uint32_t* a) {
    uint32_t o = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; ++i) o |= a[i];
    return o == 0;
}

// Inversion by the binary extended Euclidean algorithm (HAC 14.61 with both cofactors kept mod p):
// ~2*254 halving/subtraction steps of 8-limb add/shift chains, about a tenth of the instructions of
// the Fermat ladder.  The callers are the normalisations at the end of an MSM (one lone thread,
// where the ladder's 384 dependent multiplications cost ~0.15 ms) and the per-point to-affine of the
// table/CRS preparation kernels.  Data-dependent control flow; the result is the unique canonical
// inverse either way.  0 -> 0.
template <class P>
ZKG_NI Fp<P> fp_inv(const Fp<P>& a) {
    if (a.is_zero()) return a;
    uint32_t u[8], v[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = a.v[i]; v[i] = P::mod(i); b[i] = 0; c[i] = 0; }
    b[0] = 1;
    // invariants:  b * A == u,  c * A == v   (mod p),  A = the integer held in a (= a_std * R)
    while (!big_is_one(u) && !big_is_one(v)) {
        while ((u[0] & 1u) == 0) {
            big_shr1(u);
            if (b[0] & 1u) big_add_p<P>(b);
            big_shr1(b);
        }
        while ((v[0] & 1u) == 0) {
            big_shr1(v);
            if (c[0] & 1u) big_add_p<P>(c);
            big_shr1(c);
        }
        if (!big_lt(u, v)) {
            big_sub(u, v);
            if (big_lt(b, c)) big_add_p<P>(b);
            big_sub(b, c);
        } else {
            big_sub(v, u);
            if (big_lt(c, b)) big_add_p<P>(c);
            big_sub(c, b);
        }
    }
    Fp<P> r;
    const bool from_b = big_is_one(u);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = from_b ? b[i] : c[i];
    // r = A^-1 = a_std^-1 * R^-1 ;  two Montgomery products by R^2 give a_std^-1 * R
    r = fp_mul(r, Fp<P>::r2());
    return fp_mul(r, Fp<P>::r2());
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace zkg
