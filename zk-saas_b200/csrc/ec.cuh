// ec.cuh -- BN254 G1 (over Fq) and G2 (over Fq2 = Fq[u]/(u^2+1)) group arithmetic for the MSM
// kernels.  Replaces ark-ec 0.4 `short_weierstrass::{Affine,Projective}` arithmetic underneath
// `VariableBaseMSM::msm` (reference call site dist-primitives/src/dmsm/mod.rs:73).
//
// Working representation is XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity: ZZ = 0): the
// mixed addition that dominates bucket accumulation costs 8M + 2S instead of the 7M + 4S of the
// Jacobian madd-2007-bl arkworks uses, and needs no field inversion.  A group element has a
// unique normalised affine form, so results are bit-identical to arkworks' after into_affine().
//
// Device-side affine bases are stored packed as (x, y) with the point at infinity encoded as
// (0, 0), which is not on either curve (b != 0).
#pragma once
#include "fp.cuh"

namespace zkg {

// --------------------------------------------------------------------------------------------
// uniform field interface: f_add/f_sub/f_mul/f_sqr/f_dbl/f_neg/f_inv over Fq and Fq2
// --------------------------------------------------------------------------------------------
template <class P> ZKG_D Fp<P> f_add(const Fp<P>& a, const Fp<P>& b) { return fp_add(a, b); }
template <class P> ZKG_D Fp<P> f_sub(const Fp<P>& a, const Fp<P>& b) { return fp_sub(a, b); }
template <class P> ZKG_D Fp<P> f_mul(const Fp<P>& a, const Fp<P>& b) { return fp_mul(a, b); }
template <class P> ZKG_D Fp<P> f_sqr(const Fp<P>& a) { return fp_sqr(a); }
// (tried for the cold formulas: an out-of-line shared multiplier body -- slower, 2.89 vs 2.67 ms tail;
//  the flag-free fp_mul_r29 so that ptxas may interleave independent products -- slower, 2.47 vs 1.35 ms)
template <class P> ZKG_D Fp<P> f_mul_hot(const Fp<P>& a, const Fp<P>& b) { return fp_mul(a, b); }
template <class P> ZKG_D Fp<P> f_sqr_hot(const Fp<P>& a) { return fp_sqr(a); }
// a*b - c*d with ONE Montgomery reduction (fp_dot: 192 wide MADs instead of 256); canonical, so it is
// bit-identical to f_sub(f_mul(a, b), f_mul(c, d)).  Every group formula ends its Y coordinate this way.
template <class P> ZKG_D Fp<P> f_mulsub(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    Fp<P> x[2] = {a, fp_neg(c)}, y[2] = {b, d};
    return fp_dot<P, 2>(x, y);
}
template <class P> ZKG_D Fp<P> f_dbl(const Fp<P>& a) { return fp_dbl(a); }
template <class P> ZKG_D Fp<P> f_neg(const Fp<P>& a) { return fp_neg(a); }
template <class P> ZKG_D Fp<P> f_inv(const Fp<P>& a) { return fp_inv(a); }

struct Fq2 {
    Fq c0, c1;
    ZKG_HD static Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
    ZKG_HD static Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
    ZKG_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    ZKG_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    ZKG_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
};
ZKG_D Fq2 f_add(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fp_add(a.c0, b.c0); r.c1 = fp_add(a.c1, b.c1); return r; }
ZKG_D Fq2 f_sub(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fp_sub(a.c0, b.c0); r.c1 = fp_sub(a.c1, b.c1); return r; }
ZKG_D Fq2 f_dbl(const Fq2& a) { Fq2 r; r.c0 = fp_dbl(a.c0); r.c1 = fp_dbl(a.c1); return r; }
ZKG_D Fq2 f_neg(const Fq2& a) { Fq2 r; r.c0 = fp_neg(a.c0); r.c1 = fp_neg(a.c1); return r; }
// (a0 + a1 u)(b0 + b1 u), u^2 = -1, Karatsuba: 3 base-field products
ZKG_NI Fq2 f_mul(const Fq2& a, const Fq2& b) {
    Fq v0 = fp_mul(a.c0, b.c0);
    Fq v1 = fp_mul(a.c1, b.c1);
    Fq s = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
    Fq2 r;
    r.c0 = fp_sub(v0, v1);
    r.c1 = fp_sub(fp_sub(s, v0), v1);
    return r;
}
// (a0 + a1 u)^2 = (a0+a1)(a0-a1) + 2 a0 a1 u : 2 base-field products
ZKG_NI Fq2 f_sqr(const Fq2& a) {
    Fq m = fp_mul(a.c0, a.c1);
    Fq2 r;
    r.c0 = fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1));
    r.c1 = fp_dbl(m);
    return r;
}
// a*b - c*d over Fq2 as two 4-term inner products: 640 wide MADs instead of the 768 of two Karatsuba products
//   re = a0 b0 - a1 b1 - c0 d0 + c1 d1,   im = a0 b1 + a1 b0 - c0 d1 - c1 d0
ZKG_NI Fq2 f_mulsub(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) {
    Fq nc0 = fp_neg(c.c0), nc1 = fp_neg(c.c1);
    Fq x[4] = {a.c0, fp_neg(a.c1), nc0, c.c1}, y[4] = {b.c0, b.c1, d.c0, d.c1};
    Fq2 r;
    r.c0 = fp_dot<FqParams, 4>(x, y);
    Fq x2[4] = {a.c0, a.c1, nc0, nc1}, y2[4] = {b.c1, b.c0, d.c1, d.c0};
    r.c1 = fp_dot<FqParams, 4>(x2, y2);
    return r;
}
ZKG_D Fq2 f_mul_hot(const Fq2& a, const Fq2& b) { return f_mul(a, b); }
ZKG_D Fq2 f_sqr_hot(const Fq2& a) { return f_sqr(a); }
ZKG_NI Fq2 f_inv(const Fq2& a) {
    Fq n = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
    Fq2 r;
    r.c0 = fp_mul(a.c0, n);
    r.c1 = fp_neg(fp_mul(a.c1, n));
    return r;
}

// --------------------------------------------------------------------------------------------
// points
// --------------------------------------------------------------------------------------------
template <class F>
struct Affine {
    F x, y;                                   // infinity <=> x == 0 && y == 0
    ZKG_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    ZKG_HD static Affine inf() { Affine r; r.x = F::zero(); r.y = F::zero(); return r; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    ZKG_HD bool is_inf() const { return zz.is_zero(); }
    ZKG_HD static XYZZ inf() { XYZZ r; r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero(); return r; }
    ZKG_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        XYZZ r; r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one(); return r;
    }
};

// acc = 2 * (affine p)      (mdbl-2008-s-1, a = 0)
template <class F>
ZKG_NI XYZZ<F> xyzz_dbl_affine(const Affine<F>& p) {
    if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();   // no 2-torsion on BN254, kept for safety
    XYZZ<F> r;
    F u = f_dbl(p.y);
    F v = f_sqr(u);
    F w = f_mul(u, v);
    F s = f_mul(p.x, v);
    F xx = f_sqr(p.x);
    F m = f_add(f_dbl(xx), xx);
    r.x = f_sub(f_sub(f_sqr(m), s), s);
    r.y = f_mulsub(m, f_sub(s, r.x), w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// a = 2a      (dbl-2008-s-1, a = 0)
template <class F>
ZKG_NI void xyzz_dbl(XYZZ<F>& a) {
    if (a.is_inf()) return;
    F u = f_dbl(a.y);
    F v = f_sqr(u);
    F w = f_mul(u, v);
    F s = f_mul(a.x, v);
    F xx = f_sqr(a.x);
    F m = f_add(f_dbl(xx), xx);
    F x3 = f_sub(f_sub(f_sqr(m), s), s);
    a.y = f_mulsub(m, f_sub(s, x3), w, a.y);
    a.x = x3;
    a.zz = f_mul(v, a.zz);
    a.zzz = f_mul(w, a.zzz);
}

// a = 2^k * a through Jacobian coordinates (dbl-2009-l, a = 0: 2M + 5S per doubling against 6M + 4S
// for XYZZ).  Used by the serial Horner tail of the MSM, whose c*(W-1) dependent doublings are pure
// latency.  XYZZ -> Jacobian without inversion: scale by lambda = zz, i.e. (zz^2 X, zz^3 Y, zzz).
template <class F>
ZKG_NI void xyzz_dbl_k(XYZZ<F>& a, int k) {
    if (a.is_inf() || k <= 0) return;
    F t = f_sqr(a.zz);
    F X = f_mul(a.x, t);
    F Y = f_mul(a.y, f_mul(t, a.zz));
    F Z = a.zzz;
    for (int i = 0; i < k; ++i) {
        F A = f_sqr(X);
        F B = f_sqr(Y);
        F C = f_sqr(B);
        F D = f_dbl(f_sub(f_sub(f_sqr(f_add(X, B)), A), C));
        F E = f_add(f_dbl(A), A);
        F X3 = f_sub(f_sub(f_sqr(E), D), D);
        F C8 = f_dbl(f_dbl(f_dbl(C)));
        Z = f_dbl(f_mul(Y, Z));
        Y = f_sub(f_mul(E, f_sub(D, X3)), C8);
        X = X3;
    }
    a.x = X;
    a.y = Y;
    a.zz = f_sqr(Z);
    a.zzz = f_mul(a.zz, Z);
}

// acc += (neg ? -p : p), p affine      (madd-2008-s: 8M + 2S)
template <class F>
ZKG_D void xyzz_madd(XYZZ<F>& acc, const Affine<F>& p_in, bool neg) {
    if (p_in.is_inf()) return;
    Affine<F> p = p_in;
    if (neg) p.y = f_neg(p.y);
    if (acc.is_inf()) { acc.x = p.x; acc.y = p.y; acc.zz = F::one(); acc.zzz = F::one(); return; }
    F u2 = f_mul_hot(p.x, acc.zz);
    F s2 = f_mul_hot(p.y, acc.zzz);
    F pp_ = f_sub(u2, acc.x);
    F r = f_sub(s2, acc.y);
    if (pp_.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl_affine(p);     // P == Q
        else acc = XYZZ<F>::inf();                     // P == -Q
        return;
    }
    F pp = f_sqr_hot(pp_);
    F ppp = f_mul_hot(pp_, pp);
    F q = f_mul_hot(acc.x, pp);
    F x3 = f_sub(f_sub(f_sub(f_sqr_hot(r), ppp), q), q);
    acc.y = f_mulsub(r, f_sub(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.zz = f_mul_hot(acc.zz, pp);
    acc.zzz = f_mul_hot(acc.zzz, ppp);
}

// a += b      (add-2008-s: 12M + 2S)
template <class F>
ZKG_NI void xyzz_add(XYZZ<F>& a, const XYZZ<F>& b) {
    if (b.is_inf()) return;
    if (a.is_inf()) { a = b; return; }
    F u1 = f_mul(a.x, b.zz);
    F u2 = f_mul(b.x, a.zz);
    F s1 = f_mul(a.y, b.zzz);
    F s2 = f_mul(b.y, a.zzz);
    F pp_ = f_sub(u2, u1);
    F r = f_sub(s2, s1);
    if (pp_.is_zero()) {
        if (r.is_zero()) xyzz_dbl(a);
        else a = XYZZ<F>::inf();
        return;
    }
    F pp = f_sqr(pp_);
    F ppp = f_mul(pp_, pp);
    F q = f_mul(u1, pp);
    F x3 = f_sub(f_sub(f_sub(f_sqr(r), ppp), q), q);
    a.y = f_mulsub(r, f_sub(q, x3), s1, ppp);
    a.x = x3;
    a.zz = f_mul(f_mul(a.zz, b.zz), pp);
    a.zzz = f_mul(f_mul(a.zzz, b.zzz), ppp);
}

// unique normal form (into_affine): x = X/ZZ, y = Y/ZZZ with a single inversion
template <class F>
ZKG_NI Affine<F> xyzz_to_affine(const XYZZ<F>& a) {
    if (a.is_inf()) return Affine<F>::inf();
    F d = f_inv(f_mul(a.zz, a.zzz));
    Affine<F> r;
    r.x = f_mul(a.x, f_mul(a.zzz, d));
    r.y = f_mul(a.y, f_mul(a.zz, d));
    return r;
}

// (Round-1 experiment, removed: "quad-cooperative" add/double with four lanes evaluating the
// independent products of a formula level and width-4 shuffles between levels.  A lone warp runs
// this code at ~6.3 cycles per instruction (ncu: 3.2 wait + 1.5 branch-resolving stalls per issue),
// so the select/shuffle overhead of ~250 instructions per level cost as much as the products it
// saved: k_final 1.55 ms vs 1.37 ms single-lane.)


typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

}  // namespace zkg
