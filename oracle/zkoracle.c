/* oracle/zkoracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, gcc) of the zk-SaaS hot path over BN254:
 *   - the arkworks arithmetic the reference calls (ark-ff Fp, ark-ec short Weierstrass +
 *     VariableBaseMSM, ark-poly Radix2EvaluationDomain; versions ^0.4, un-vendored and
 *     un-pinned: dist-primitives/Cargo.toml:9-14, secret-sharing/Cargo.toml:7-10, no Cargo.lock),
 *   - literal loops of the reference's own code: secret-sharing/src/pss.rs,
 *     secret-sharing/src/utils.rs, dist-primitives/src/dfft/mod.rs,
 *     dist-primitives/src/utils/{pack,deg_red}.rs, dist-primitives/src/dmsm/mod.rs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / the reported CPU baseline.  The product
 * (libzksaas_gpu.so) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" at the arkworks binary (no Rust toolchain here, no golden
 * vectors in the reference).  This restatement is pinned instead against (i) oracle/pyref.py, an
 * independent big-integer model, (ii) the BN254 constants / on-curve points held in the
 * reference tree (fixtures/verifier.sol, fixtures/verification_key.json) and (iii) the
 * equalities the reference's own tests assert.  See tests/test_oracle_pins.py.
 *
 * Memory images are arkworks' in-memory forms: Fp = 4 LE u64 Montgomery limbs (R = 2^256);
 * G1 affine = {x, y, infinity:bool} (72 B), G2 affine = {x.c0,x.c1,y.c0,y.c1,infinity} (136 B);
 * group results are Jacobian (X,Y,Z) images normalised to Z = 1 (identity = (1,1,0)).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "bn254_consts.h"

typedef unsigned __int128 u128;

#define FP fr
#define FP_MOD BN254_FR_MOD
#define FP_R BN254_FR_R
#define FP_R2 BN254_FR_R2
#define FP_INV BN254_FR_INV
#include "fp_tmpl.h"

#define FP fq
#define FP_MOD BN254_FQ_MOD
#define FP_R BN254_FQ_R
#define FP_R2 BN254_FQ_R2
#define FP_INV BN254_FQ_INV
#include "fp_tmpl.h"

/* ---------------------------------------------------------------------------------------
 * Fq2 = Fq[u]/(u^2 + 1)  (ark-bn254 Fq2Config: NONRESIDUE = -1); element = c0 || c1 (8 u64)
 * ------------------------------------------------------------------------------------- */
static inline void fq2_set(uint64_t *o, const uint64_t *a) { memcpy(o, a, 64); }
static inline void fq2_zero(uint64_t *o) { memset(o, 0, 64); }
static inline void fq2_one(uint64_t *o) { fq_one(o); fq_zero(o + 4); }
static inline int fq2_is_zero(const uint64_t *a) { return fq_is_zero(a) && fq_is_zero(a + 4); }
static inline int fq2_eq(const uint64_t *a, const uint64_t *b) { return memcmp(a, b, 64) == 0; }
static inline void fq2_add(uint64_t *o, const uint64_t *a, const uint64_t *b) { fq_add(o, a, b); fq_add(o + 4, a + 4, b + 4); }
static inline void fq2_sub(uint64_t *o, const uint64_t *a, const uint64_t *b) { fq_sub(o, a, b); fq_sub(o + 4, a + 4, b + 4); }
static inline void fq2_neg(uint64_t *o, const uint64_t *a) { fq_neg(o, a); fq_neg(o + 4, a + 4); }
static inline void fq2_dbl(uint64_t *o, const uint64_t *a) { fq2_add(o, a, a); }
static inline void fq2_mul(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    /* Karatsuba: 3 base multiplications */
    uint64_t v0[4], v1[4], s[4], t[4];
    fq_mul(v0, a, b);
    fq_mul(v1, a + 4, b + 4);
    fq_add(s, a, a + 4);
    fq_add(t, b, b + 4);
    fq_mul(s, s, t);
    fq_sub(s, s, v0);
    fq_sub(o + 4, s, v1);
    fq_sub(o, v0, v1);
}
static inline void fq2_sqr(uint64_t *o, const uint64_t *a) {
    /* (a0+a1)(a0-a1), 2 a0 a1 */
    uint64_t s[4], d[4], m[4];
    fq_add(s, a, a + 4);
    fq_sub(d, a, a + 4);
    fq_mul(m, a, a + 4);
    fq_mul(o, s, d);
    fq_dbl(o + 4, m);
}
static void fq2_inv(uint64_t *o, const uint64_t *a) {
    uint64_t n[4], t[4];
    fq_sqr(n, a);
    fq_sqr(t, a + 4);
    fq_add(n, n, t);
    fq_inv(n, n);
    fq_mul(o, a, n);
    fq_mul(t, a + 4, n);
    fq_neg(o + 4, t);
}

/* ark_std::log2 = ceil(log2 x) */
static int ceil_log2(size_t x) {
    int r = 0;
    while (((size_t)1 << r) < x) ++r;
    return r;
}

/* ark-ec 0.4.2: c = size < 32 ? 3 : ln_without_floats(size) + 2, ln_without_floats(a) = log2(a)*69/100 */
static int ark_window_size(size_t size) {
    if (size < 32) return 3;
    return ceil_log2(size) * 69 / 100 + 2;
}

/* ark-ec 0.4.2 make_digits: signed radix-2^c digits of a canonical 256-bit scalar */
static void make_digits(const uint64_t *a, int w, int num_bits, int64_t *out) {
    uint64_t radix = (uint64_t)1 << w;
    uint64_t window_mask = radix - 1;
    uint64_t carry = 0;
    int digits_count = (num_bits + w - 1) / w;
    for (int i = 0; i < digits_count; ++i) {
        size_t bit_offset = (size_t)i * w;
        size_t u64_idx = bit_offset / 64;
        size_t bit_idx = bit_offset % 64;
        uint64_t bit_buf;
        if (bit_idx < 64 - (size_t)w || u64_idx == 3) {
            bit_buf = a[u64_idx] >> bit_idx;
        } else {
            bit_buf = (a[u64_idx] >> bit_idx) | (a[1 + u64_idx] << (64 - bit_idx));
        }
        uint64_t coef = carry + (bit_buf & window_mask);
        carry = (coef + radix / 2) >> w;
        int64_t d = (int64_t)coef - (int64_t)(carry << w);
        if (i == digits_count - 1) d += (int64_t)(carry << w);
        out[i] = d;
    }
}

#define EC g1
#define BF fq
#define BFW 4
#include "ec_tmpl.h"

#define EC g2
#define BF fq2
#define BFW 8
#include "ec_tmpl.h"

/* ---------------------------------------------------------------------------------------
 * DomainCoeff abstraction: pss.rs is generic over T: DomainCoeff<F> (field elements AND
 * group elements, dist-primitives/src/dmsm/mod.rs:34,38,85).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    size_t words;                                                  /* u64 words per element */
    void (*zero)(uint64_t *);
    void (*add)(uint64_t *, const uint64_t *, const uint64_t *);
    void (*sub)(uint64_t *, const uint64_t *, const uint64_t *);
    void (*mul_fr)(uint64_t *, const uint64_t *, const uint64_t *fr_mont);
    int (*is_zero)(const uint64_t *);
} coeff_ops;

static void c_fr_zero(uint64_t *o) { fr_zero(o); }
static void c_fr_add(uint64_t *o, const uint64_t *a, const uint64_t *b) { fr_add(o, a, b); }
static void c_fr_sub(uint64_t *o, const uint64_t *a, const uint64_t *b) { fr_sub(o, a, b); }
static void c_fr_mul(uint64_t *o, const uint64_t *a, const uint64_t *s) { fr_mul(o, a, s); }
static int c_fr_is_zero(const uint64_t *a) { return fr_is_zero(a); }
static const coeff_ops FR_OPS = {4, c_fr_zero, c_fr_add, c_fr_sub, c_fr_mul, c_fr_is_zero};

#define DEF_GROUP_OPS(G, WORDS, NEGF)                                                                 \
    static void c_##G##_zero(uint64_t *o) { G##_set_identity((G##_jac *)o); }                     \
    static void c_##G##_add(uint64_t *o, const uint64_t *a, const uint64_t *b) {                  \
        G##_jac t = *(const G##_jac *)a;                                                          \
        G##_add(&t, (const G##_jac *)b);                                                          \
        *(G##_jac *)o = t;                                                                        \
    }                                                                                             \
    static void c_##G##_sub(uint64_t *o, const uint64_t *a, const uint64_t *b) {                  \
        G##_jac nb = *(const G##_jac *)b;                                                         \
        G##_jac t = *(const G##_jac *)a;                                                          \
        if (!G##_is_identity(&nb)) NEGF(nb.Y, nb.Y);                                              \
        G##_add(&t, &nb);                                                                         \
        *(G##_jac *)o = t;                                                                        \
    }                                                                                             \
    static void c_##G##_mul(uint64_t *o, const uint64_t *a, const uint64_t *s) {                  \
        uint64_t k[4];                                                                            \
        fr_from_mont(k, s);                                                                       \
        G##_jac t;                                                                                \
        G##_mul_scalar(&t, (const G##_jac *)a, k);                                                \
        *(G##_jac *)o = t;                                                                        \
    }                                                                                             \
    static int c_##G##_is_zero(const uint64_t *a) { return G##_is_identity((const G##_jac *)a); } \
    static const coeff_ops G##_OPS = {WORDS, c_##G##_zero, c_##G##_add, c_##G##_sub, c_##G##_mul, c_##G##_is_zero};

DEF_GROUP_OPS(g1, 12, fq_neg)
DEF_GROUP_OPS(g2, 24, fq2_neg)

/* ---------------------------------------------------------------------------------------
 * ark-poly 0.4 Radix2EvaluationDomain<Fr>
 * ------------------------------------------------------------------------------------- */
typedef struct {
    size_t size;
    int log_size;
    uint64_t group_gen[4], group_gen_inv[4], size_inv[4], offset[4], offset_inv[4];
    int has_offset;
} domain_t;

/* F::get_root_of_unity(n) = TWO_ADIC_ROOT ^ (2^(28 - log n)) */
static void fr_root_of_unity(uint64_t *o, size_t n) {
    int lg = ceil_log2(n);
    fr_set(o, BN254_FR_TWO_ADIC_ROOT_MONT);
    for (int i = lg; i < BN254_FR_TWO_ADICITY; ++i) fr_sqr(o, o);
}

static void domain_new(domain_t *d, size_t num_coeffs) {
    d->log_size = ceil_log2(num_coeffs);
    d->size = (size_t)1 << d->log_size;
    fr_root_of_unity(d->group_gen, d->size);
    fr_inv(d->group_gen_inv, d->group_gen);
    uint64_t s[4];
    fr_from_u64(s, (uint64_t)d->size);
    fr_inv(d->size_inv, s);
    fr_one(d->offset);
    fr_one(d->offset_inv);
    d->has_offset = 0;
}

static void domain_get_coset(domain_t *d, const uint64_t *offset) {
    fr_set(d->offset, offset);
    fr_inv(d->offset_inv, offset);
    d->has_offset = !fr_eq(offset, BN254_FR_R);
}

static size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

/* in-order radix-2 DFT: out[k] = sum_j v[j] root^(jk); v has exactly `n` elements */
static void dft_in_place(uint64_t *v, size_t n, int lg, const uint64_t *root, const coeff_ops *ops) {
    size_t W = ops->words;
    uint64_t tmp[24], u[24];
    for (size_t i = 0; i < n; ++i) {
        size_t j = bitrev(i, lg);
        if (j > i) {
            memcpy(tmp, v + i * W, W * 8);
            memcpy(v + i * W, v + j * W, W * 8);
            memcpy(v + j * W, tmp, W * 8);
        }
    }
    /* twiddles root^k, k < n/2 */
    uint64_t *tw = (uint64_t *)malloc((n / 2 + 1) * 32);
    fr_one(tw);
    for (size_t k = 1; k < n / 2; ++k) fr_mul(tw + 4 * k, tw + 4 * (k - 1), root);
    for (int s = 0; s < lg; ++s) {
        size_t half = (size_t)1 << s, step = n >> (s + 1);
        for (size_t base = 0; base < n; base += 2 * half)
            for (size_t k = 0; k < half; ++k) {
                uint64_t *a = v + (base + k) * W, *b = v + (base + k + half) * W;
                if (k == 0) memcpy(tmp, b, W * 8);
                else ops->mul_fr(tmp, b, tw + 4 * (k * step));
                memcpy(u, a, W * 8);
                ops->add(a, u, tmp);
                ops->sub(b, u, tmp);
            }
    }
    free(tw);
}

/* fft_in_place on a buffer that already has domain size (caller did the resize) */
static void domain_fft(const domain_t *d, uint64_t *v, const coeff_ops *ops) {
    size_t W = ops->words;
    if (d->has_offset) {
        uint64_t pw[4];
        fr_one(pw);
        for (size_t i = 0; i < d->size; ++i) {
            ops->mul_fr(v + i * W, v + i * W, pw);
            fr_mul(pw, pw, d->offset);
        }
    }
    dft_in_place(v, d->size, d->log_size, d->group_gen, ops);
}

static void domain_ifft(const domain_t *d, uint64_t *v, const coeff_ops *ops) {
    size_t W = ops->words;
    dft_in_place(v, d->size, d->log_size, d->group_gen_inv, ops);
    uint64_t pw[4];
    fr_set(pw, d->size_inv);
    for (size_t i = 0; i < d->size; ++i) {
        ops->mul_fr(v + i * W, v + i * W, pw);
        if (d->has_offset) fr_mul(pw, pw, d->offset_inv);
    }
}

/* ---------------------------------------------------------------------------------------
 * secret-sharing/src/pss.rs
 * ------------------------------------------------------------------------------------- */
typedef struct {
    size_t t, l, n;
    domain_t share, secret, secret2;
} pss_t;

/* pss.rs:39-66 */
static void pss_new(pss_t *pp, size_t l) {
    pp->l = l;
    pp->n = 4 * l;
    pp->t = l;
    domain_new(&pp->share, pp->n);
    domain_new(&pp->secret, l + pp->t);
    domain_get_coset(&pp->secret, BN254_FR_GENERATOR_MONT);
    domain_new(&pp->secret2, 2 * (l + pp->t));
    domain_get_coset(&pp->secret2, BN254_FR_GENERATOR_MONT);
}

/* pss.rs:90-122 (rand != NULL) and pss.rs:69-87 (rand == NULL: det_pack, zero padding) */
static void pss_pack(const pss_t *pp, const uint64_t *secrets, const uint64_t *rand, uint64_t *out,
                     const coeff_ops *ops) {
    size_t W = ops->words;
    memcpy(out, secrets, pp->l * W * 8);
    for (size_t i = 0; i < pp->t; ++i) {
        if (rand) memcpy(out + (pp->l + i) * W, rand + i * W, W * 8);
        else ops->zero(out + (pp->l + i) * W);
    }
    domain_ifft(&pp->secret, out, ops);                 /* interpolate on the secrets coset */
    for (size_t i = pp->l + pp->t; i < pp->n; ++i) ops->zero(out + i * W); /* resize to n */
    domain_fft(&pp->share, out, ops);                   /* evaluate on the share domain */
}

/* pss.rs:125-138 */
static void pss_unpack(const pss_t *pp, const uint64_t *shares, uint64_t *out, const coeff_ops *ops) {
    size_t W = ops->words;
    uint64_t *buf = (uint64_t *)malloc(pp->n * W * 8);
    memcpy(buf, shares, pp->n * W * 8);
    domain_ifft(&pp->share, buf, ops);
    domain_fft(&pp->secret, buf, ops);                  /* fft_in_place truncates to l+t */
    memcpy(out, buf, pp->l * W * 8);
    free(buf);
}

/* evaluate on secret2 and keep entries 0,2,..,2(l-1): shared tail of unpack2 / lagrange_unpack */
static void pss_secret2_tail(const pss_t *pp, uint64_t *buf, uint64_t *out, const coeff_ops *ops) {
    size_t W = ops->words;
    domain_fft(&pp->secret2, buf, ops);
    for (size_t i = 0; i < pp->l; ++i) memcpy(out + i * W, buf + 2 * i * W, W * 8);
}

/* pss.rs:141-166 */
static void pss_unpack2(const pss_t *pp, const uint64_t *shares, uint64_t *out, const coeff_ops *ops) {
    size_t W = ops->words;
    uint64_t *buf = (uint64_t *)malloc(pp->n * W * 8);
    memcpy(buf, shares, pp->n * W * 8);
    domain_ifft(&pp->share, buf, ops);
    pss_secret2_tail(pp, buf, out, ops);
    free(buf);
}

/* secret-sharing/src/utils.rs:120-135 */
static void get_zero_roots(const uint64_t *xs, size_t k, uint64_t *result /* k+1 */) {
    for (size_t i = 0; i <= k; ++i) fr_zero(result + 4 * i);
    size_t n = k;
    fr_one(result + 4 * n);
    for (size_t i = 0; i < k; ++i) {
        n -= 1;
        fr_zero(result + 4 * n);
        for (size_t j = n; j < k; ++j) {
            uint64_t t[4];
            fr_mul(t, result + 4 * (j + 1), xs + 4 * i);
            fr_sub(result + 4 * j, result + 4 * j, t);
        }
    }
}

/* secret-sharing/src/utils.rs:78-116 + pss.rs:170-207 */
static int pss_lagrange_unpack(const pss_t *pp, const uint64_t *shares, const uint32_t *parties, size_t k,
                               uint64_t *out, const coeff_ops *ops) {
    size_t W = ops->words;
    if (!(k > 2 * (pp->t + pp->l - 1)) || k > pp->n) return -1;
    uint64_t *xs = (uint64_t *)malloc(k * 32);
    for (size_t i = 0; i < k; ++i) {
        if (parties[i] >= pp->n) { free(xs); return -1; }
        fr_pow_u64(xs + 4 * i, pp->share.group_gen, parties[i]);
    }
    uint64_t *roots = (uint64_t *)malloc((k + 1) * 32);
    get_zero_roots(xs, k, roots);
    uint64_t *num = (uint64_t *)malloc(k * (k + 1) * 32);   /* numerators[i] = roots / (x - xs[i]) */
    uint64_t *den = (uint64_t *)malloc(k * 32);
    for (size_t i = 0; i < k; ++i) {
        uint64_t *f = num + i * (k + 1) * 4;
        memcpy(f, roots, (k + 1) * 32);
        /* syn_div_in_place, a == 1 (utils.rs:47-55) */
        uint64_t c[4], t[4];
        fr_zero(c);
        for (size_t j = k + 1; j-- > 0;) {
            fr_mul(t, xs + 4 * i, c);
            fr_add(f + 4 * j, f + 4 * j, t);
            memcpy(t, f + 4 * j, 32); memcpy(f + 4 * j, c, 32); memcpy(c, t, 32);
        }
        /* eval(f, x) Horner (utils.rs:7-15) */
        uint64_t acc[4];
        fr_zero(acc);
        for (size_t j = k + 1; j-- > 0;) { fr_mul(acc, acc, xs + 4 * i); fr_add(acc, acc, f + 4 * j); }
        fr_inv(den + 4 * i, acc);
    }
    /* numerators.len() == k; result has k coefficients */
    uint64_t *buf = (uint64_t *)malloc(pp->n * W * 8);
    for (size_t j = 0; j < pp->n; ++j) ops->zero(buf + j * W);
    uint64_t y[24], tmp[24];
    for (size_t i = 0; i < k; ++i) {
        ops->mul_fr(y, shares + i * W, den + 4 * i);
        for (size_t j = 0; j < k; ++j) {
            ops->mul_fr(tmp, y, num + (i * (k + 1) + j) * 4);
            ops->add(buf + j * W, buf + j * W, tmp);
        }
    }
    /* (leading-zero truncation + fft's resize back to n cancel out) */
    pss_secret2_tail(pp, buf, out, ops);
    free(buf); free(den); free(num); free(roots); free(xs);
    return 0;
}

/* pss.rs:210-221 */
static int pss_unpack_missing_shares(const pss_t *pp, const uint64_t *shares, const uint32_t *parties, size_t k,
                                     uint64_t *out, const coeff_ops *ops) {
    if (k == pp->n) { pss_unpack2(pp, shares, out, ops); return 0; }
    return pss_lagrange_unpack(pp, shares, parties, k, out, ops);
}

/* ---------------------------------------------------------------------------------------
 * dist-primitives/src/dfft/mod.rs
 * ------------------------------------------------------------------------------------- */
/* dfft/mod.rs:322-335, literal */
static void fft_in_place_rearrange(uint64_t *data, size_t len) {
    size_t target = 0;
    uint64_t t[4];
    for (size_t pos = 0; pos < len; ++pos) {
        if (target > pos) {
            memcpy(t, data + 4 * target, 32);
            memcpy(data + 4 * target, data + 4 * pos, 32);
            memcpy(data + 4 * pos, t, 32);
        }
        size_t mask = len >> 1;
        while (target & mask) { target &= ~mask; mask >>= 1; }
        target |= mask;
    }
}

/* dfft/mod.rs:178-208, literal */
static void fft1_in_place(uint64_t *px, size_t mbyl, size_t l, const uint64_t *gen) {
    size_t dom_size = mbyl * l;
    for (int i = ceil_log2(dom_size); i >= ceil_log2(l) + 1; --i) {
        size_t poly_size = dom_size >> i;
        uint64_t factor_stride[4], factor[4];
        fr_pow_u64(factor_stride, gen, (uint64_t)1 << (i - 1));
        fr_set(factor, factor_stride);
        for (size_t k = 0; k < poly_size; ++k) {
            for (size_t j = 0; j < ((size_t)1 << (i - 1)) / l; ++j) {
                uint64_t x[4], y[4];
                fr_set(x, px + 4 * ((2 * j) * poly_size + k));
                fr_mul(y, px + 4 * ((2 * j + 1) * poly_size + k), factor);
                fr_add(px + 4 * (j * (2 * poly_size) + k), x, y);
                fr_sub(px + 4 * (j * (2 * poly_size) + k + poly_size), x, y);
            }
            fr_mul(factor, factor, factor_stride);
        }
    }
}

/* dfft/mod.rs:210-237, literal */
static void fft2_in_place(uint64_t *s1_io, size_t dom_size, size_t l, const uint64_t *gen) {
    uint64_t *s1 = s1_io;
    uint64_t *s2 = (uint64_t *)calloc(dom_size, 32);
    for (int i = ceil_log2(l); i >= 1; --i) {
        size_t poly_size = dom_size >> i;
        uint64_t factor_stride[4], factor[4];
        fr_pow_u64(factor_stride, gen, (uint64_t)1 << (i - 1));
        fr_set(factor, factor_stride);
        for (size_t k = 0; k < poly_size; ++k) {
            for (size_t j = 0; j < ((size_t)1 << (i - 1)); ++j) {
                uint64_t x[4], y[4];
                fr_set(x, s1 + 4 * (k * ((size_t)1 << i) + 2 * j));
                fr_mul(y, s1 + 4 * (k * ((size_t)1 << i) + 2 * j + 1), factor);
                fr_add(s2 + 4 * (k * ((size_t)1 << (i - 1)) + j), x, y);
                fr_sub(s2 + 4 * ((k + poly_size) * ((size_t)1 << (i - 1)) + j), x, y);
            }
            fr_mul(factor, factor, factor_stride);
        }
        uint64_t *tmp = s1; s1 = s2; s2 = tmp;            /* mem::swap */
    }
    /* rotate_right(1) (dfft/mod.rs:236), written into the caller's buffer */
    if (s1 == s1_io) { memcpy(s2, s1, dom_size * 32); uint64_t *tmp = s1; s1 = s2; s2 = tmp; }
    memcpy(s1_io, s1 + 4 * (dom_size - 1), 32);
    memcpy(s1_io + 4, s1, (dom_size - 1) * 32);
    free(s1);
}

static void distribute_powers(uint64_t *v, size_t n, const uint64_t *g) {
    uint64_t pw[4];
    fr_one(pw);
    for (size_t i = 0; i < n; ++i) { fr_mul(v + 4 * i, v + 4 * i, pw); fr_mul(pw, pw, g); }
}

/* =======================================================================================
 * Exported C interface (ctypes from tests/, smoke(), bench.py cpu_baseline only)
 * ===================================================================================== */
#define API __attribute__((visibility("default")))

API int zko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define DEF_FIELD_API(P)                                                                             \
    API void zko_##P##_mul(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {            \
        for (size_t i = 0; i < n; ++i) P##_mul(o + 4 * i, a + 4 * i, b + 4 * i);                     \
    }                                                                                                \
    API void zko_##P##_add(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {            \
        for (size_t i = 0; i < n; ++i) P##_add(o + 4 * i, a + 4 * i, b + 4 * i);                     \
    }                                                                                                \
    API void zko_##P##_sub(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {            \
        for (size_t i = 0; i < n; ++i) P##_sub(o + 4 * i, a + 4 * i, b + 4 * i);                     \
    }                                                                                                \
    API void zko_##P##_inv(const uint64_t *a, uint64_t *o, size_t n) {                               \
        for (size_t i = 0; i < n; ++i) P##_inv(o + 4 * i, a + 4 * i);                                \
    }                                                                                                \
    API void zko_##P##_to_mont(const uint64_t *a, uint64_t *o, size_t n) {                           \
        for (size_t i = 0; i < n; ++i) P##_to_mont(o + 4 * i, a + 4 * i);                            \
    }                                                                                                \
    API void zko_##P##_from_mont(const uint64_t *a, uint64_t *o, size_t n) {                         \
        for (size_t i = 0; i < n; ++i) P##_from_mont(o + 4 * i, a + 4 * i);                          \
    }                                                                                                \
    API void zko_##P##_pow_u64(const uint64_t *a, uint64_t e, uint64_t *o) { P##_pow_u64(o, a, e); }

DEF_FIELD_API(fr)
DEF_FIELD_API(fq)

API void zko_fq2_mul(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    for (size_t i = 0; i < n; ++i) fq2_mul(o + 8 * i, a + 8 * i, b + 8 * i);
}
API void zko_fq2_sqr(const uint64_t *a, uint64_t *o, size_t n) {
    for (size_t i = 0; i < n; ++i) fq2_sqr(o + 8 * i, a + 8 * i);
}
API void zko_fq2_inv(const uint64_t *a, uint64_t *o, size_t n) {
    for (size_t i = 0; i < n; ++i) fq2_inv(o + 8 * i, a + 8 * i);
}

/* ---- domains ---- */
API void zko_fr_root_of_unity(size_t n, uint64_t *o) { fr_root_of_unity(o, n); }

/* Radix2EvaluationDomain::{fft,ifft}_in_place on exactly n = 2^k elements; offset NULL = 1 */
API void zko_fr_fft(uint64_t *v, size_t n, const uint64_t *offset, int inverse) {
    domain_t d;
    domain_new(&d, n);
    if (offset) domain_get_coset(&d, offset);
    if (inverse) domain_ifft(&d, v, &FR_OPS);
    else domain_fft(&d, v, &FR_OPS);
}
API void zko_fr_distribute_powers(uint64_t *v, size_t n, const uint64_t *g) { distribute_powers(v, n, g); }
API void zko_fr_rearrange(uint64_t *v, size_t n) { fft_in_place_rearrange(v, n); }
API void zko_fft1_in_place(uint64_t *px, size_t mbyl, uint32_t l, const uint64_t *gen) { fft1_in_place(px, mbyl, l, gen); }
API void zko_fft2_in_place(uint64_t *s1, size_t m, uint32_t l, const uint64_t *gen) { fft2_in_place(s1, m, l, gen); }

/* ---- PSS over Fr, batched: column c uses secrets[c*l..], rand[c*t..] (NULL = det_pack), out[c*n..] ---- */
API void zko_pss_pack_fr(uint32_t l, const uint64_t *secrets, const uint64_t *rand, uint64_t *out, size_t cols) {
    pss_t pp;
    pss_new(&pp, l);
    for (size_t c = 0; c < cols; ++c)
        pss_pack(&pp, secrets + c * pp.l * 4, rand ? rand + c * pp.t * 4 : NULL, out + c * pp.n * 4, &FR_OPS);
}
API void zko_pss_unpack_fr(uint32_t l, const uint64_t *shares, uint64_t *out, size_t cols) {
    pss_t pp;
    pss_new(&pp, l);
    for (size_t c = 0; c < cols; ++c) pss_unpack(&pp, shares + c * pp.n * 4, out + c * pp.l * 4, &FR_OPS);
}
API void zko_pss_unpack2_fr(uint32_t l, const uint64_t *shares, uint64_t *out, size_t cols) {
    pss_t pp;
    pss_new(&pp, l);
    for (size_t c = 0; c < cols; ++c) pss_unpack2(&pp, shares + c * pp.n * 4, out + c * pp.l * 4, &FR_OPS);
}
API int zko_pss_lagrange_unpack_fr(uint32_t l, const uint64_t *shares, const uint32_t *parties, uint32_t k,
                                   uint64_t *out, size_t cols) {
    pss_t pp;
    pss_new(&pp, l);
    for (size_t c = 0; c < cols; ++c)
        if (pss_lagrange_unpack(&pp, shares + c * k * 4, parties, k, out + c * pp.l * 4, &FR_OPS)) return -1;
    return 0;
}

/* ---- PSS over group elements (Jacobian images; results normalised to Z = 1) ---- */
#define DEF_GROUP_PSS(G, WORDS)                                                                          \
    static void G##_normalize_img(uint64_t *p) {                                                         \
        G##_aff a;                                                                                       \
        G##_normalize(&a, (const G##_jac *)p);                                                           \
        G##_from_affine((G##_jac *)p, &a);                                                               \
    }                                                                                                    \
    API void zko_pss_pack_##G(uint32_t l, const uint64_t *secrets, const uint64_t *rand, uint64_t *out) {\
        pss_t pp;                                                                                        \
        pss_new(&pp, l);                                                                                 \
        pss_pack(&pp, secrets, rand, out, &G##_OPS);                                                     \
        for (size_t i = 0; i < pp.n; ++i) G##_normalize_img(out + i * WORDS);                            \
    }                                                                                                    \
    API void zko_pss_unpack2_##G(uint32_t l, const uint64_t *shares, uint64_t *out) {                    \
        pss_t pp;                                                                                        \
        pss_new(&pp, l);                                                                                 \
        pss_unpack2(&pp, shares, out, &G##_OPS);                                                         \
        for (size_t i = 0; i < pp.l; ++i) G##_normalize_img(out + i * WORDS);                            \
    }                                                                                                    \
    API void zko_pss_unpack_##G(uint32_t l, const uint64_t *shares, uint64_t *out) {                     \
        pss_t pp;                                                                                        \
        pss_new(&pp, l);                                                                                 \
        pss_unpack(&pp, shares, out, &G##_OPS);                                                          \
        for (size_t i = 0; i < pp.l; ++i) G##_normalize_img(out + i * WORDS);                            \
    }                                                                                                    \
    API void zko_##G##_add(const uint64_t *a, const uint64_t *b, uint64_t *o) {                          \
        c_##G##_add(o, a, b);                                                                            \
        G##_normalize_img(o);                                                                            \
    }                                                                                                    \
    API void zko_##G##_mul(const uint64_t *a, const uint64_t *scalar_mont, uint64_t *o) {                \
        c_##G##_mul(o, a, scalar_mont);                                                                  \
        G##_normalize_img(o);                                                                            \
    }                                                                                                    \
    API void zko_##G##_normalize(uint64_t *p) { G##_normalize_img(p); }

DEF_GROUP_PSS(g1, 12)
DEF_GROUP_PSS(g2, 24)

/* ---- MSM: VariableBaseMSM::msm (call site dist-primitives/src/dmsm/mod.rs:73) ----
 * bases: arkworks Affine images, `stride` bytes apart (72 for G1, 136 for G2);
 * scalars: Fr Montgomery images.  Returns 0, writes normalised Jacobian image. */
API int zko_g1_msm(const void *bases, size_t stride, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz,
                   int threads, int c_override) {
    g1_aff *b = (g1_aff *)malloc(sizeof(g1_aff) * (n ? n : 1));
    uint64_t *s = (uint64_t *)malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) {
        const uint8_t *p = (const uint8_t *)bases + i * stride;
        memcpy(b[i].x, p, 32);
        memcpy(b[i].y, p + 32, 32);
        b[i].inf = p[64] != 0;
        fr_from_mont(s + 4 * i, scalars_mont + 4 * i);    /* into_bigint() */
    }
    g1_jac r;
    g1_msm_bigint(&r, b, s, n, threads, c_override);
    memcpy(out_xyz, &r, sizeof r);
    g1_normalize_img(out_xyz);
    free(b); free(s);
    return 0;
}

API int zko_g2_msm(const void *bases, size_t stride, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz,
                   int threads, int c_override) {
    g2_aff *b = (g2_aff *)malloc(sizeof(g2_aff) * (n ? n : 1));
    uint64_t *s = (uint64_t *)malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) {
        const uint8_t *p = (const uint8_t *)bases + i * stride;
        memcpy(b[i].x, p, 64);
        memcpy(b[i].y, p + 64, 64);
        b[i].inf = p[128] != 0;
        fr_from_mont(s + 4 * i, scalars_mont + 4 * i);
    }
    g2_jac r;
    g2_msm_bigint(&r, b, s, n, threads, c_override);
    memcpy(out_xyz, &r, sizeof r);
    g2_normalize_img(out_xyz);
    free(b); free(s);
    return 0;
}

/* bases[i] = scalars[i] * generator, written as arkworks Affine images (test-data helper) */
API void zko_g1_fixed_base(const uint64_t *scalars_mont, size_t n, void *out, size_t stride) {
    g1_jac g;
    fq_set(g.X, BN254_G1_GEN_X_MONT); fq_set(g.Y, BN254_G1_GEN_Y_MONT); fq_one(g.Z);
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (size_t i = 0; i < n; ++i) {
        uint64_t k[4];
        fr_from_mont(k, scalars_mont + 4 * i);
        g1_jac r;
        g1_mul_scalar(&r, &g, k);
        g1_aff a;
        g1_normalize(&a, &r);
        uint8_t *p = (uint8_t *)out + i * stride;
        memset(p, 0, stride);
        memcpy(p, a.x, 32); memcpy(p + 32, a.y, 32); p[64] = (uint8_t)a.inf;
    }
}

API void zko_g2_fixed_base(const uint64_t *scalars_mont, size_t n, void *out, size_t stride) {
    g2_jac g;
    memcpy(g.X, BN254_G2_GEN_X_C0_MONT, 32); memcpy(g.X + 4, BN254_G2_GEN_X_C1_MONT, 32);
    memcpy(g.Y, BN254_G2_GEN_Y_C0_MONT, 32); memcpy(g.Y + 4, BN254_G2_GEN_Y_C1_MONT, 32);
    fq2_one(g.Z);
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (size_t i = 0; i < n; ++i) {
        uint64_t k[4];
        fr_from_mont(k, scalars_mont + 4 * i);
        g2_jac r;
        g2_mul_scalar(&r, &g, k);
        g2_aff a;
        g2_normalize(&a, &r);
        uint8_t *p = (uint8_t *)out + i * stride;
        memset(p, 0, stride);
        memcpy(p, a.x, 64); memcpy(p + 64, a.y, 64); p[128] = (uint8_t)a.inf;
    }
}

/* on-curve check of an Affine image (pins the fixtures' points) */
API int zko_g1_on_curve(const void *aff) {
    const uint8_t *p = (const uint8_t *)aff;
    if (p[64]) return 1;
    uint64_t x[4], y[4], l[4], r[4];
    memcpy(x, p, 32); memcpy(y, p + 32, 32);
    fq_sqr(l, y);
    fq_sqr(r, x); fq_mul(r, r, x); fq_add(r, r, BN254_G1_B_MONT);
    return fq_eq(l, r);
}
API int zko_g2_on_curve(const void *aff) {
    const uint8_t *p = (const uint8_t *)aff;
    if (p[128]) return 1;
    uint64_t x[8], y[8], l[8], r[8], b[8];
    memcpy(x, p, 64); memcpy(y, p + 64, 64);
    memcpy(b, BN254_G2_B_C0_MONT, 32); memcpy(b + 4, BN254_G2_B_C1_MONT, 32);
    fq2_sqr(l, y);
    fq2_sqr(r, x); fq2_mul(r, r, x); fq2_add(r, r, b);
    return fq2_eq(l, r);
}

/* ---- King closure of fft2_with_rearrange: dist-primitives/src/dfft/mod.rs:264-304 ----
 * shares_by_party[r] = vector (mbyl x Fr) received from parties[r]; rand = mbyl x t random points
 * (column-major: column i uses rand[i*t .. i*t+t)); out_by_party[p] = mbyl x Fr for each of n parties. */
API int zko_king_fft2(const uint64_t *const *shares_by_party, const uint32_t *parties, uint32_t n_recv, size_t mbyl,
                      uint32_t l, const uint64_t *gen, const uint64_t *g, int rearrange, const uint64_t *rand,
                      uint64_t *const *out_by_party) {
    pss_t pp;
    pss_new(&pp, l);
    size_t m = mbyl * pp.l;
    uint64_t *s1 = (uint64_t *)calloc(m, 32);
    uint64_t col[32 * 4], tmp[32 * 4];
    if (pp.n > 32) return -1;
    for (size_t i = 0; i < mbyl; ++i) {                         /* transpose + unpack (:265-274) */
        for (uint32_t r = 0; r < n_recv; ++r) memcpy(col + 4 * r, shares_by_party[r] + 4 * i, 32);
        if (pss_unpack_missing_shares(&pp, col, parties, n_recv, tmp, &FR_OPS)) { free(s1); return -1; }
        memcpy(s1 + 4 * i * pp.l, tmp, pp.l * 32);
    }
    fft2_in_place(s1, m, pp.l, gen);                            /* :276 */
    if (!fr_eq(g, BN254_FR_R)) distribute_powers(s1, m, g);     /* :278-280 */
    uint64_t sec[32 * 4];
    if (rearrange) {                                            /* :284-300 */
        fft_in_place_rearrange(s1, m);
        for (size_t i = 0; i < mbyl; ++i) {
            for (size_t j = 0; j < pp.l; ++j) memcpy(sec + 4 * j, s1 + 4 * (i + j * mbyl), 32);
            pss_pack(&pp, sec, rand + i * pp.t * 4, col, &FR_OPS);
            for (size_t p = 0; p < pp.n; ++p) memcpy(out_by_party[p] + 4 * i, col + 4 * p, 32);
        }
    } else {                                                    /* pack_vec, :302 */
        for (size_t i = 0; i < mbyl; ++i) {
            pss_pack(&pp, s1 + 4 * i * pp.l, rand + i * pp.t * 4, col, &FR_OPS);
            for (size_t p = 0; p < pp.n; ++p) memcpy(out_by_party[p] + 4 * i, col + 4 * p, 32);
        }
    }
    free(s1);
    return 0;
}

/* King closure of deg_red: dist-primitives/src/utils/deg_red.rs:103-111 */
API int zko_deg_red_king(const uint64_t *const *shares_by_party, const uint32_t *parties, uint32_t n_recv, size_t cols,
                         uint32_t l, const uint64_t *rand, uint64_t *const *out_by_party) {
    pss_t pp;
    pss_new(&pp, l);
    if (pp.n > 32) return -1;
    uint64_t col[32 * 4], xi[32 * 4];
    for (size_t i = 0; i < cols; ++i) {
        for (uint32_t r = 0; r < n_recv; ++r) memcpy(col + 4 * r, shares_by_party[r] + 4 * i, 32);
        if (pss_unpack_missing_shares(&pp, col, parties, n_recv, xi, &FR_OPS)) return -1;
        pss_pack(&pp, xi, rand + i * pp.t * 4, col, &FR_OPS);
        for (size_t p = 0; p < pp.n; ++p) memcpy(out_by_party[p] + 4 * i, col + 4 * p, 32);
    }
    return 0;
}

/* bases[i] = (start + i*step) * G1 generator, i < n, as arkworks Affine images (benchmark data for the
 * CPU baseline leg: random-looking distinct points at ~1 us each instead of a 254-bit scalar
 * multiplication each).  Chunks run in parallel; each chunk is normalised with one batched inversion. */
API void zko_g1_sequence(const uint64_t *start_mont, const uint64_t *step_mont, size_t n, void *out, size_t stride) {
    const size_t CH = 4096;
    g1_jac g;
    fq_set(g.X, BN254_G1_GEN_X_MONT); fq_set(g.Y, BN254_G1_GEN_Y_MONT); fq_one(g.Z);
    uint64_t kstep[4];
    fr_from_mont(kstep, step_mont);
    g1_jac dj;
    g1_mul_scalar(&dj, &g, kstep);
    g1_aff d;
    g1_normalize(&d, &dj);
    size_t nch = (n + CH - 1) / CH;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (size_t c = 0; c < nch; ++c) {
        size_t lo = c * CH, hi = lo + CH < n ? lo + CH : n, cnt = hi - lo;
        uint64_t s[4], off[4], k[4];
        fr_from_u64(off, (uint64_t)lo);
        fr_mul(off, off, step_mont);
        fr_add(s, start_mont, off);
        fr_from_mont(k, s);
        g1_jac *pts = (g1_jac *)malloc(sizeof(g1_jac) * cnt);
        uint64_t *pref = (uint64_t *)malloc(32 * cnt);
        g1_mul_scalar(&pts[0], &g, k);
        for (size_t i = 1; i < cnt; ++i) { pts[i] = pts[i - 1]; g1_add_mixed(&pts[i], &d); }
        /* Montgomery batch inversion of the Z coordinates (identity points have Z = 0: skipped) */
        uint64_t acc[4], inv[4];
        fq_one(acc);
        for (size_t i = 0; i < cnt; ++i) {
            fq_set(pref + 4 * i, acc);
            if (!fq_is_zero(pts[i].Z)) fq_mul(acc, acc, pts[i].Z);
        }
        fq_inv(inv, acc);
        for (size_t i = cnt; i-- > 0;) {
            uint8_t *p = (uint8_t *)out + (lo + i) * stride;
            memset(p, 0, stride);
            if (fq_is_zero(pts[i].Z)) { p[64] = 1; continue; }
            uint64_t zi[4], zi2[4], zi3[4], x[4], y[4];
            fq_mul(zi, inv, pref + 4 * i);
            fq_mul(inv, inv, pts[i].Z);
            fq_sqr(zi2, zi); fq_mul(zi3, zi2, zi);
            fq_mul(x, pts[i].X, zi2); fq_mul(y, pts[i].Y, zi3);
            memcpy(p, x, 32); memcpy(p + 32, y, 32);
        }
        free(pts); free(pref);
    }
}

/* sum of n Fr elements (closed-form MSM checks at sizes the big-int model cannot reach) */
API void zko_fr_sum(const uint64_t *a, size_t n, uint64_t *o) {
    uint64_t acc[4];
    fr_zero(acc);
    for (size_t i = 0; i < n; ++i) fr_add(acc, acc, a + 4 * i);
    fr_set(o, acc);
}


/* One output of Radix2EvaluationDomain::fft by its definition: out = sum_j x_j * point^j (Horner), point = offset * w^k.
 * O(n) per output: lets a test check sampled outputs of a transform far larger than the in-order oracle FFT can hold
 * in seconds (d_fft at m = 2^24).  OpenMP over `blocks` contiguous segments, each folded by Horner and weighted. */
API void zko_fr_eval_poly(const uint64_t *x, size_t n, const uint64_t *point, uint64_t *o, int threads) {
    if (threads < 1) threads = 1;
    size_t blocks = (size_t)threads * 4;
    if (blocks > n) blocks = n ? n : 1;
    uint64_t *part = (uint64_t *)malloc(blocks * 32);
    size_t seg = (n + blocks - 1) / blocks;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (long b = 0; b < (long)blocks; ++b) {
        size_t lo = (size_t)b * seg, hi = lo + seg < n ? lo + seg : n;
        uint64_t acc[4];
        fr_zero(acc);
        for (size_t j = hi; j-- > lo;) { fr_mul(acc, acc, point); fr_add(acc, acc, x + 4 * j); }
        fr_set(part + 4 * b, acc);
    }
    /* total = sum_b part_b * point^(b*seg) */
    uint64_t pseg[4], w[4], acc[4], t[4];
    fr_set(pseg, point);
    { /* point^seg by square-and-multiply */
        uint64_t base[4], r[4];
        fr_set(base, point); fr_one(r);
        for (size_t e = seg; e; e >>= 1) { if (e & 1) fr_mul(r, r, base); fr_sqr(base, base); }
        fr_set(pseg, r);
    }
    fr_one(w); fr_zero(acc);
    for (size_t b = 0; b < blocks; ++b) {
        fr_mul(t, part + 4 * b, w);
        fr_add(acc, acc, t);
        fr_mul(w, w, pseg);
    }
    fr_set(o, acc);
    free(part);
}

/* King closure of d_pp: dist-primitives/src/dpp/mod.rs:41-76.  shares_by_party[r] = the 2*cols-element
 * vector (num shares then den shares) received from parties[r]; rand = cols x t packing randomness;
 * out_by_party[p] = cols elements.  Returns -2 if a denominator is zero (the reference unwraps). */
API int zko_dpp_king(const uint64_t *const *shares_by_party, const uint32_t *parties, uint32_t n_recv, size_t cols,
                     uint32_t l, const uint64_t *rand, uint64_t *const *out_by_party) {
    pss_t pp;
    pss_new(&pp, l);
    if (pp.n > 32) return -1;
    size_t m = cols * pp.l;
    uint64_t *numden = (uint64_t *)malloc(2 * m * 32);
    uint64_t col[32 * 4], tmp[32 * 4];
    for (size_t i = 0; i < 2 * cols; ++i) {                       /* transpose + unpack, :44-53 */
        for (uint32_t r = 0; r < n_recv; ++r) memcpy(col + 4 * r, shares_by_party[r] + 4 * i, 32);
        if (pss_unpack_missing_shares(&pp, col, parties, n_recv, tmp, &FR_OPS)) { free(numden); return -1; }
        memcpy(numden + 4 * i * pp.l, tmp, pp.l * 32);
    }
    for (size_t i = 0; i < m; ++i) {                              /* :55-58 */
        uint64_t den[4];
        if (fr_is_zero(numden + 4 * (i + m))) { free(numden); return -2; }
        fr_inv(den, numden + 4 * (i + m));
        fr_mul(numden + 4 * i, numden + 4 * i, den);
    }
    for (size_t i = 1; i < m; ++i) fr_mul(numden + 4 * i, numden + 4 * i, numden + 4 * (i - 1));   /* :62-66 */
    for (size_t i = 0; i < cols; ++i) {                           /* pack_vec + transpose, :71-76 */
        pss_pack(&pp, numden + 4 * i * pp.l, rand + i * pp.t * 4, col, &FR_OPS);
        for (size_t p = 0; p < pp.n; ++p) memcpy(out_by_party[p] + 4 * i, col + 4 * p, 32);
    }
    free(numden);
    return 0;
}
