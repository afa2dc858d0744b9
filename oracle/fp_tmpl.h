/* oracle/fp_tmpl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * 4 x u64 Montgomery prime-field template, instantiated twice (Fr, Fq) by zkoracle.c.
 * Restates ark-ff 0.4.2 `Fp<MontBackend<C,4>,4>` (un-vendored dependency of the reference,
 * dist-primitives/Cargo.toml:9-14): value a*2^256 mod p, fully reduced, 4 little-endian u64
 * limbs.  BN254 moduli are 254-bit, so the CIOS accumulator never needs a sixth limb.
 *
 * Parameters (macros defined by the includer):  FP (name prefix), FP_MOD, FP_R, FP_R2, FP_INV
 */
#define FP_CAT_(a, b) a##_##b
#define FP_CAT(a, b) FP_CAT_(a, b)
#define FN(name) FP_CAT(FP, name)

static inline void FN(set)(uint64_t *o, const uint64_t *a) { memcpy(o, a, 32); }
static inline void FN(zero)(uint64_t *o) { memset(o, 0, 32); }
static inline void FN(one)(uint64_t *o) { memcpy(o, FP_R, 32); }
static inline int FN(is_zero)(const uint64_t *a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int FN(eq)(const uint64_t *a, const uint64_t *b) { return memcmp(a, b, 32) == 0; }

static inline int FN(geq_mod)(const uint64_t *a) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > FP_MOD[i]) return 1;
        if (a[i] < FP_MOD[i]) return 0;
    }
    return 1;
}

static inline void FN(sub_mod_raw)(uint64_t *a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - FP_MOD[i] - (uint64_t)br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}

static inline void FN(add)(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; ++i) {
        c += (u128)a[i] + b[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (FN(geq_mod)(t)) FN(sub_mod_raw)(t);
    memcpy(o, t, 32);
}

static inline void FN(sub)(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t[4];
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)t[i] + FP_MOD[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(o, t, 32);
}

static inline void FN(neg)(uint64_t *o, const uint64_t *a) {
    uint64_t z[4] = {0, 0, 0, 0};
    FN(sub)(o, z, a);
}

static inline void FN(dbl)(uint64_t *o, const uint64_t *a) { FN(add)(o, a, a); }

/* CIOS Montgomery product: o = a*b*2^-256 mod p */
static inline void FN(mul)(uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; ++i) {
        u128 c;
        uint64_t bi = b[i];
        c = (u128)a[0] * bi + t0; t0 = (uint64_t)c; c >>= 64;
        c += (u128)a[1] * bi + t1; t1 = (uint64_t)c; c >>= 64;
        c += (u128)a[2] * bi + t2; t2 = (uint64_t)c; c >>= 64;
        c += (u128)a[3] * bi + t3; t3 = (uint64_t)c; c >>= 64;
        t4 += (uint64_t)c;
        uint64_t m = t0 * FP_INV;
        c = (u128)m * FP_MOD[0] + t0; c >>= 64;
        c += (u128)m * FP_MOD[1] + t1; t0 = (uint64_t)c; c >>= 64;
        c += (u128)m * FP_MOD[2] + t2; t1 = (uint64_t)c; c >>= 64;
        c += (u128)m * FP_MOD[3] + t3; t2 = (uint64_t)c; c >>= 64;
        c += t4; t3 = (uint64_t)c; t4 = (uint64_t)(c >> 64);
    }
    uint64_t t[4] = {t0, t1, t2, t3};
    if (t4 || FN(geq_mod)(t)) FN(sub_mod_raw)(t);
    memcpy(o, t, 32);
}

static inline void FN(sqr)(uint64_t *o, const uint64_t *a) { FN(mul)(o, a, a); }

/* Montgomery <-> canonical (ark-ff into_bigint / from_bigint) */
static inline void FN(from_mont)(uint64_t *o, const uint64_t *a) {
    uint64_t one[4] = {1, 0, 0, 0};
    FN(mul)(o, a, one);
}
static inline void FN(to_mont)(uint64_t *o, const uint64_t *a) { FN(mul)(o, a, FP_R2); }

static inline void FN(from_u64)(uint64_t *o, uint64_t v) {
    uint64_t t[4] = {v, 0, 0, 0};
    FN(to_mont)(o, t);
}

/* o = a^e, e a canonical little-endian multi-limb exponent */
static void FN(pow)(uint64_t *o, const uint64_t *a, const uint64_t *e, int elimbs) {
    uint64_t acc[4], base[4];
    FN(one)(acc);
    FN(set)(base, a);
    for (int i = 0; i < elimbs; ++i)
        for (int b = 0; b < 64; ++b) {
            if ((e[i] >> b) & 1) FN(mul)(acc, acc, base);
            FN(sqr)(base, base);
        }
    FN(set)(o, acc);
}

static void FN(pow_u64)(uint64_t *o, const uint64_t *a, uint64_t e) { FN(pow)(o, a, &e, 1); }

/* Fermat inverse a^(p-2); 0 -> 0 */
static void FN(inv)(uint64_t *o, const uint64_t *a) {
    uint64_t e[4];
    memcpy(e, FP_MOD, 32);
    e[0] -= 2; /* both moduli end in ...01 / ...47, no borrow */
    FN(pow)(o, a, e, 4);
}

#undef FN
#undef FP
#undef FP_MOD
#undef FP_R
#undef FP_R2
#undef FP_INV
