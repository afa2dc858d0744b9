/* oracle/ec_tmpl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Short-Weierstrass (a = 0) Jacobian arithmetic + arkworks-style Pippenger, instantiated for
 * BN254 G1 (over Fq) and G2 (over Fq2) by zkoracle.c.  Restates ark-ec 0.4.2
 * `short_weierstrass::{Affine,Projective}` and `VariableBaseMSM::msm` -> `msm_bigint_wnaf`
 * (un-vendored dependency; the reference's call site is dist-primitives/src/dmsm/mod.rs:73):
 *   Projective = Jacobian (X/Z^2, Y/Z^3), identity Z = 0 constructed as (1,1,0);
 *   `+= &Affine`  = madd-2007-bl with the P==Q -> double and P==-Q -> identity branches;
 *   `+= &Projective` = add-2007-bl;  `double_in_place` = dbl-2009-l.
 *
 * Parameters: EC (prefix), BF (base-field prefix), BFW (u64 words per base-field element)
 */
#define EC_CAT_(a, b) a##_##b
#define EC_CAT(a, b) EC_CAT_(a, b)
#define EN(name) EC_CAT(EC, name)
#define F(name) EC_CAT(BF, name)
#define W BFW

typedef struct { uint64_t X[W], Y[W], Z[W]; } EN(jac);
typedef struct { uint64_t x[W], y[W]; int inf; } EN(aff);

static inline void EN(set_identity)(EN(jac) * p) { F(one)(p->X); F(one)(p->Y); F(zero)(p->Z); }
static inline int EN(is_identity)(const EN(jac) * p) { return F(is_zero)(p->Z); }

static inline void EN(from_affine)(EN(jac) * p, const EN(aff) * a) {
    if (a->inf) { EN(set_identity)(p); return; }
    F(set)(p->X, a->x); F(set)(p->Y, a->y); F(one)(p->Z);
}

static inline void EN(neg_aff)(EN(aff) * o, const EN(aff) * a) {
    *o = *a;
    if (!a->inf) F(neg)(o->y, a->y);
}

/* dbl-2009-l, a = 0 */
static void EN(double_in_place)(EN(jac) * p) {
    if (EN(is_identity)(p)) return;
    uint64_t A[W], B[W], C[W], D[W], E[W], FF[W], t[W];
    F(sqr)(A, p->X);
    F(sqr)(B, p->Y);
    F(sqr)(C, B);
    F(add)(t, p->X, B); F(sqr)(t, t); F(sub)(t, t, A); F(sub)(t, t, C); F(dbl)(D, t);
    F(dbl)(E, A); F(add)(E, E, A);
    F(sqr)(FF, E);
    F(mul)(p->Z, p->Y, p->Z); F(dbl)(p->Z, p->Z);      /* Z3 = 2*Y1*Z1 (uses old Y) */
    F(sub)(p->X, FF, D); F(sub)(p->X, p->X, D);          /* X3 = F - 2D */
    F(sub)(t, D, p->X); F(mul)(t, E, t);
    F(dbl)(C, C); F(dbl)(C, C); F(dbl)(C, C);            /* 8C */
    F(sub)(p->Y, t, C);
}

/* madd-2007-bl: p += q (q affine) */
static void EN(add_mixed)(EN(jac) * p, const EN(aff) * q) {
    if (q->inf) return;
    if (EN(is_identity)(p)) { F(set)(p->X, q->x); F(set)(p->Y, q->y); F(one)(p->Z); return; }
    uint64_t Z1Z1[W], U2[W], S2[W], H[W], HH[W], I[W], J[W], r[W], V[W], t[W];
    F(sqr)(Z1Z1, p->Z);
    F(mul)(U2, q->x, Z1Z1);
    F(mul)(S2, q->y, p->Z); F(mul)(S2, S2, Z1Z1);
    if (F(eq)(p->X, U2)) {
        if (F(eq)(p->Y, S2)) EN(double_in_place)(p);
        else EN(set_identity)(p);
        return;
    }
    F(sub)(H, U2, p->X);
    F(sqr)(HH, H);
    F(dbl)(I, HH); F(dbl)(I, I);
    F(mul)(J, H, I);
    F(sub)(r, S2, p->Y); F(dbl)(r, r);
    F(mul)(V, p->X, I);
    /* Z3 = (Z1+H)^2 - Z1Z1 - HH */
    F(add)(t, p->Z, H); F(sqr)(t, t); F(sub)(t, t, Z1Z1); F(sub)(p->Z, t, HH);
    /* X3 = r^2 - J - 2V */
    F(sqr)(t, r); F(sub)(t, t, J); F(sub)(t, t, V); F(sub)(p->X, t, V);
    /* Y3 = r*(V - X3) - 2*Y1*J */
    F(mul)(J, p->Y, J); F(dbl)(J, J);
    F(sub)(t, V, p->X); F(mul)(t, r, t); F(sub)(p->Y, t, J);
}

/* add-2007-bl: p += q */
static void EN(add)(EN(jac) * p, const EN(jac) * q) {
    if (EN(is_identity)(q)) return;
    if (EN(is_identity)(p)) { *p = *q; return; }
    uint64_t Z1Z1[W], Z2Z2[W], U1[W], U2[W], S1[W], S2[W], H[W], I[W], J[W], r[W], V[W], t[W];
    F(sqr)(Z1Z1, p->Z);
    F(sqr)(Z2Z2, q->Z);
    F(mul)(U1, p->X, Z2Z2);
    F(mul)(U2, q->X, Z1Z1);
    F(mul)(S1, p->Y, q->Z); F(mul)(S1, S1, Z2Z2);
    F(mul)(S2, q->Y, p->Z); F(mul)(S2, S2, Z1Z1);
    if (F(eq)(U1, U2)) {
        if (F(eq)(S1, S2)) EN(double_in_place)(p);
        else EN(set_identity)(p);
        return;
    }
    F(sub)(H, U2, U1);
    F(dbl)(I, H); F(sqr)(I, I);
    F(mul)(J, H, I);
    F(sub)(r, S2, S1); F(dbl)(r, r);
    F(mul)(V, U1, I);
    /* Z3 = ((Z1+Z2)^2 - Z1Z1 - Z2Z2) * H */
    F(add)(t, p->Z, q->Z); F(sqr)(t, t); F(sub)(t, t, Z1Z1); F(sub)(t, t, Z2Z2); F(mul)(p->Z, t, H);
    F(sqr)(t, r); F(sub)(t, t, J); F(sub)(t, t, V); F(sub)(p->X, t, V);
    F(mul)(J, S1, J); F(dbl)(J, J);
    F(sub)(t, V, p->X); F(mul)(t, r, t); F(sub)(p->Y, t, J);
}

/* into_affine: unique normal form (x, y, inf) -- what "bit-exact" means for a group element */
static void EN(normalize)(EN(aff) * o, const EN(jac) * p) {
    if (EN(is_identity)(p)) { F(zero)(o->x); F(zero)(o->y); o->inf = 1; return; }
    uint64_t zi[W], zi2[W], zi3[W];
    F(inv)(zi, p->Z);
    F(sqr)(zi2, zi);
    F(mul)(zi3, zi2, zi);
    F(mul)(o->x, p->X, zi2);
    F(mul)(o->y, p->Y, zi3);
    o->inf = 0;
}

/* p = k * a, k canonical 4-limb scalar (MSB-first double-and-add) */
static void EN(mul_scalar)(EN(jac) * out, const EN(jac) * a, const uint64_t k[4]) {
    EN(jac) acc;
    EN(set_identity)(&acc);
    for (int i = 255; i >= 0; --i) {
        EN(double_in_place)(&acc);
        if ((k[i >> 6] >> (i & 63)) & 1) EN(add)(&acc, a);
    }
    *out = acc;
}

/* ark-ec 0.4.2 msm_bigint_wnaf.  `scalars` are canonical (already into_bigint'ed) 4-limb values.
 * threads > 1 parallelises over windows (what rayon does under feature "parallel").
 * c_override > 0 replaces arkworks' window rule (used by tests to cross-check window sizes). */
static void EN(msm_bigint)(EN(jac) * out, const EN(aff) * bases, const uint64_t *scalars, size_t size,
                           int threads, int c_override) {
    EN(set_identity)(out);
    if (size == 0) return;
    int c = c_override > 0 ? c_override : ark_window_size(size);
    const int num_bits = 254;
    int digits_count = (num_bits + c - 1) / c;
    int64_t *digits = (int64_t *)malloc(sizeof(int64_t) * size * (size_t)digits_count);
    for (size_t i = 0; i < size; ++i) make_digits(scalars + 4 * i, c, num_bits, digits + i * digits_count);
    EN(jac) *window_sums = (EN(jac) *)malloc(sizeof(EN(jac)) * digits_count);
    size_t nb = (size_t)1 << c;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
    for (int w = 0; w < digits_count; ++w) {
        EN(jac) *buckets = (EN(jac) *)malloc(sizeof(EN(jac)) * nb);
        for (size_t b = 0; b < nb; ++b) EN(set_identity)(&buckets[b]);
        for (size_t i = 0; i < size; ++i) {
            int64_t d = digits[i * digits_count + w];
            if (d > 0) {
                EN(add_mixed)(&buckets[d - 1], &bases[i]);
            } else if (d < 0) {
                EN(aff) nq;
                EN(neg_aff)(&nq, &bases[i]);
                EN(add_mixed)(&buckets[-d - 1], &nq);
            }
        }
        EN(jac) running, res;
        EN(set_identity)(&running);
        EN(set_identity)(&res);
        for (size_t b = nb; b-- > 0;) {
            EN(add)(&running, &buckets[b]);
            EN(add)(&res, &running);
        }
        window_sums[w] = res;
        free(buckets);
    }
    EN(jac) total;
    EN(set_identity)(&total);
    for (int w = digits_count - 1; w >= 1; --w) {
        EN(add)(&total, &window_sums[w]);
        for (int k = 0; k < c; ++k) EN(double_in_place)(&total);
    }
    EN(add)(&total, &window_sums[0]);
    *out = total;
    free(window_sums);
    free(digits);
}

#undef EN
#undef F
#undef W
#undef EC
#undef BF
#undef BFW
