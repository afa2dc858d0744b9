// host_fr.hpp -- minimal BN254 Fr arithmetic on the HOST, used only to derive launch parameters
// (roots of unity, PackedSharingParams matrices, Lagrange weights).  All bulk arithmetic runs in
// the CUDA kernels; nothing here touches per-element data.
// Restates the parameter set-up of secret-sharing/src/pss.rs:39-66 (domains) and the linear maps
// behind pack/unpack/unpack2/lagrange_unpack (pss.rs:69-207).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "bn254_consts.cuh"

namespace zkg {
namespace host {

struct HFr {
    uint64_t v[4];
    bool operator==(const HFr& o) const { return memcmp(v, o.v, 32) == 0; }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
};

typedef unsigned __int128 u128;

inline HFr h_zero() { HFr r; memset(r.v, 0, 32); return r; }
inline HFr h_one() { HFr r; memcpy(r.v, BN254_FR_R_64, 32); return r; }
inline HFr h_load(const uint64_t* p) { HFr r; memcpy(r.v, p, 32); return r; }

inline bool h_geq_mod(const uint64_t* a) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > BN254_FR_MOD_64[i]) return true;
        if (a[i] < BN254_FR_MOD_64[i]) return false;
    }
    return true;
}
inline void h_sub_mod(uint64_t* a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - BN254_FR_MOD_64[i] - (uint64_t)br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
inline HFr h_add(const HFr& a, const HFr& b) {
    HFr r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (h_geq_mod(r.v)) h_sub_mod(r.v);
    return r;
}
inline HFr h_sub(const HFr& a, const HFr& b) {
    HFr r;
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.v[i] - b.v[i] - (uint64_t)br;
        r.v[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)r.v[i] + BN254_FR_MOD_64[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
}
// Montgomery product, operand-scanning with a 9-limb accumulator then word-by-word reduction
inline HFr h_mul(const HFr& a, const HFr& b) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.v[j] * b.v[i] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + 4] = (uint64_t)c;
    }
    for (int i = 0; i < 4; ++i) {
        uint64_t m = t[i] * BN254_FR_INV_64;
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)m * BN254_FR_MOD_64[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        for (int k = i + 4; c != 0 && k < 9; ++k) { c += t[k]; t[k] = (uint64_t)c; c >>= 64; }
    }
    HFr r;
    memcpy(r.v, t + 4, 32);
    if (t[8] || h_geq_mod(r.v)) h_sub_mod(r.v);
    return r;
}
inline HFr h_pow(HFr a, uint64_t e) {
    HFr acc = h_one();
    while (e) {
        if (e & 1) acc = h_mul(acc, a);
        a = h_mul(a, a);
        e >>= 1;
    }
    return acc;
}
inline HFr h_inv(const HFr& a) {      // a^(r-2)
    uint64_t e[4];
    memcpy(e, BN254_FR_MOD_64, 32);
    e[0] -= 2;
    HFr acc = h_one(), base = a;
    for (int i = 0; i < 4; ++i)
        for (int b = 0; b < 64; ++b) {
            if ((e[i] >> b) & 1) acc = h_mul(acc, base);
            base = h_mul(base, base);
        }
    return acc;
}
inline HFr h_from_u64(uint64_t x) {
    HFr t = h_zero();
    t.v[0] = x;
    return h_mul(t, h_load(BN254_FR_R2_64));
}
// F::get_root_of_unity(n), n = 2^k <= 2^28
inline HFr h_root_of_unity(uint64_t n) {
    int lg = 0;
    while (((uint64_t)1 << lg) < n) ++lg;
    HFr w = h_load(BN254_FR_TWO_ADIC_ROOT_MONT_64);
    for (int i = lg; i < BN254_FR_TWO_ADICITY; ++i) w = h_mul(w, w);
    return w;
}

// The linear maps of PackedSharingParams (n = 4l parties, t = l), as dense row-major matrices.
struct PssMatrices {
    uint32_t l = 0, t = 0, n = 0;
    std::vector<HFr> pack;      // n x (l+t):  shares  = pack    * (secrets || rand)     pss.rs:90-122
    std::vector<HFr> unpack;    // l x n    :  secrets = unpack  * shares  (degree < l+t) pss.rs:125-138
    std::vector<HFr> unpack2;   // l x n    :  secrets = unpack2 * shares  (degree < n)   pss.rs:141-166
};

// All three maps are "interpolate on one domain, evaluate on another":
//   share domain   x_j = zeta_n^j           (j < n)
//   secret domain  s_i = g * zeta_{l+t}^i   (i < l+t),   secret2: u_i = g * zeta_{2(l+t)}^i,  g = 5
inline void pss_matrices(uint32_t l, PssMatrices* out) {
    const uint32_t t = l, n = 4 * l, k = l + t;
    out->l = l; out->t = t; out->n = n;
    HFr g = h_load(BN254_FR_GENERATOR_MONT_64);
    HFr zn = h_root_of_unity(n), zk = h_root_of_unity(k), z2k = h_root_of_unity(2 * k);
    HFr n_inv = h_inv(h_from_u64(n)), k_inv = h_inv(h_from_u64(k));
    HFr g_inv = h_inv(g), zn_inv = h_inv(zn), zk_inv = h_inv(zk);
    // pack: coeffs c_d = g^-d * (1/k) sum_i v_i zk^(-i d)  (d < k);  share_j = sum_d c_d zn^(j d)
    out->pack.assign((size_t)n * k, h_zero());
    for (uint32_t j = 0; j < n; ++j)
        for (uint32_t i = 0; i < k; ++i) {
            HFr acc = h_zero();
            for (uint32_t d = 0; d < k; ++d) {
                HFr term = h_mul(h_pow(g_inv, d), h_pow(zk_inv, (uint64_t)i * d));
                term = h_mul(term, h_pow(zn, (uint64_t)j * d));
                acc = h_add(acc, term);
            }
            out->pack[(size_t)j * k + i] = h_mul(acc, k_inv);
        }
    // share-domain interpolation: c_d = (1/n) sum_j share_j zn^(-j d)   (d < n)
    // unpack : truncate to d < k, evaluate at s_i = g zk^i
    // unpack2: all d < n,       evaluate at u_{2i} = g z2k^(2i)
    out->unpack.assign((size_t)l * n, h_zero());
    out->unpack2.assign((size_t)l * n, h_zero());
    for (uint32_t i = 0; i < l; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            HFr a1 = h_zero(), a2 = h_zero();
            HFr si = h_mul(g, h_pow(zk, i));
            HFr ui = h_mul(g, h_pow(z2k, 2ull * i));
            for (uint32_t d = 0; d < n; ++d) {
                HFr w = h_pow(zn_inv, (uint64_t)j * d);
                if (d < k) a1 = h_add(a1, h_mul(w, h_pow(si, d)));
                a2 = h_add(a2, h_mul(w, h_pow(ui, d)));
            }
            out->unpack[(size_t)i * n + j] = h_mul(a1, n_inv);
            out->unpack2[(size_t)i * n + j] = h_mul(a2, n_inv);
        }
}

// lagrange_unpack (pss.rs:170-207) for the received party subset: l x k matrix M with
// secrets_i = sum_r M[i][r] * shares_r = sum_r L_r(u_{2i}) * shares_r, L_r the Lagrange basis on
// the points x_r = zeta_n^{parties[r]}.  (Interpolate-then-evaluate collapsed into one map.)
inline bool pss_lagrange_matrix(uint32_t l, const uint32_t* parties, uint32_t k, std::vector<HFr>* out) {
    const uint32_t t = l, n = 4 * l;
    if (!(k > 2 * (t + l - 1)) || k > n) return false;
    HFr g = h_load(BN254_FR_GENERATOR_MONT_64);
    HFr zn = h_root_of_unity(n), z2k = h_root_of_unity(2 * (l + t));
    std::vector<HFr> xs(k);
    for (uint32_t r = 0; r < k; ++r) {
        if (parties[r] >= n) return false;
        for (uint32_t q = 0; q < r; ++q) if (parties[q] == parties[r]) return false;
        xs[r] = h_pow(zn, parties[r]);
    }
    out->assign((size_t)l * k, h_zero());
    for (uint32_t i = 0; i < l; ++i) {
        HFr u = h_mul(g, h_pow(z2k, 2ull * i));
        for (uint32_t r = 0; r < k; ++r) {
            HFr num = h_one(), den = h_one();
            for (uint32_t q = 0; q < k; ++q) {
                if (q == r) continue;
                num = h_mul(num, h_sub(u, xs[q]));
                den = h_mul(den, h_sub(xs[r], xs[q]));
            }
            (*out)[(size_t)i * k + r] = h_mul(num, h_inv(den));
        }
    }
    return true;
}

}  // namespace host
}  // namespace zkg
