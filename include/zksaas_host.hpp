// zksaas_host.hpp -- C++17 host-side mirror of the reference's interface for the hot path, over the C ABI of
// include/zksaas_gpu.h.  Header-only; link with -lzksaas_gpu.
//
// The reference is compiled (Rust) code and there is no Rust toolchain in this image, so this is the host layer a
// maintainer would otherwise write in Rust (INTEGRATION.md shows that binding): the same names, argument meaning and
// error behaviour as the reference's functions on the path, one C-ABI call per reference call, no arithmetic of its
// own beyond the O(log n) field operations a `Radix2EvaluationDomain` needs for its constants.
//
//   reference                                                    here
//   ark_bn254::{Fr, G1Affine, G2Affine, G1Projective, ...}       Fr, G1Affine, G2Affine, G1Projective, G2Projective
//                                                                (the SAME memory images: 4xu64 Montgomery, 72 / 136 B affine)
//   ark_poly::Radix2EvaluationDomain                             Radix2EvaluationDomain        (fft / ifft on the device)
//   secret-sharing/src/pss.rs:19-166  PackedSharingParams        PackedSharingParams           (pack, det_pack, unpack, unpack2)
//   dist-primitives/src/utils/pack.rs:8-35                       pack_vec, transpose
//   dist-primitives/src/dfft/mod.rs:16-95     FftMask            FftMask::sample               (random draws passed in)
//   dist-primitives/src/dfft/mod.rs:99-175    d_fft / d_ifft     d_fft, d_ifft                 (all parties of a LocalTestNet at once)
//   dist-primitives/src/dfft/mod.rs:178-335   fft1 / fft2 / ...  fft1_in_place, fft2_in_place, fft_in_place_rearrange
//   dist-primitives/src/utils/deg_red.rs:14-126                  DegRedMask::sample, deg_red
//   dist-primitives/src/dpp/mod.rs:15-87      d_pp               d_pp
//   groth16/src/proving_key.rs:72-104, qap.rs:99-112            crs_det_pack<G>, qap_pss_pack
//   groth16/src/ext_wit.rs:14-181   libsnark_h, circom_h         libsnark_h, circom_h
//   dist-primitives/src/dmsm/mod.rs:10-102    MsmMask, d_msm     MsmMask<G>::sample, d_msm<G>
//   G::msm(bases, scalars) -> Result<G, usize>                   msm<G>(bases, scalars)        (throws MsmLengthMismatch{min_len})
//   mpc-net/src/multi.rs LocalTestNet (+ lossy round :330-363)   LocalTestNet{n, dropouts}
//
// The host RNG stays with the caller, as it stays in Rust: every function that draws randomness in the reference
// (`pack`, the mask samplers) takes the draws as arguments, so results are comparable bit for bit.
// Every compute call fails loudly (zksaas::Error) without the CUDA library or a device: there is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "zksaas_gpu.h"

namespace zksaas {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};
// `G::msm` returns Err(min(bases.len(), scalars.len())) on a length mismatch (ark-ec 0.4.2; dmsm/mod.rs:73 `?`)
struct MsmLengthMismatch : std::runtime_error {
    size_t min_len;
    explicit MsmLengthMismatch(size_t m) : std::runtime_error("msm length mismatch"), min_len(m) {}
};
inline void check(int32_t rc) {
    if (rc != ZKG_OK) throw Error(rc, zkg_last_error());
}

// ---------------------------------------------------------------------------------------------
// ark_bn254::Fr as its memory image (Montgomery residue a * 2^256 mod r, 4 little-endian u64 limbs).  Host arithmetic
// is only used for domain constants (roots of unity, inverses of sizes, coset shifts).
// ---------------------------------------------------------------------------------------------
struct Fr {
    uint64_t v[4];
    static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr uint64_t R1[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
    static constexpr uint64_t INV = 0xc2e1f593efffffffULL;                  // -r^-1 mod 2^64
    // F::TWO_ADIC_ROOT_OF_UNITY (2^28-th root), Montgomery image
    static constexpr uint64_t ROOT[4] = {0x636e735580d13d9cULL, 0xa22bf3742445ffd6ULL, 0x56452ac01eb203d8ULL, 0x1860ef942963f9e7ULL};
    static constexpr int TWO_ADICITY = 28;

    static Fr zero() { return Fr{{0, 0, 0, 0}}; }
    static Fr one() { return Fr{{R1[0], R1[1], R1[2], R1[3]}}; }
    static Fr from_u64(uint64_t x) { Fr a{{x, 0, 0, 0}}, r2{{R2[0], R2[1], R2[2], R2[3]}}; return a * r2; }
    static Fr generator() { return from_u64(5); }                           // F::GENERATOR
    bool operator==(const Fr& o) const { return std::memcmp(v, o.v, 32) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }

    static bool geq_mod(const uint64_t* a) {
        for (int i = 3; i >= 0; --i) { if (a[i] != MOD[i]) return a[i] > MOD[i]; }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        unsigned __int128 b = 0;
        for (int i = 0; i < 4; ++i) { unsigned __int128 d = (unsigned __int128)a[i] - MOD[i] - (uint64_t)b; a[i] = (uint64_t)d; b = (d >> 64) & 1; }
    }
    Fr operator+(const Fr& o) const {
        Fr r; unsigned __int128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (unsigned __int128)v[i] + o.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    Fr operator-() const {
        if (is_zero()) return *this;
        Fr r; unsigned __int128 b = 0;
        for (int i = 0; i < 4; ++i) { unsigned __int128 d = (unsigned __int128)MOD[i] - v[i] - (uint64_t)b; r.v[i] = (uint64_t)d; b = (d >> 64) & 1; }
        return r;
    }
    Fr operator-(const Fr& o) const { return *this + (-o); }
    Fr operator*(const Fr& o) const {                                       // CIOS Montgomery product
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; ++j) { c += (unsigned __int128)v[j] * o.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * INV;
            c = (unsigned __int128)m * MOD[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) { c += (unsigned __int128)m * MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fr r{{t[0], t[1], t[2], t[3]}};
        if (t[4] || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    Fr pow(const uint64_t* e, int limbs) const {
        Fr r = one();
        for (int i = limbs - 1; i >= 0; --i)
            for (int b = 63; b >= 0; --b) { r = r * r; if ((e[i] >> b) & 1) r = r * *this; }
        return r;
    }
    Fr pow(uint64_t e) const { return pow(&e, 1); }
    Fr inverse() const {                                                    // a^(r-2)
        uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
        return pow(e, 4);
    }
};
static_assert(sizeof(Fr) == 32, "Fr image");

// ark_ec::short_weierstrass::{Affine, Projective} images: Affine {x, y, infinity: bool} padded to 8 bytes; Projective = Jacobian
template <int W>
struct Affine {
    uint64_t x[W], y[W];
    uint8_t infinity;
    uint8_t pad[7];
};
template <int W>
struct Projective {
    uint64_t x[W], y[W], z[W];
    bool operator==(const Projective& o) const { return std::memcmp(this, &o, sizeof *this) == 0; }   // results are normalised (Z = 1)
    bool is_identity() const { for (int i = 0; i < W; ++i) if (z[i]) return false; return true; }
};
using G1Affine = Affine<4>;
using G2Affine = Affine<8>;
using G1Projective = Projective<4>;
using G2Projective = Projective<8>;
static_assert(sizeof(G1Affine) == ZKG_G1_AFFINE_BYTES && sizeof(G2Affine) == ZKG_G2_AFFINE_BYTES, "arkworks affine images");
static_assert(sizeof(G1Projective) == 96 && sizeof(G2Projective) == 192, "arkworks projective images");

namespace detail {
constexpr uint64_t FQ_ONE[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
template <int W> struct group_tag { static constexpr int32_t id = W == 4 ? 1 : 2; };
}  // namespace detail

// Projective (normalised) -> Affine image, `into_affine()` of a normalised point
template <int W>
inline Affine<W> into_affine(const Projective<W>& p) {
    Affine<W> a;
    std::memset(&a, 0, sizeof a);
    if (p.is_identity()) { a.infinity = 1; return a; }
    std::memcpy(a.x, p.x, sizeof a.x);
    std::memcpy(a.y, p.y, sizeof a.y);
    return a;
}
// Affine -> Projective (`into_group()`): identity = (1, 1, 0)
template <int W>
inline Projective<W> into_group(const Affine<W>& a) {
    Projective<W> p;
    std::memset(&p, 0, sizeof p);
    if (a.infinity) { std::memcpy(p.x, detail::FQ_ONE, 32); std::memcpy(p.y, detail::FQ_ONE, 32); return p; }
    std::memcpy(p.x, a.x, sizeof p.x);
    std::memcpy(p.y, a.y, sizeof p.y);
    std::memcpy(p.z, detail::FQ_ONE, 32);
    return p;
}
template <int W>
inline Projective<W> identity() { Affine<W> a; std::memset(&a, 0, sizeof a); a.infinity = 1; return into_group(a); }

// ---------------------------------------------------------------------------------------------
// `G::msm(bases, scalars)` -- ark-ec VariableBaseMSM::msm at dist-primitives/src/dmsm/mod.rs:73
// ---------------------------------------------------------------------------------------------
template <int W>
inline Projective<W> msm(const std::vector<Affine<W>>& bases, const std::vector<Fr>& scalars, int32_t device = 0) {
    Projective<W> out;
    auto fn = W == 4 ? zkg_msm_bn254_g1 : zkg_msm_bn254_g2;
    int32_t rc = fn(device, bases.data(), sizeof(Affine<W>), bases.size(), (const uint64_t*)scalars.data(), scalars.size(), (uint64_t*)&out);
    if (rc == ZKG_ERR_LEN_MISMATCH) throw MsmLengthMismatch(bases.size() < scalars.size() ? bases.size() : scalars.size());
    check(rc);
    return out;
}
// sum_i coeffs[i] * points[i] (a tiny MSM): the group operations around d_msm in groth16/src/prove.rs
template <int W>
inline Projective<W> lincomb(const std::vector<Projective<W>>& points, const std::vector<Fr>& coeffs, int32_t device = 0) {
    std::vector<Affine<W>> aff;
    for (const auto& p : points) aff.push_back(into_affine(p));
    return msm<W>(aff, coeffs, device);
}
template <int W>
inline Projective<W> add(const Projective<W>& a, const Projective<W>& b, int32_t device = 0) {
    return lincomb<W>({a, b}, {Fr::one(), Fr::one()}, device);
}

// pack_from_arkworks_proving_key's inner step (groth16/src/proving_key.rs:72-104): `pp.det_pack::<G>(chunk)` for every l-chunk of
// a CRS query (bases.size() a multiple of l; pad a short last chunk with the identity), returned as the n parties' share vectors
template <int W>
inline std::vector<std::vector<Affine<W>>> crs_det_pack(const std::vector<Affine<W>>& bases, uint32_t l, int32_t device = 0) {
    if (bases.size() % l) throw Error(ZKG_ERR_BAD_ARG, "crs_det_pack: a multiple of l bases expected");
    const size_t chunks = bases.size() / l;
    std::vector<std::vector<Affine<W>>> out(4 * l, std::vector<Affine<W>>(chunks));
    std::vector<void*> p;
    for (auto& v : out) p.push_back(v.data());
    check(zkg_crs_det_pack_bn254(device, detail::group_tag<W>::id, bases.data(), sizeof(Affine<W>), bases.size(), l, p.data(), sizeof(Affine<W>)));
    return out;
}

// ---------------------------------------------------------------------------------------------
// ark_poly::Radix2EvaluationDomain<Fr>
// ---------------------------------------------------------------------------------------------
struct Radix2EvaluationDomain {
    size_t size_;
    int log_size;
    Fr group_gen_, group_gen_inv_, size_inv_;
    static Radix2EvaluationDomain new_(size_t num_coeffs) {                 // `new` is a keyword
        Radix2EvaluationDomain d;
        d.log_size = 0;
        while (((size_t)1 << d.log_size) < num_coeffs) ++d.log_size;       // ark_std::log2 = ceil
        if (d.log_size > Fr::TWO_ADICITY) throw Error(ZKG_ERR_BAD_ARG, "domain larger than 2^28");
        d.size_ = (size_t)1 << d.log_size;
        Fr g{{Fr::ROOT[0], Fr::ROOT[1], Fr::ROOT[2], Fr::ROOT[3]}};
        for (int i = d.log_size; i < Fr::TWO_ADICITY; ++i) g = g * g;      // get_root_of_unity(size)
        d.group_gen_ = g;
        d.group_gen_inv_ = g.inverse();
        d.size_inv_ = Fr::from_u64(d.size_).inverse();
        return d;
    }
    size_t size() const { return size_; }
    Fr group_gen() const { return group_gen_; }
    Fr group_gen_inv() const { return group_gen_inv_; }
    Fr size_inv() const { return size_inv_; }
    Fr element(size_t i) const { return group_gen_.pow((uint64_t)i); }
    // fft_in_place / ifft_in_place (resize to the domain size, optional coset offset), on the device
    void fft_in_place(std::vector<Fr>& v, const Fr* offset = nullptr, int32_t device = 0) const {
        v.resize(size_, Fr::zero());
        check(zkg_fr_fft_bn254(device, (uint64_t*)v.data(), size_, offset ? offset->v : nullptr, 0));
    }
    void ifft_in_place(std::vector<Fr>& v, const Fr* offset = nullptr, int32_t device = 0) const {
        v.resize(size_, Fr::zero());
        check(zkg_fr_fft_bn254(device, (uint64_t*)v.data(), size_, offset ? offset->v : nullptr, 1));
    }
};

// ---------------------------------------------------------------------------------------------
// secret-sharing/src/pss.rs  PackedSharingParams  (:19-66 new, :69-87 det_pack, :90-122 pack, :125-138 unpack, :141-166 unpack2)
// Batched over columns: `secrets` holds cols*l values, `rand_points` cols*t, `shares` cols*n (column-major).
// ---------------------------------------------------------------------------------------------
struct PackedSharingParams {
    uint32_t t, l, n;
    int32_t device;
    static PackedSharingParams new_(uint32_t l, int32_t device = 0) { return PackedSharingParams{l, l, 4 * l, device}; }
    std::vector<Fr> pack(const std::vector<Fr>& secrets, const std::vector<Fr>& rand_points) const {
        if (secrets.size() % l || rand_points.size() != secrets.size() / l * t) throw Error(ZKG_ERR_BAD_ARG, "Secrets length mismatch");
        const size_t cols = secrets.size() / l;
        std::vector<Fr> shares(cols * n);
        check(zkg_pss_pack_bn254_fr(device, l, (const uint64_t*)secrets.data(), (const uint64_t*)rand_points.data(), (uint64_t*)shares.data(), cols));
        return shares;
    }
    std::vector<Fr> det_pack(const std::vector<Fr>& secrets) const {
        if (secrets.size() % l) throw Error(ZKG_ERR_BAD_ARG, "Secrets length mismatch");
        const size_t cols = secrets.size() / l;
        std::vector<Fr> shares(cols * n);
        check(zkg_pss_pack_bn254_fr(device, l, (const uint64_t*)secrets.data(), nullptr, (uint64_t*)shares.data(), cols));
        return shares;
    }
    std::vector<Fr> unpack(const std::vector<Fr>& shares) const { return unpack_with(shares, zkg_pss_unpack_bn254_fr); }
    std::vector<Fr> unpack2(const std::vector<Fr>& shares) const { return unpack_with(shares, zkg_pss_unpack2_bn254_fr); }
    // unpack_missing_shares over GROUP elements + the sum (dmsm/mod.rs:85-86; sha256.rs:375-377): returns the l unpacked
    // points, *sum (nullable) their sum
    template <int W>
    std::vector<Projective<W>> unpack_missing_shares(const std::vector<Projective<W>>& shares, const std::vector<uint32_t>& parties,
                                                     Projective<W>* sum = nullptr) const {
        if (shares.size() != parties.size()) throw Error(ZKG_ERR_BAD_ARG, "one share per received party expected");
        std::vector<Projective<W>> out(l);
        auto fn = W == 4 ? zkg_pss_unpack2_bn254_g1 : zkg_pss_unpack2_bn254_g2;
        check(fn(device, l, (const uint64_t*)shares.data(), parties.data(), (uint32_t)parties.size(), (uint64_t*)out.data(), (uint64_t*)sum));
        return out;
    }

  private:
    template <class FN>
    std::vector<Fr> unpack_with(const std::vector<Fr>& shares, FN fn) const {
        if (shares.size() % n) throw Error(ZKG_ERR_BAD_ARG, "Shares length mismatch");
        const size_t cols = shares.size() / n;
        std::vector<Fr> secrets(cols * l);
        check(fn(device, l, (const uint64_t*)shares.data(), (uint64_t*)secrets.data(), cols));
        return secrets;
    }
};

using Shares = std::vector<std::vector<Fr>>;                                // party-major: [party][column]

namespace detail {
inline std::vector<uint64_t*> ptrs(Shares& s) {
    std::vector<uint64_t*> p;
    for (auto& v : s) p.push_back((uint64_t*)v.data());
    return p;
}
inline std::vector<const uint64_t*> cptrs(const Shares& s) {
    std::vector<const uint64_t*> p;
    for (auto& v : s) p.push_back((const uint64_t*)v.data());
    return p;
}
}  // namespace detail

// dist-primitives/src/utils/pack.rs:22-35
template <class T>
inline std::vector<std::vector<T>> transpose(const std::vector<std::vector<T>>& m) {
    std::vector<std::vector<T>> r(m.empty() ? 0 : m[0].size(), std::vector<T>(m.size()));
    for (size_t i = 0; i < m.size(); ++i)
        for (size_t j = 0; j < m[i].size(); ++j) r[j][i] = m[i][j];
    return r;
}
// pack_vec (pack.rs:8-20) followed by the transpose every caller applies: the n parties' share vectors of the l-chunks of
// `secrets` (last chunk zero-padded as pack_from_witness does, groth16/examples/sha256.rs:131-156)
inline Shares pack_vec(const std::vector<Fr>& secrets, const PackedSharingParams& pp, const std::vector<Fr>& rand_points) {
    const size_t chunks = (secrets.size() + pp.l - 1) / pp.l;
    if (rand_points.size() != chunks * pp.t) throw Error(ZKG_ERR_BAD_ARG, "pack_vec: ceil(len/l) * t random points expected");
    Shares out(pp.n, std::vector<Fr>(chunks));
    auto p = detail::ptrs(out);
    check(zkg_pss_pack_vec_bn254_fr(pp.device, pp.l, 0, (const uint64_t*)secrets.data(), secrets.size(), (const uint64_t*)rand_points.data(), p.data()));
    return out;
}
// the `pack` closure of QAP::pss (groth16/src/qap.rs:99-112): bit-reverse, column i packs x[i], x[i + m/l], ...
inline Shares qap_pss_pack(const std::vector<Fr>& x, const PackedSharingParams& pp, const std::vector<Fr>& rand_points) {
    const size_t chunks = x.size() / pp.l;
    if (x.size() % pp.l || rand_points.size() != chunks * pp.t) throw Error(ZKG_ERR_BAD_ARG, "qap_pss_pack: len/l * t random points expected");
    Shares out(pp.n, std::vector<Fr>(chunks));
    auto p = detail::ptrs(out);
    check(zkg_pss_pack_vec_bn254_fr(pp.device, pp.l, 1, (const uint64_t*)x.data(), x.size(), (const uint64_t*)rand_points.data(), p.data()));
    return out;
}

// dist-primitives/src/dfft/mod.rs:322-335, :178-208, :210-237; Radix2EvaluationDomain::distribute_powers (:49, :279)
inline void fft_in_place_rearrange(std::vector<Fr>& data, int32_t device = 0) { check(zkg_bitrev_bn254(device, (uint64_t*)data.data(), data.size())); }
inline void fft1_in_place(std::vector<Fr>& px, const PackedSharingParams& pp, const Fr& gen, const Fr* pre_scale = nullptr,
                          const std::vector<Fr>* in_mask = nullptr, int32_t device = -1) {
    if (in_mask && in_mask->size() != px.size()) throw Error(ZKG_ERR_BAD_ARG, "fft1: mask length differs from the share vector");
    check(zkg_fft1_bn254(device < 0 ? pp.device : device, (uint64_t*)px.data(), px.size(), pp.l, gen.v, pre_scale ? pre_scale->v : nullptr,
                         in_mask ? (const uint64_t*)in_mask->data() : nullptr));
}
inline void fft2_in_place(std::vector<Fr>& s1, const PackedSharingParams& pp, const Fr& gen) {
    check(zkg_fft2_bn254(pp.device, (uint64_t*)s1.data(), s1.size(), pp.l, gen.v));
}
inline void distribute_powers(std::vector<Fr>& v, const Fr& g, int32_t device = 0) {
    check(zkg_distribute_powers_bn254(device, (uint64_t*)v.data(), v.size(), g.v));
}

// ---------------------------------------------------------------------------------------------
// masks
// ---------------------------------------------------------------------------------------------
struct FftMask {                                                            // dfft/mod.rs:16-95 (one party's share)
    std::vector<Fr> in_mask, out_mask;
    static FftMask zero(size_t mbyl) { return FftMask{std::vector<Fr>(mbyl, Fr::zero()), std::vector<Fr>(mbyl, Fr::zero())}; }
    // :30-85 with the draws passed in: mask_values (m), rand_in / rand_out (m/l * t each).  Returns the n parties' masks.
    static std::vector<FftMask> sample(bool rearrange, const Fr& g, const Fr& gen, size_t m, const PackedSharingParams& pp,
                                       const std::vector<Fr>& mask_values, const std::vector<Fr>& rand_in, const std::vector<Fr>& rand_out) {
        const size_t mbyl = m / pp.l;
        if (mask_values.size() != m || rand_in.size() != mbyl * pp.t || rand_out.size() != mbyl * pp.t)
            throw Error(ZKG_ERR_BAD_ARG, "FftMask::sample: m mask values and m/l * t + m/l * t packing draws expected");
        Shares in(pp.n, std::vector<Fr>(mbyl)), out(pp.n, std::vector<Fr>(mbyl));
        auto pi = detail::ptrs(in), po = detail::ptrs(out);
        check(zkg_fft_mask_sample_bn254(pp.device, rearrange ? 1 : 0, g.v, gen.v, m, pp.l, (const uint64_t*)mask_values.data(),
                                        (const uint64_t*)rand_in.data(), (const uint64_t*)rand_out.data(), pi.data(), po.data()));
        std::vector<FftMask> r;
        for (uint32_t p = 0; p < pp.n; ++p) r.push_back(FftMask{std::move(in[p]), std::move(out[p])});
        return r;
    }
};
struct DegRedMask {                                                         // utils/deg_red.rs:14-77 over Fr, gen = 1
    std::vector<Fr> in_mask, out_mask;
    static std::vector<DegRedMask> sample(const PackedSharingParams& pp, size_t num, const std::vector<Fr>& mask_values,
                                          const std::vector<Fr>& rand_in, const std::vector<Fr>& rand_out) {
        if (mask_values.size() != num * pp.l || rand_in.size() != num * pp.t || rand_out.size() != num * pp.t)
            throw Error(ZKG_ERR_BAD_ARG, "DegRedMask::sample: num * l mask values and num * t + num * t packing draws expected");
        Shares in(pp.n, std::vector<Fr>(num)), out(pp.n, std::vector<Fr>(num));
        auto pi = detail::ptrs(in), po = detail::ptrs(out);
        check(zkg_deg_red_mask_sample_bn254(pp.device, num, pp.l, (const uint64_t*)mask_values.data(), (const uint64_t*)rand_in.data(),
                                            (const uint64_t*)rand_out.data(), pi.data(), po.data()));
        std::vector<DegRedMask> r;
        for (uint32_t p = 0; p < pp.n; ++p) r.push_back(DegRedMask{std::move(in[p]), std::move(out[p])});
        return r;
    }
};
// rows of the pack matrix (n x (l + t)): shares = M (secrets || rand), read off the device by packing unit vectors
inline std::vector<std::vector<Fr>> pack_matrix(const PackedSharingParams& pp) {
    const uint32_t k = pp.l + pp.t;
    std::vector<std::vector<Fr>> M(pp.n, std::vector<Fr>(k));
    for (uint32_t j = 0; j < k; ++j) {
        std::vector<Fr> sec(pp.l, Fr::zero()), rnd(pp.t, Fr::zero());
        (j < pp.l ? sec[j] : rnd[j - pp.l]) = Fr::one();
        auto col = pp.pack(sec, rnd);
        for (uint32_t i = 0; i < pp.n; ++i) M[i][j] = col[i];
    }
    return M;
}
template <int W>
struct MsmMask {                                                            // dmsm/mod.rs:10-57
    Projective<W> in_mask, out_mask;
    static MsmMask zero() { return MsmMask{identity<W>(), identity<W>()}; }
    // :21-48: mask value i = gen * mask_scalars[i]; in = pack(values), out = pack([-(sum of the values); l]); the t random GROUP
    // elements each `pp.pack` call draws are passed in.  `gen` = G::generator().
    static std::vector<MsmMask> sample(const PackedSharingParams& pp, const Projective<W>& gen, const std::vector<Fr>& mask_scalars,
                                       const std::vector<Projective<W>>& rand_in, const std::vector<Projective<W>>& rand_out) {
        if (mask_scalars.size() != pp.l || rand_in.size() != pp.t || rand_out.size() != pp.t)
            throw Error(ZKG_ERR_BAD_ARG, "MsmMask::sample: l mask scalars and t + t random points expected");
        std::vector<Projective<W>> values;
        for (const Fr& x : mask_scalars) values.push_back(lincomb<W>({gen}, {x}, pp.device));
        Projective<W> out_value = lincomb<W>(values, std::vector<Fr>(pp.l, -Fr::one()), pp.device);
        auto M = pack_matrix(pp);
        std::vector<Projective<W>> in_pts = values, out_pts(pp.l, out_value);
        in_pts.insert(in_pts.end(), rand_in.begin(), rand_in.end());
        out_pts.insert(out_pts.end(), rand_out.begin(), rand_out.end());
        std::vector<MsmMask> r;
        for (uint32_t i = 0; i < pp.n; ++i) r.push_back(MsmMask{lincomb<W>(in_pts, M[i], pp.device), lincomb<W>(out_pts, M[i], pp.device)});
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// N parties in one process; party 0 is the king (mpc-net/src/lib.rs:65-67).  `dropouts`: parties whose message the king
// does not receive (simulate_lossy_network_round, mpc-net/src/multi.rs:330-363).
// ---------------------------------------------------------------------------------------------
struct LocalTestNet {
    uint32_t n;
    std::vector<uint32_t> dropouts;
    uint32_t n_parties() const { return n; }
    std::vector<uint32_t> parties() const {
        std::vector<uint32_t> p;
        for (uint32_t i = 0; i < n; ++i) {
            bool dropped = false;
            for (uint32_t d : dropouts) dropped |= d == i;
            if (!dropped) p.push_back(i);
        }
        return p;
    }
};

// king closures: dfft/mod.rs:264-304, deg_red.rs:103-111
inline Shares king_fft2(const Shares& recv, const std::vector<uint32_t>& parties, const PackedSharingParams& pp, const Fr& gen, const Fr& g,
                        bool rearrange, const std::vector<Fr>& rand_points) {
    if (recv.empty() || recv.size() != parties.size()) throw Error(ZKG_ERR_BAD_ARG, "king_fft2: one share vector per received party expected");
    const size_t mbyl = recv[0].size();
    for (auto& v : recv) if (v.size() != mbyl) throw Error(ZKG_ERR_BAD_ARG, "king_fft2: share vectors of different lengths");
    if (rand_points.size() != mbyl * pp.t) throw Error(ZKG_ERR_BAD_ARG, "king_fft2: m/l * t random points expected");
    Shares out(pp.n, std::vector<Fr>(mbyl));
    auto pi = detail::cptrs(recv);
    auto po = detail::ptrs(out);
    check(zkg_king_fft2_bn254(pp.device, pi.data(), parties.data(), (uint32_t)parties.size(), mbyl, pp.l, gen.v, g.v, rearrange ? 1 : 0,
                              (const uint64_t*)rand_points.data(), po.data()));
    return out;
}
inline Shares deg_red_king(const Shares& recv, const std::vector<uint32_t>& parties, const PackedSharingParams& pp, const std::vector<Fr>& rand_points) {
    if (recv.empty() || recv.size() != parties.size()) throw Error(ZKG_ERR_BAD_ARG, "deg_red_king: one share vector per received party expected");
    const size_t cols = recv[0].size();
    if (rand_points.size() != cols * pp.t) throw Error(ZKG_ERR_BAD_ARG, "deg_red_king: cols * t random points expected");
    Shares out(pp.n, std::vector<Fr>(cols));
    auto pi = detail::cptrs(recv);
    auto po = detail::ptrs(out);
    check(zkg_deg_red_king_bn254(pp.device, pi.data(), parties.data(), (uint32_t)parties.size(), cols, pp.l, (const uint64_t*)rand_points.data(), po.data()));
    return out;
}

namespace detail {
inline std::vector<Fr> vec_add(const std::vector<Fr>& a, const std::vector<Fr>& b, int32_t device) {
    if (a.size() != b.size()) throw Error(ZKG_ERR_BAD_ARG, "mask length differs from the share vector");
    std::vector<Fr> r(a.size());
    check(zkg_field_op(device, 0, 1, (const uint64_t*)a.data(), (const uint64_t*)b.data(), (uint64_t*)r.data(), a.size()));
    return r;
}
inline Shares d_fft_impl(const Shares& shares, const std::vector<FftMask>& masks, bool rearrange, const Radix2EvaluationDomain& dom, const Fr& g,
                         const PackedSharingParams& pp, const LocalTestNet& net, const std::vector<Fr>& rand_points, bool inverse) {
    if (shares.size() != net.n || masks.size() != net.n) throw Error(ZKG_ERR_BAD_ARG, "d_fft: one share vector and one mask per party expected");
    const size_t mbyl = shares[0].size();
    if (mbyl * pp.l != dom.size()) throw Error(ZKG_ERR_BAD_ARG, "Mismatch of size in FFT");          // dfft/mod.rs:112-118
    const Fr gen = inverse ? dom.group_gen_inv() : dom.group_gen();
    const Fr pre = dom.size_inv();
    Shares sent(net.n);
    for (uint32_t p = 0; p < net.n; ++p) {
        sent[p] = shares[p];
        fft1_in_place(sent[p], pp, gen, inverse ? &pre : nullptr, &masks[p].in_mask);               // :121 / :159-162, :254-258
    }
    const auto parties = net.parties();
    Shares recv;
    for (uint32_t p : parties) recv.push_back(sent[p]);
    Shares out = king_fft2(recv, parties, pp, gen, g, rearrange, rand_points);                      // :264-304
    for (uint32_t p = 0; p < net.n; ++p) out[p] = vec_add(out[p], masks[p].out_mask, pp.device);   // :313-317
    return out;
}
}  // namespace detail

// dist-primitives/src/dfft/mod.rs:99-134 / :137-175, for all parties of the net at once: returns the n output share vectors
inline Shares d_fft(const Shares& pcoeff_shares, const std::vector<FftMask>& masks, bool rearrange, const Radix2EvaluationDomain& dom,
                    const PackedSharingParams& pp, const LocalTestNet& net, const std::vector<Fr>& rand_points) {
    return detail::d_fft_impl(pcoeff_shares, masks, rearrange, dom, Fr::one(), pp, net, rand_points, false);
}
inline Shares d_ifft(const Shares& peval_shares, const std::vector<FftMask>& masks, bool rearrange, const Radix2EvaluationDomain& dom, const Fr& g,
                     const PackedSharingParams& pp, const LocalTestNet& net, const std::vector<Fr>& rand_points) {
    return detail::d_fft_impl(peval_shares, masks, rearrange, dom, g, pp, net, rand_points, true);
}
// dist-primitives/src/utils/deg_red.rs:80-126
inline Shares deg_red(const Shares& x_shares, const std::vector<DegRedMask>& masks, const PackedSharingParams& pp, const LocalTestNet& net,
                      const std::vector<Fr>& rand_points) {
    Shares sent(net.n);
    for (uint32_t p = 0; p < net.n; ++p) sent[p] = detail::vec_add(x_shares[p], masks[p].in_mask, pp.device);
    const auto parties = net.parties();
    Shares recv;
    for (uint32_t p : parties) recv.push_back(sent[p]);
    Shares out = deg_red_king(recv, parties, pp, rand_points);
    for (uint32_t p = 0; p < net.n; ++p) out[p] = detail::vec_add(out[p], masks[p].out_mask, pp.device);
    return out;
}
// dist-primitives/src/dpp/mod.rs:15-87 (partial products of num / den; the dummy randomness s = 1 of :24-25 included):
// the king unpacks, divides (a zero denominator is ZKG_ERR_BAD_ARG where the reference's `.inverse().unwrap()` panics),
// takes the running product over all secrets and re-packs (:41-76); then deg_red (:86)
inline Shares dpp_king(const Shares& recv, const std::vector<uint32_t>& parties, const PackedSharingParams& pp, const std::vector<Fr>& rand_points) {
    if (recv.empty() || recv.size() != parties.size() || recv[0].size() % 2) throw Error(ZKG_ERR_BAD_ARG, "dpp_king: num || den share vectors expected");
    const size_t cols = recv[0].size() / 2;
    if (rand_points.size() != cols * pp.t) throw Error(ZKG_ERR_BAD_ARG, "dpp_king: cols * t random points expected");
    Shares out(pp.n, std::vector<Fr>(cols));
    auto pi = detail::cptrs(recv);
    auto po = detail::ptrs(out);
    check(zkg_dpp_king_bn254(pp.device, pi.data(), parties.data(), (uint32_t)parties.size(), cols, pp.l, (const uint64_t*)rand_points.data(), po.data()));
    return out;
}
inline Shares d_pp(const Shares& num, const Shares& den, const std::vector<DegRedMask>& masks, const PackedSharingParams& pp, const LocalTestNet& net,
                   const std::vector<Fr>& rand_king, const std::vector<Fr>& rand_degred) {
    Shares sent(net.n);
    for (uint32_t p = 0; p < net.n; ++p) { sent[p] = num[p]; sent[p].insert(sent[p].end(), den[p].begin(), den[p].end()); }    // :31-32
    const auto parties = net.parties();
    Shares recv;
    for (uint32_t p : parties) recv.push_back(sent[p]);
    return deg_red(dpp_king(recv, parties, pp, rand_king), masks, pp, net, rand_degred);
}
// share-wise h = (a * b - c) [* factor]: groth16/src/ext_wit.rs:173-177 / :82-86
inline std::vector<Fr> qap_h(const std::vector<Fr>& a, const std::vector<Fr>& b, const std::vector<Fr>& c, const Fr* factor = nullptr, int32_t device = 0) {
    if (a.size() != b.size() || a.size() != c.size()) throw Error(ZKG_ERR_BAD_ARG, "a, b, c of different lengths");
    std::vector<Fr> out(a.size());
    check(zkg_qap_h_bn254(device, (const uint64_t*)a.data(), (const uint64_t*)b.data(), (const uint64_t*)c.data(), nullptr, nullptr, nullptr,
                          factor ? factor->v : nullptr, (uint64_t*)out.data(), a.size()));
    return out;
}

// groth16/src/ext_wit.rs:104-181 `circom_h`: three d_ifft (coset shift w_2m on the way out, rearranged), three d_fft,
// h = a*b - c share-wise, deg_red.  qap_shares[party] = {a, b, c} share vectors of a PackedQAPShare (qap.rs:30-41);
// fft_masks: 3 ifft + 3 fft masks, each for all parties; rand: the king's packing draws of the seven rounds (m/l * t each).
struct PackedQAPShare { std::vector<Fr> a, b, c; };
inline Shares circom_h(const std::vector<PackedQAPShare>& qap_shares, const std::vector<std::vector<FftMask>>& fft_masks,
                       const std::vector<DegRedMask>& degred_mask, const Radix2EvaluationDomain& domain, const PackedSharingParams& pp,
                       const LocalTestNet& net, const std::vector<std::vector<Fr>>& rand) {
    if (qap_shares.size() != net.n || fft_masks.size() != 6 || rand.size() != 7) throw Error(ZKG_ERR_BAD_ARG, "circom_h: n QAP shares, 6 FFT masks, 7 draw vectors expected");
    const Fr root_of_unity = Radix2EvaluationDomain::new_(2 * domain.size()).element(1);          // :117-122
    Shares in[3];
    for (const auto& q : qap_shares) { in[0].push_back(q.a); in[1].push_back(q.b); in[2].push_back(q.c); }
    Shares ev[3];
    for (int k = 0; k < 3; ++k) {
        Shares coeff = d_ifft(in[k], fft_masks[k], true, domain, root_of_unity, pp, net, rand[k]);   // :124-159
        ev[k] = d_fft(coeff, fft_masks[3 + k], false, domain, pp, net, rand[3 + k]);                 // :161-170
    }
    Shares h(net.n);
    for (uint32_t p = 0; p < net.n; ++p) h[p] = qap_h(ev[0][p], ev[1][p], ev[2][p], nullptr, pp.device);   // :173-177
    return deg_red(h, degred_mask, pp, net, rand[6]);                                               // :179
}
// groth16/src/ext_wit.rs:14-102 `libsnark_h`: coset (offset = F::GENERATOR) d_ifft x3 and d_fft x3 (all rearranged),
// h = (a*b - c) / Z(g), coset d_ifft back to coefficients (fft_masks: 7; rand: 7)
inline Shares libsnark_h(const std::vector<PackedQAPShare>& qap_shares, const std::vector<std::vector<FftMask>>& fft_masks,
                         const Radix2EvaluationDomain& domain, const PackedSharingParams& pp, const LocalTestNet& net,
                         const std::vector<std::vector<Fr>>& rand) {
    if (qap_shares.size() != net.n || fft_masks.size() != 7 || rand.size() != 7) throw Error(ZKG_ERR_BAD_ARG, "libsnark_h: n QAP shares, 7 FFT masks, 7 draw vectors expected");
    const Fr g = Fr::generator(), ginv = g.inverse();                                                // coset_dom.coset_offset(), :29
    Shares in[3];
    for (const auto& q : qap_shares) { in[0].push_back(q.a); in[1].push_back(q.b); in[2].push_back(q.c); }
    Shares ev[3];
    for (int k = 0; k < 3; ++k) {
        Shares coeff = d_ifft(in[k], fft_masks[k], true, domain, g, pp, net, rand[k]);               // :31-64
        ev[k] = d_fft(coeff, fft_masks[3 + k], true, domain, pp, net, rand[3 + k]);                  // :66-75
    }
    const Fr vinv = (g.pow((uint64_t)domain.size()) - Fr::one()).inverse();                          // evaluate_vanishing_polynomial(g)^-1, :78-81
    Shares h(net.n);
    for (uint32_t p = 0; p < net.n; ++p) h[p] = qap_h(ev[0][p], ev[1][p], ev[2][p], &vinv, pp.device);     // :83-88
    return d_ifft(h, fft_masks[6], false, domain, ginv, pp, net, rand[6]);                           // :91-100
}

// dist-primitives/src/dmsm/mod.rs:59-102 for all parties at once.  Every share crosses the "network" as an ark-serialize
// compressed point (mpc-net/src/ser_net.rs:25,40); the king's unpack_missing_shares + sum run on the device.
template <int W>
inline std::vector<Projective<W>> d_msm(const std::vector<std::vector<Affine<W>>>& bases, const Shares& scalars, const std::vector<MsmMask<W>>& masks,
                                        const PackedSharingParams& pp, const LocalTestNet& net) {
    if (bases.size() != net.n || scalars.size() != net.n || masks.size() != net.n) throw Error(ZKG_ERR_BAD_ARG, "d_msm: one input per party expected");
    constexpr size_t WIRE = W == 4 ? 32 : 64;
    auto to_wire = W == 4 ? zkg_g1_to_wire_bn254 : zkg_g2_to_wire_bn254;
    auto from_wire = W == 4 ? zkg_g1_from_wire_bn254 : zkg_g2_from_wire_bn254;
    std::vector<Projective<W>> c_shares;
    for (uint32_t p = 0; p < net.n; ++p) {
        Projective<W> c = msm<W>(bases[p], scalars[p], pp.device);                                  // :73
        c_shares.push_back(add<W>(c, masks[p].in_mask, pp.device));                                 // :74
    }
    const auto parties = net.parties();
    std::vector<uint8_t> frames(parties.size() * WIRE);
    for (size_t r = 0; r < parties.size(); ++r) check(to_wire(pp.device, (const uint64_t*)&c_shares[parties[r]], frames.data() + r * WIRE, 1));   // :79-81
    std::vector<Projective<W>> recv(parties.size());
    check(from_wire(pp.device, frames.data(), (uint64_t*)recv.data(), parties.size()));
    Projective<W> output;
    pp.unpack_missing_shares<W>(recv, parties, &output);                                            // :85-86
    uint8_t frame[WIRE];
    check(to_wire(pp.device, (const uint64_t*)&output, frame, 1));                                  // :90-92
    check(from_wire(pp.device, frame, (uint64_t*)&output, 1));
    std::vector<Projective<W>> out;
    for (uint32_t p = 0; p < net.n; ++p) out.push_back(add<W>(output, masks[p].out_mask, pp.device));   // :98
    return out;
}

}  // namespace zksaas
