// tests/cpp/host_mirror_test.cpp -- the reference's own protocol-level tests, written against the C++ host mirror
// (include/zksaas_host.hpp) the way they are written against the Rust functions, and run on the GPU through the C ABI:
//   dist-primitives/src/dfft/tests.rs        d_ifft_works :20-79, d_fft_works :81-140, d_ifftxd_fft_works :142-220
//   dist-primitives/src/utils/deg_red.rs     :142-191 (squared sharings, L = 4, lossy round)
//   dist-primitives/src/dmsm/mod.rs          :127-180 / examples/dmsm_test.rs:13-93 (d_msm output unpacks to the plain MSM)
//   secret-sharing/src/pss.rs                :250-311 (pack/unpack, det_pack, multiplication)
//   groth16/src/ext_wit.rs                   :287-538 (libsnark_h / circom_h against the plain pipelines), examples/dpp_test.rs
// plus what a compiled host must get right at the boundary: Err(min_len) on a length mismatch, and concurrent calls
// from OS threads (mpc-net/src/multi.rs:320-325 polls one task per party).
// Expected values: the reference's own assertions (plain fft / plain msm through the same library) AND the CPU oracle
// (oracle/libzkoracle.so: zko_fr_fft, zko_g1_msm, zko_g1_fixed_base) -- test infrastructure, linked only here.
//   host_mirror_test --host-only   prints host-side constants (no GPU): checked against pyref by the CPU suite
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "zksaas_host.hpp"

extern "C" {
void zko_fr_fft(uint64_t* v, size_t n, const uint64_t* offset, int inverse);
int zko_g1_msm(const void* bases, size_t stride, const uint64_t* scalars_mont, size_t n, uint64_t* out_xyz, int threads, int c_override);
int zko_g2_msm(const void* bases, size_t stride, const uint64_t* scalars_mont, size_t n, uint64_t* out_xyz, int threads, int c_override);
void zko_g1_fixed_base(const uint64_t* scalars_mont, size_t n, void* out, size_t stride);
void zko_g2_fixed_base(const uint64_t* scalars_mont, size_t n, void* out, size_t stride);
}

using namespace zksaas;

static int g_fail = 0;
#define EXPECT(cond, name)                                                                   \
    do {                                                                                     \
        if (cond) std::printf("PASS  %s\n", name);                                           \
        else { std::printf("FAIL  %s  (%s:%d)\n", name, __FILE__, __LINE__); ++g_fail; }     \
    } while (0)

// counter-based generator (SplitMix64) + the arkworks Fp::rand recipe (256 bits, clear the top two, reject >= r); the
// sampled value is used as the Montgomery image (uniform either way)
struct Rng {
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
    Fr fr() {
        for (;;) {
            Fr a{{next(), next(), next(), next() & ((1ULL << 62) - 1)}};
            if (!Fr::geq_mod(a.v)) return a;
        }
    }
    std::vector<Fr> frs(size_t n) { std::vector<Fr> v(n); for (auto& x : v) x = fr(); return v; }
};

static Shares pack_rearranged(Rng& rng, const PackedSharingParams& pp, std::vector<Fr> x) {     // dfft/tests.rs:29-39 == QAP::pss
    return qap_pss_pack(x, pp, rng.frs(x.size() / pp.l * pp.t));
}
static std::vector<Fr> unpack_all(const PackedSharingParams& pp, const Shares& by_party, bool two = false) {
    std::vector<Fr> cols;
    for (size_t c = 0; c < by_party[0].size(); ++c)
        for (uint32_t p = 0; p < pp.n; ++p) cols.push_back(by_party[p][c]);
    return two ? pp.unpack2(cols) : pp.unpack(cols);
}
static std::vector<FftMask> sample_fft_mask(Rng& rng, bool rearrange, const Fr& g, const Fr& gen, size_t m, const PackedSharingParams& pp) {
    return FftMask::sample(rearrange, g, gen, m, pp, rng.frs(m), rng.frs(m / pp.l * pp.t), rng.frs(m / pp.l * pp.t));
}

static void test_pss(Rng& rng) {                                           // pss.rs:250-311
    for (uint32_t l : {2u, 4u, 8u}) {
        auto pp = PackedSharingParams::new_(l);
        auto secrets = rng.frs(l * 5);
        auto shares = pp.pack(secrets, rng.frs(pp.t * 5));
        EXPECT(pp.unpack(shares) == secrets, "pss: pack -> unpack");
        EXPECT(pp.unpack(pp.det_pack(secrets)) == secrets, "pss: det_pack -> unpack");
        std::vector<Fr> sq(shares.size()), exp(secrets.size());
        for (size_t i = 0; i < shares.size(); ++i) sq[i] = shares[i] * shares[i];           // host Fr arithmetic
        for (size_t i = 0; i < secrets.size(); ++i) exp[i] = secrets[i] * secrets[i];
        EXPECT(pp.unpack2(sq) == exp, "pss: unpack2 of squared shares == squared secrets");
    }
}

static void test_d_ifft_d_fft(Rng& rng, uint32_t l, size_t m) {           // dfft/tests.rs:20-140
    auto pp = PackedSharingParams::new_(l);
    auto dom = Radix2EvaluationDomain::new_(m);
    LocalTestNet net{pp.n, {}};
    auto evals = rng.frs(m);
    auto coeffs = evals;
    dom.ifft_in_place(coeffs);
    {   // the library's plain transform against the oracle's
        auto o = evals;
        zko_fr_fft((uint64_t*)o.data(), m, nullptr, 1);
        EXPECT(o == coeffs, "ifft_in_place == oracle");
    }
    auto masks = sample_fft_mask(rng, false, Fr::one(), dom.group_gen_inv(), m, pp);
    auto out = d_ifft(pack_rearranged(rng, pp, evals), masks, false, dom, Fr::one(), pp, net, rng.frs(m / l * pp.t));
    EXPECT(unpack_all(pp, out) == coeffs, "d_ifft_works");
    auto masks2 = sample_fft_mask(rng, false, Fr::one(), dom.group_gen(), m, pp);
    auto out2 = d_fft(pack_rearranged(rng, pp, coeffs), masks2, false, dom, pp, net, rng.frs(m / l * pp.t));
    EXPECT(unpack_all(pp, out2) == evals, "d_fft_works");
}

static void test_ifft_then_fft(Rng& rng) {                                 // dfft/tests.rs:142-220, with a lossy king round
    const uint32_t l = 2;
    const size_t m = 1 << 9;
    auto pp = PackedSharingParams::new_(l);
    auto dom = Radix2EvaluationDomain::new_(m);
    LocalTestNet net{pp.n, {7}};
    auto evals = rng.frs(m);
    auto m1 = sample_fft_mask(rng, true, Fr::one(), dom.group_gen_inv(), m, pp);
    auto m2 = sample_fft_mask(rng, false, Fr::one(), dom.group_gen(), m, pp);
    auto coeff_sh = d_ifft(pack_rearranged(rng, pp, evals), m1, true, dom, Fr::one(), pp, net, rng.frs(m / l * pp.t));
    auto eval_sh = d_fft(coeff_sh, m2, false, dom, pp, net, rng.frs(m / l * pp.t));
    EXPECT(unpack_all(pp, eval_sh) == evals, "d_ifftxd_fft_works (one party dropped at the king)");
}

static void test_deg_red(Rng& rng) {                                       // utils/deg_red.rs:142-191
    const uint32_t l = 4;
    const size_t num = 64;
    auto pp = PackedSharingParams::new_(l);
    LocalTestNet net{pp.n, {15}};
    auto secrets = rng.frs(num * l);
    auto shares = pack_vec(secrets, pp, rng.frs(num * pp.t));
    Shares sq = shares;
    for (auto& v : sq) for (auto& x : v) x = x * x;
    auto masks = DegRedMask::sample(pp, num, rng.frs(num * l), rng.frs(num * pp.t), rng.frs(num * pp.t));
    auto out = deg_red(sq, masks, pp, net, rng.frs(num * pp.t));
    std::vector<Fr> exp(secrets.size());
    for (size_t i = 0; i < secrets.size(); ++i) exp[i] = secrets[i] * secrets[i];
    EXPECT(unpack_all(pp, out) == exp, "deg_red of squared sharings (lossy round)");
}

template <int W>
static void test_d_msm(Rng& rng, size_t M, std::vector<uint32_t> dropouts, const char* name) {   // dmsm_test.rs:13-93
    const uint32_t l = 2;
    auto pp = PackedSharingParams::new_(l);
    LocalTestNet net{pp.n, dropouts};
    auto fixed = W == 4 ? zko_g1_fixed_base : zko_g2_fixed_base;
    auto y_pub = rng.frs(M), dl = rng.frs(M);
    std::vector<Affine<W>> x_pub(M);
    fixed((const uint64_t*)dl.data(), M, x_pub.data(), sizeof(Affine<W>));
    Projective<W> should_be = msm<W>(x_pub, y_pub);
    Projective<W> oracle;
    (W == 4 ? zko_g1_msm : zko_g2_msm)(x_pub.data(), sizeof(Affine<W>), (const uint64_t*)y_pub.data(), M, (uint64_t*)&oracle, 4, 0);
    EXPECT(should_be == oracle, "G::msm == oracle (arkworks Pippenger restated)");
    // pack the bases through their discrete logs (pack is linear), the scalars directly
    Shares dl_sh = pack_vec(dl, pp, rng.frs(M / l * pp.t));
    std::vector<std::vector<Affine<W>>> x_sh(pp.n, std::vector<Affine<W>>(M / l));
    for (uint32_t p = 0; p < pp.n; ++p) fixed((const uint64_t*)dl_sh[p].data(), M / l, x_sh[p].data(), sizeof(Affine<W>));
    Shares y_sh = pack_vec(y_pub, pp, rng.frs(M / l * pp.t));
    Affine<W> gen_aff;
    Fr one = Fr::one();
    fixed(one.v, 1, &gen_aff, sizeof gen_aff);                               // G::generator()
    auto rnd_pts = [&](size_t k) { std::vector<Affine<W>> a(k); auto s = rng.frs(k); fixed((const uint64_t*)s.data(), k, a.data(), sizeof(Affine<W>));
                                   std::vector<Projective<W>> p; for (auto& x : a) p.push_back(into_group(x)); return p; };
    auto masks = MsmMask<W>::sample(pp, into_group(gen_aff), rng.frs(l), rnd_pts(pp.t), rnd_pts(pp.t));
    auto out = d_msm<W>(x_sh, y_sh, masks, pp, net);
    std::vector<uint32_t> all(pp.n);
    for (uint32_t i = 0; i < pp.n; ++i) all[i] = i;
    auto res = pp.unpack_missing_shares<W>(out, all);                        // every party holds a "repeated" sharing of the output
    bool ok = true;
    for (auto& r : res) ok &= r == should_be;
    EXPECT(ok, name);
}

static void test_d_pp(Rng& rng) {                                          // examples/dpp_test.rs: d_pp(x, x) == all ones; plus a random num / den
    const uint32_t l = 2;
    const size_t m = 1 << 5, cols = m / l;
    auto pp = PackedSharingParams::new_(l);
    LocalTestNet net{pp.n, {}};
    std::vector<Fr> x(m);
    for (size_t i = 0; i < m; ++i) x[i] = Fr::from_u64(i + 1);
    auto px = pack_vec(x, pp, rng.frs(cols * pp.t));
    auto masks = DegRedMask::sample(pp, cols, rng.frs(cols * l), rng.frs(cols * pp.t), rng.frs(cols * pp.t));
    auto out = d_pp(px, px, masks, pp, net, rng.frs(cols * pp.t), rng.frs(cols * pp.t));
    EXPECT(unpack_all(pp, out) == std::vector<Fr>(m, Fr::one()), "d_pp(x, x) == ones (dpp_test.rs)");
    auto num = rng.frs(m), den = rng.frs(m);
    auto out2 = d_pp(pack_vec(num, pp, rng.frs(cols * pp.t)), pack_vec(den, pp, rng.frs(cols * pp.t)), masks, pp, net, rng.frs(cols * pp.t), rng.frs(cols * pp.t));
    std::vector<Fr> exp(m);
    Fr run = Fr::one();
    for (size_t i = 0; i < m; ++i) { run = run * num[i] * den[i].inverse(); exp[i] = run; }    // host arithmetic: prefix products of num/den
    EXPECT(unpack_all(pp, out2) == exp, "d_pp: prefix products of num / den");
}

template <int W>
static void test_crs_det_pack(Rng& rng, const char* name) {                // groth16/src/proving_key.rs:72-104
    const uint32_t l = 2;
    const size_t n = 2 * 9;
    auto pp = PackedSharingParams::new_(l);
    auto fixed = W == 4 ? zko_g1_fixed_base : zko_g2_fixed_base;
    std::vector<Affine<W>> bases(n);
    auto dl = rng.frs(n);
    fixed((const uint64_t*)dl.data(), n, bases.data(), sizeof(Affine<W>));
    std::memset(&bases[5], 0, sizeof bases[5]);
    bases[5].infinity = 1;                                                  // an identity among the secrets
    auto shares = crs_det_pack<W>(bases, l);
    std::vector<uint32_t> all(pp.n);
    for (uint32_t i = 0; i < pp.n; ++i) all[i] = i;
    bool ok = true;
    for (size_t c = 0; c < n / l; ++c) {
        std::vector<Projective<W>> col;
        for (uint32_t p = 0; p < pp.n; ++p) col.push_back(into_group(shares[p][c]));
        auto sec = pp.unpack_missing_shares<W>(col, all);                   // degree < l + t: unpack2 recovers the secrets
        for (uint32_t j = 0; j < l; ++j) ok &= sec[j] == into_group(bases[c * l + j]);
    }
    EXPECT(ok, name);
}

// groth16/src/ext_wit.rs:287-409 libsnark_dummy_ext_witness / :411-538 circom_dummy_ext_witness: a = b = (0..m), c = a*b;
// expected h from the plain pipelines libsnark_ref / circom_ref (:204-285) run through the library's own fft / powers
static void test_ext_witness(Rng& rng, size_t m) {
    const uint32_t l = 2;
    auto pp = PackedSharingParams::new_(l);
    auto dom = Radix2EvaluationDomain::new_(m);
    LocalTestNet net{pp.n, {}};
    const size_t mbyl = m / l;
    std::vector<Fr> a(m), c(m);
    for (size_t i = 0; i < m; ++i) { a[i] = Fr::from_u64(i); c[i] = a[i] * a[i]; }
    // QAP::pss (groth16/src/qap.rs:92-134)
    Shares pa = qap_pss_pack(a, pp, rng.frs(mbyl * pp.t)), pb = qap_pss_pack(a, pp, rng.frs(mbyl * pp.t)), pc = qap_pss_pack(c, pp, rng.frs(mbyl * pp.t));
    std::vector<PackedQAPShare> qap;
    for (uint32_t p = 0; p < pp.n; ++p) qap.push_back(PackedQAPShare{pa[p], pb[p], pc[p]});
    auto draws = [&] { std::vector<std::vector<Fr>> r; for (int i = 0; i < 7; ++i) r.push_back(rng.frs(mbyl * pp.t)); return r; };
    {   // circom
        const Fr root = Radix2EvaluationDomain::new_(2 * m).element(1);
        std::vector<std::vector<FftMask>> fm;
        for (int i = 0; i < 3; ++i) fm.push_back(sample_fft_mask(rng, true, root, dom.group_gen_inv(), m, pp));
        for (int i = 0; i < 3; ++i) fm.push_back(sample_fft_mask(rng, false, Fr::one(), dom.group_gen(), m, pp));
        auto dm = DegRedMask::sample(pp, mbyl, rng.frs(mbyl * l), rng.frs(mbyl * pp.t), rng.frs(mbyl * pp.t));
        auto h = circom_h(qap, fm, dm, dom, pp, net, draws());
        auto coset_eval = [&](std::vector<Fr> v) { dom.ifft_in_place(v); distribute_powers(v, root); dom.fft_in_place(v); return v; };   // circom_ref :239-285
        auto ae = coset_eval(a), ce = coset_eval(c);
        std::vector<Fr> exp(m);
        for (size_t i = 0; i < m; ++i) exp[i] = ae[i] * ae[i] - ce[i];
        EXPECT(unpack_all(pp, h, true) == exp, "circom_h == circom_ref (ext_wit.rs:411-538)");
    }
    {   // libsnark
        const Fr g = Fr::generator(), ginv = g.inverse();
        std::vector<std::vector<FftMask>> fm;
        for (int i = 0; i < 3; ++i) fm.push_back(sample_fft_mask(rng, true, g, dom.group_gen_inv(), m, pp));
        for (int i = 0; i < 3; ++i) fm.push_back(sample_fft_mask(rng, true, Fr::one(), dom.group_gen(), m, pp));
        fm.push_back(sample_fft_mask(rng, false, ginv, dom.group_gen_inv(), m, pp));
        auto h = libsnark_h(qap, fm, dom, pp, net, draws());
        auto coset5 = [&](std::vector<Fr> v) { dom.ifft_in_place(v); dom.fft_in_place(v, &g); return v; };                            // libsnark_ref :204-237
        auto ae = coset5(a), ce = coset5(c);
        const Fr vinv = (g.pow((uint64_t)m) - Fr::one()).inverse();
        std::vector<Fr> exp(m);
        for (size_t i = 0; i < m; ++i) exp[i] = (ae[i] * ae[i] - ce[i]) * vinv;
        dom.ifft_in_place(exp, &g);
        EXPECT(unpack_all(pp, h, true) == exp, "libsnark_h == libsnark_ref (ext_wit.rs:287-409)");
    }
}

static void test_length_mismatch(Rng& rng) {
    std::vector<G1Affine> bases(5);
    auto s = rng.frs(5);
    zko_g1_fixed_base((const uint64_t*)s.data(), 5, bases.data(), sizeof(G1Affine));
    bool caught = false;
    try { msm<4>(bases, rng.frs(3)); } catch (const MsmLengthMismatch& e) { caught = e.min_len == 3; }
    EXPECT(caught, "G::msm -> Err(min(len)) on a length mismatch");
    bool bad = false;
    try { auto pp = PackedSharingParams::new_(3); pp.det_pack(rng.frs(3)); } catch (const Error& e) { bad = e.code == ZKG_ERR_BAD_ARG; }
    EXPECT(bad, "unsupported packing factor is an error, not a crash");
}

static void test_concurrent(Rng& rng) {                                    // multi.rs:320-325: one task per party, different OS threads
    const size_t n = 1 << 12;
    std::vector<G1Affine> bases(n);
    auto dl = rng.frs(n), sc = rng.frs(n);
    zko_g1_fixed_base((const uint64_t*)dl.data(), n, bases.data(), sizeof(G1Affine));
    G1Projective ref = msm<4>(bases, sc);
    auto pp = PackedSharingParams::new_(2);
    auto secrets = rng.frs(2 * 256), rnd = rng.frs(2 * 256);
    auto ref_sh = pp.pack(secrets, rnd);
    std::vector<int> ok(8, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < 8; ++t)
        th.emplace_back([&, t] {
            bool good = true;
            for (int it = 0; it < 4; ++it) { good &= msm<4>(bases, sc) == ref; good &= pp.pack(secrets, rnd) == ref_sh; }
            ok[t] = good;
        });
    for (auto& t : th) t.join();
    bool all = true;
    for (int v : ok) all &= v == 1;
    EXPECT(all, "8 threads issue MSM and pack calls concurrently");
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--host-only") {
        auto d = Radix2EvaluationDomain::new_(1000);
        auto pr = [](const char* k, const Fr& f) { std::printf("%s %016llx%016llx%016llx%016llx\n", k, (unsigned long long)f.v[3], (unsigned long long)f.v[2], (unsigned long long)f.v[1], (unsigned long long)f.v[0]); };
        std::printf("size %zu\n", d.size());
        pr("group_gen", d.group_gen());
        pr("group_gen_inv", d.group_gen_inv());
        pr("size_inv", d.size_inv());
        pr("element5", d.element(5));
        pr("generator", Fr::generator());
        pr("minus_one", -Fr::one());
        pr("gen_times_inv", d.group_gen() * d.group_gen_inv());
        return 0;
    }
    Rng rng{0x7A6B53616153ULL};
    try {
        test_pss(rng);
        test_d_ifft_d_fft(rng, 2, 8);
        test_d_ifft_d_fft(rng, 2, 1 << 10);
        test_d_ifft_d_fft(rng, 4, 64);
        test_ifft_then_fft(rng);
        test_deg_red(rng);
        test_d_pp(rng);
        test_ext_witness(rng, 32);
        test_ext_witness(rng, 1 << 10);
        test_crs_det_pack<4>(rng, "crs det_pack over G1 chunks unpacks to the CRS elements");
        test_crs_det_pack<8>(rng, "crs det_pack over G2 chunks unpacks to the CRS elements");
        test_d_msm<4>(rng, 1 << 10, {}, "d_msm G1 2^10 points, sampled masks, compressed wire == plain MSM");
        test_d_msm<4>(rng, 1 << 8, {3}, "d_msm G1 with a dropped party == plain MSM");
        test_d_msm<8>(rng, 1 << 6, {}, "d_msm G2 == plain MSM");
        test_length_mismatch(rng);
        test_concurrent(rng);
    } catch (const std::exception& e) {
        std::printf("FAIL  exception: %s\n", e.what());
        ++g_fail;
    }
    std::printf(g_fail ? "%d FAILED\n" : "ALL PASS\n", g_fail);
    return g_fail ? 1 : 0;
}
