"""GPU: the reference's own protocol-level tests, re-stated against the CUDA path (BN254 Fr / G1):

  dist-primitives/src/dfft/tests.rs      d_ifft_works, d_fft_works, d_ifftxd_fft_works, coset_d_ifftxd_fft_works
  dist-primitives/src/utils/deg_red.rs   :142-191 (degree reduction of squared sharings, L = 4, with a dropout)
  dist-primitives/examples/dmsm_test.rs  :13-53   (d_msm output unpacks to the plain MSM)
  groth16/src/ext_wit.rs                 :411-538 circom_dummy_ext_witness (expected h = golden circom_ref)
                                         :287-409 libsnark_dummy_ext_witness (expected h = golden libsnark_ref)
plus the committed golden fixtures (tests/golden) through the GPU entry points."""
import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol
from oracle_lib import _p, pyref

pytestmark = pytest.mark.gpu
R = pyref.R_MOD


@pytest.fixture(scope="module")
def z():
    import zksaas_b200
    return zksaas_b200


def strided_secrets(x, l):
    """column i takes x[i + j*m/l], j < l  (dfft/tests.rs:29-39, groth16/src/qap.rs:103-113)"""
    mbyl = x.shape[0] // l
    return np.ascontiguousarray(x.reshape(l, mbyl, 4).transpose(1, 0, 2)).reshape(-1, 4)


def unpack_all(z, pp, shares_by_party, two=False):
    cols = np.ascontiguousarray(np.stack(shares_by_party, axis=1)).reshape(-1, 4)
    return pp.unpack2(cols) if two else pp.unpack(cols)


def sample_fft_mask(z, rng, rearrange, g, gen, m, pp):
    mbyl = m // pp.l
    return z.FftMask.sample(rearrange, g, gen, m, pp, ol.rand_fr(rng, m), ol.rand_fr(rng, mbyl * pp.t),
                            ol.rand_fr(rng, mbyl * pp.t))


def pack_rearranged(z, rng, pp, x):
    xr = z.fft_in_place_rearrange(x.copy())
    return z.transpose(z.pack_vec(strided_secrets(xr, pp.l), pp, ol.rand_fr(rng, xr.shape[0] // pp.l * pp.t)))


@pytest.mark.parametrize("l,m", [(2, 8), (2, 1 << 10), (4, 64)])
def test_d_ifft_works(z, l, m):
    """dfft/tests.rs:20-79"""
    from zksaas_b200 import api
    rng = np.random.default_rng(m + l)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    evals = ol.rand_fr(rng, m)
    coeffs = dom.ifft(evals)
    shares = pack_rearranged(z, rng, pp, evals)
    masks = sample_fft_mask(z, rng, False, api.fr_image(1), dom.group_gen_inv(), m, pp)
    out = z.d_ifft(shares, masks, False, dom, api.fr_image(1), pp, z.LocalTestNet(pp.n), ol.rand_fr(rng, m // l * pp.t))
    assert (unpack_all(z, pp, out) == coeffs).all()
    o = ol.oracle()
    exp = evals.copy()
    o.zko_fr_fft(_p(exp), m, None, 1)
    assert (coeffs == exp).all()


@pytest.mark.parametrize("l,m", [(2, 8), (2, 1 << 10), (4, 64)])
def test_d_fft_works(z, l, m):
    """dfft/tests.rs:81-140"""
    from zksaas_b200 import api
    rng = np.random.default_rng(2 * m + l)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    coeffs = ol.rand_fr(rng, m)
    evals = dom.fft(coeffs)
    shares = pack_rearranged(z, rng, pp, coeffs)
    masks = sample_fft_mask(z, rng, False, api.fr_image(1), dom.group_gen(), m, pp)
    out = z.d_fft(shares, masks, False, dom, pp, z.LocalTestNet(pp.n), ol.rand_fr(rng, m // l * pp.t))
    assert (unpack_all(z, pp, out) == evals).all()


def test_d_ifft_then_d_fft_is_identity(z):
    """dfft/tests.rs:142-220: ifft with rearrange=true feeds the fft directly."""
    from zksaas_b200 import api
    l, m = 2, 1 << 9
    rng = np.random.default_rng(99)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    evals = ol.rand_fr(rng, m)
    shares = pack_rearranged(z, rng, pp, evals)
    one = api.fr_image(1)
    m1 = sample_fft_mask(z, rng, True, one, dom.group_gen_inv(), m, pp)
    m2 = sample_fft_mask(z, rng, False, one, dom.group_gen(), m, pp)
    net = z.LocalTestNet(pp.n)
    coeff_sh = z.d_ifft(shares, m1, True, dom, one, pp, net, ol.rand_fr(rng, m // l * pp.t))
    eval_sh = z.d_fft(coeff_sh, m2, False, dom, pp, net, ol.rand_fr(rng, m // l * pp.t))
    assert (unpack_all(z, pp, eval_sh) == evals).all()


def test_coset_d_ifft_d_fft_chain(z):
    """dfft/tests.rs:222-357: ifft (coset shift g on the way out) -> fft -> ifft (g^-1) -> fft."""
    from zksaas_b200 import api
    l, m = 2, 1 << 8
    rng = np.random.default_rng(17)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    net = z.LocalTestNet(pp.n)
    g = api.fr_image(5)
    ginv = api.fr_image(pow(5, -1, R))
    one = api.fr_image(1)
    evals = ol.rand_fr(rng, m)
    shares = pack_rearranged(z, rng, pp, evals)
    rp = lambda: ol.rand_fr(rng, m // l * pp.t)
    c1 = z.d_ifft(shares, sample_fft_mask(z, rng, True, g, dom.group_gen_inv(), m, pp), True, dom, g, pp, net, rp())
    e1 = z.d_fft(c1, sample_fft_mask(z, rng, True, one, dom.group_gen(), m, pp), True, dom, pp, net, rp())
    # e1 = evaluations over the coset g*H, re-arranged: compare with the plain coset FFT
    coset_evals = dom.fft(dom.ifft(evals), offset=g)
    got = unpack_all(z, pp, e1)                         # strided packing of the bit-reversed vector
    mbyl = m // l
    rev = np.ascontiguousarray(got.reshape(mbyl, l, 4).transpose(1, 0, 2)).reshape(-1, 4)
    assert (z.fft_in_place_rearrange(rev.copy()) == coset_evals).all()
    c2 = z.d_ifft(e1, sample_fft_mask(z, rng, True, ginv, dom.group_gen_inv(), m, pp), True, dom, ginv, pp, net, rp())
    e2 = z.d_fft(c2, sample_fft_mask(z, rng, False, one, dom.group_gen(), m, pp), False, dom, pp, net, rp())
    assert (unpack_all(z, pp, e2) == evals).all()


@pytest.mark.parametrize("dropouts", [(), (15,)])
def test_deg_red_squares(z, dropouts):
    """utils/deg_red.rs:142-191, L = 4 (N = 16); lossy round drops the last party."""
    from zksaas_b200 import api
    l, num = 4, 64
    rng = np.random.default_rng(5)
    pp = z.PackedSharingParams.new(l)
    secrets = ol.rand_fr(rng, num * l)
    shares = z.transpose(z.pack_vec(secrets, pp, ol.rand_fr(rng, num * pp.t)))
    sq = [api.fr_mul(s, s) for s in shares]
    masks = z.DegRedMask.sample(pp, num, ol.rand_fr(rng, num * l), ol.rand_fr(rng, num * pp.t), ol.rand_fr(rng, num * pp.t))
    out = z.deg_red(sq, masks, pp, z.LocalTestNet(pp.n, dropouts), ol.rand_fr(rng, num * pp.t))
    assert (unpack_all(z, pp, out) == api.fr_mul(secrets, secrets)).all()       # degree is back to l+t-1


def test_d_msm_matches_plain_msm(z):
    """dmsm_test.rs:13-93 (BN254 G1, 2^10 public points, l = 2): config 1 of BASELINE.json."""
    o = ol.oracle()
    l, M = 2, 1 << 10
    rng = np.random.default_rng(1010)
    pp = z.PackedSharingParams.new(l)
    y_pub = ol.rand_fr(rng, M)
    dl = ol.rand_fr(rng, M)
    x_pub = np.zeros((M, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(dl), M, x_pub.ctypes.data, 72)
    should_be = z.msm_g1(x_pub, y_pub)
    assert (should_be == ol.o_g1_msm(x_pub, y_pub, threads=8)).all()
    # pack the bases through their discrete logs (pack is linear: share_j of points = (share_j of dlogs) * G)
    rand_x = ol.rand_fr(rng, M // l * pp.t)
    dl_sh = z.transpose(z.pack_vec(dl, pp, rand_x))
    x_shares = []
    for p in range(pp.n):
        aff = np.zeros((M // l, 72), dtype=np.uint8)
        o.zko_g1_fixed_base(_p(dl_sh[p]), M // l, aff.ctypes.data, 72)
        x_shares.append(aff)
    y_shares = z.transpose(z.pack_vec(y_pub, pp, ol.rand_fr(rng, M // l * pp.t)))
    masks = [z.MsmMask.zero() for _ in range(pp.n)]
    out = z.d_msm(x_shares, y_shares, masks, pp, z.LocalTestNet(pp.n))
    # every party ends with the same "repeated" sharing of the output: unpack2 -> result[0] == plain MSM
    from zksaas_b200 import api
    M2 = pp.unpack2_matrix()
    res0 = api.group_lincomb(out, [M2[0][j] for j in range(pp.n)])
    assert (res0 == should_be).all()


def test_circom_h_dataflow_vs_golden(z):
    """groth16/src/ext_wit.rs:104-181 + test :411-538 (a = b = (0..m), c = a*b, m = 2^10): the unpacked h
    must equal circom_ref's h, committed as tests/golden/ext_wit.json."""
    from zksaas_b200 import api
    l, m = 2, 1 << 10
    rng = np.random.default_rng(31)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    net = z.LocalTestNet(pp.n)
    a = ol.fr_np(list(range(m)))
    c = api.fr_mul(a, a)
    root = z.Radix2EvaluationDomain.new(2 * m).element(1)
    one = api.fr_image(1)
    rp = lambda: ol.rand_fr(rng, m // l * pp.t)
    sh = {k: pack_rearranged(z, rng, pp, v) for k, v in (("a", a), ("b", a), ("c", c))}            # QAP::pss
    coeff = {k: z.d_ifft(sh[k], sample_fft_mask(z, rng, True, root, dom.group_gen_inv(), m, pp), True, dom, root, pp, net, rp())
             for k in "abc"}                                                                       # :127-159
    ev = {k: z.d_fft(coeff[k], sample_fft_mask(z, rng, False, one, dom.group_gen(), m, pp), False, dom, pp, net, rp())
          for k in "abc"}                                                                          # :161-170
    h = [api.fr_sub(api.fr_mul(ev["a"][p], ev["b"][p]), ev["c"][p]) for p in range(pp.n)]          # :173-177
    num = m // l
    dmask = z.DegRedMask.sample(pp, num, ol.rand_fr(rng, num * l), rp(), rp())
    h_red = z.deg_red(h, dmask, pp, net, rp())                                                     # :179
    got = unpack_all(z, pp, h_red, two=True)                                                       # test :532-535
    exp = ol.fr_np(gu.ints(gu.load("ext_wit.json")[str(m)]["circom_h"]))
    assert (got == exp).all()


@pytest.mark.parametrize("m", [32, 1 << 10])
def test_libsnark_h_dataflow_vs_golden(z, m):
    """groth16/src/ext_wit.rs:14-102 `libsnark_h` + its test :287-409 (a = b = (0..m), c = a*b; m = 32 there): coset
    d_ifft x3 (rearrange), d_fft x3 (rearrange), h = (ab - c) / Z(g) fused (zkg_qap_h_bn254 with the factor), coset d_ifft
    back to coefficients; the unpack2'ed result must equal libsnark_ref's h (tests/golden/ext_wit.json)."""
    from zksaas_b200 import api
    l = 2
    rng = np.random.default_rng(32 + m)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    net = z.LocalTestNet(pp.n)
    a = ol.fr_np(list(range(m)))
    c = api.fr_mul(a, a)
    one = api.fr_image(1)
    g, ginv = api.fr_image(5), api.fr_image(pow(5, -1, R))                     # coset_dom.coset_offset() / _inv(), :29
    rp = lambda: ol.rand_fr(rng, m // l * pp.t)
    sh = {k: pack_rearranged(z, rng, pp, v) for k, v in (("a", a), ("b", a), ("c", c))}            # QAP::pss
    coeff = {k: z.d_ifft(sh[k], sample_fft_mask(z, rng, True, g, dom.group_gen_inv(), m, pp), True, dom, g, pp, net, rp())
             for k in "abc"}                                                                       # :31-64
    ev = {k: z.d_fft(coeff[k], sample_fft_mask(z, rng, True, one, dom.group_gen(), m, pp), True, dom, pp, net, rp())
          for k in "abc"}                                                                          # :66-75
    vinv = api.fr_image(pow((pow(5, m, R) - 1) % R, -1, R))                                        # :78-81
    h_eval = [api.qap_h(ev["a"][p], ev["b"][p], ev["c"][p], factor=vinv) for p in range(pp.n)]     # :83-88
    h = z.d_ifft(h_eval, sample_fft_mask(z, rng, False, ginv, dom.group_gen_inv(), m, pp), False, dom, ginv, pp, net, rp())  # :91-100
    got = unpack_all(z, pp, h, two=True)                                                           # test :403-406
    exp = ol.fr_np(gu.ints(gu.load("ext_wit.json")[str(m)]["libsnark_h"]))
    assert (got == exp).all()


def test_gpu_vs_golden_fixtures(z):
    from zksaas_b200 import api
    g = gu.load("fields.json")
    for name, p, field in (("fr", pyref.R_MOD, 0), ("fq", pyref.Q_MOD, 1)):
        f = g[name]
        A, B = ol.mont_np(gu.ints(f["a"]), p), ol.mont_np(gu.ints(f["b"]), p)
        for op, key in ((0, "mul"), (1, "add"), (2, "sub")):
            assert ol.np_ints(api._field_op(op, A, B, field), p) == gu.ints(f[key])
    gg = gu.load("groups.json")
    for case in gg["g1_msm"]:
        got = z.msm_g1(gu.g1_from_dlogs(gu.ints(case["dlogs"])), ol.fr_np(gu.ints(case["scalars"])))
        assert ol.g1_xyz_to_point(got) == gu.g1_point(case["result"])
    for case in gg["g2_msm"]:
        got = z.msm_g2(gu.g2_from_dlogs(gu.ints(case["dlogs"])), ol.fr_np(gu.ints(case["scalars"])))
        assert ol.g2_xyz_to_point(got) == gu.g2_point(case["result"])
    gen = np.zeros((1, 72), dtype=np.uint8)
    gen[0] = np.frombuffer(pyref.g1_affine_image(pyref.G1_GEN), dtype=np.uint8)
    for k, exp in gg["g1_multiples"].items():
        assert ol.g1_xyz_to_point(z.msm_g1(gen, ol.fr_np([int(k, 16)]))) == gu.g1_point(exp)
    ps = gu.load("pss.json")
    for ls, f in ps.items():
        pp = z.PackedSharingParams.new(int(ls))
        sh = pp.pack(ol.fr_np(gu.ints(f["secrets"])), ol.fr_np(gu.ints(f["rand"])))
        assert ol.np_fr(sh) == gu.ints(f["shares"])
        assert ol.np_fr(pp.det_pack(ol.fr_np(gu.ints(f["secrets"])))) == gu.ints(f["det_shares"])
        assert ol.np_fr(pp.unpack2(api.fr_mul(sh, sh))) == gu.ints(f["squared_shares_unpack2"])
    df = gu.load("dfft.json")
    for key, f in df.items():
        l, m = int(key.split("_")[0][1:]), int(key.split("_")[1][1:])
        pp = z.PackedSharingParams.new(l)
        dom = z.Radix2EvaluationDomain.new(m)
        assert ol.np_fr(dom.fft(ol.fr_np(list(range(m))))) == gu.ints(f["fft_x"])
        fft1 = []
        for p in range(pp.n):
            v = ol.fr_np(gu.ints(f["party_shares"][p]))
            z.fft1_in_place(v, pp, dom.group_gen())
            assert ol.np_fr(v) == gu.ints(f["fft1"][p])
            fft1.append(v)
        rand = ol.fr_np(sum((gu.ints(r) for r in f["rand_king"]), []))
        zeta = z.Radix2EvaluationDomain.new(2 * m).element(1)
        for rearr in (0, 1):
            for gname, gimg in (("one", api.fr_image(1)), ("zeta_2m", zeta)):
                out = z.king_fft2(fft1, list(range(pp.n)), pp, dom.group_gen(), gimg, rearr, rand)
                exp = f["king"][f"rearrange{rearr}_{gname}"]
                for p in range(pp.n):
                    assert ol.np_fr(out[p]) == gu.ints(exp[p])
        s1 = ol.fr_np(gu.ints(f["fft2_in"]))
        assert ol.np_fr(z.fft2_in_place(s1, pp, dom.group_gen())) == gu.ints(f["fft2_out"])


def test_concurrent_calls_from_threads(z):
    """The C ABI must be re-entrant: LocalTestNet polls up to n party tasks from different OS threads
    (mpc-net/src/multi.rs:320-325).  8 threads issue MSM / fft1 / king / pack calls at once (ctypes
    releases the GIL) and every result must equal the single-threaded one."""
    import threading
    from zksaas_b200 import api
    o = ol.oracle()
    l, mbyl = 2, 1 << 11
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    rng = np.random.default_rng(4242)
    jobs = []
    for t in range(8):
        n = 700 + 37 * t
        dl = ol.rand_fr(rng, n)
        bases = np.zeros((n, 72), dtype=np.uint8)
        o.zko_g1_fixed_base(_p(dl), n, bases.ctypes.data, 72)
        jobs.append({"bases": bases, "scalars": ol.rand_fr(rng, n), "px": ol.rand_fr(rng, mbyl),
                     "shares": [ol.rand_fr(rng, mbyl) for _ in range(pp.n)], "rand": ol.rand_fr(rng, mbyl * pp.t),
                     "sec": ol.rand_fr(rng, 64 * l)})

    def work(j):
        return (z.msm_g1(j["bases"], j["scalars"]),
                z.fft1_in_place(j["px"].copy(), pp, dom.group_gen()),
                z.king_fft2(j["shares"], list(range(pp.n)), pp, dom.group_gen(), api.fr_image(5), True, j["rand"]),
                pp.pack(j["sec"], j["rand"][:64 * pp.t]))

    expect = [work(j) for j in jobs]
    got = [None] * len(jobs)
    errs = []

    def run(i):
        try:
            for _ in range(3):
                got[i] = work(jobs[i])
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for e, g in zip(expect, got):
        assert (e[0] == g[0]).all() and (e[1] == g[1]).all() and (e[3] == g[3]).all()
        assert all((a == b).all() for a, b in zip(e[2], g[2]))


@pytest.mark.parametrize("l,cols", [(2, 16), (2, 5000), (4, 333)])
def test_dpp_king_vs_oracle(z, l, cols):
    """king closure of d_pp (dpp/mod.rs:41-76) vs the literal oracle loop."""
    import ctypes as C
    o = ol.oracle()
    rng = np.random.default_rng(cols + l)
    pp = z.PackedSharingParams.new(l)
    num, den = ol.rand_fr(rng, cols * l), ol.rand_fr(rng, cols * l)
    ns = z.transpose(z.pack_vec(num, pp, ol.rand_fr(rng, cols * pp.t)))
    ds = z.transpose(z.pack_vec(den, pp, ol.rand_fr(rng, cols * pp.t)))
    shares = [np.concatenate([ns[p], ds[p]]) for p in range(pp.n)]
    rand = ol.rand_fr(rng, cols * pp.t)
    exp = [np.zeros((cols, 4), dtype=np.uint64) for _ in range(pp.n)]
    par = (C.c_uint32 * pp.n)(*range(pp.n))
    assert o.zko_dpp_king(ol.ptr_array(shares), par, pp.n, cols, l, _p(rand), ol.ptr_array(exp)) == 0
    got = z.dpp_king(shares, list(range(pp.n)), pp, rand)
    for p in range(pp.n):
        assert (got[p] == exp[p]).all()
    # a zero denominator is an error, as in the reference (inverse().unwrap())
    den[7] = 0
    ds = z.transpose(z.pack_vec(den, pp, ol.rand_fr(rng, cols * pp.t)))
    shares = [np.concatenate([ns[p], ds[p]]) for p in range(pp.n)]
    with pytest.raises(z.ZkgError):
        z.dpp_king(shares, list(range(pp.n)), pp, rand)


def test_d_pp_example(z):
    """dist-primitives/examples/dpp_test.rs: num = den = (1..m) => every partial product is one."""
    from zksaas_b200 import api
    l, m = 2, 1 << 5
    rng = np.random.default_rng(55)
    pp = z.PackedSharingParams.new(l)
    x = ol.fr_np(list(range(1, m + 1)))
    px = z.transpose(z.pack_vec(x, pp, ol.rand_fr(rng, m // l * pp.t)))
    num = m // l
    masks = z.DegRedMask.sample(pp, num, ol.rand_fr(rng, num * l), ol.rand_fr(rng, num * pp.t), ol.rand_fr(rng, num * pp.t))
    out = z.d_pp(px, px, masks, pp, z.LocalTestNet(pp.n), ol.rand_fr(rng, num * pp.t), ol.rand_fr(rng, num * pp.t))
    got = unpack_all(z, pp, out)
    assert (got == ol.fr_np([1] * m)).all()


@pytest.mark.parametrize("l,g2", [(2, False), (4, False), (2, True)])
def test_crs_det_pack_group_vs_oracle(z, l, g2):
    """det_pack over group elements per l-chunk (groth16/src/proving_key.rs:72-104) vs the oracle's
    literal FFT-over-points det_pack (secret-sharing/src/pss.rs:69-87), incl. an infinity element."""
    from zksaas_b200 import api
    o = ol.oracle()
    rng = np.random.default_rng(l * 10 + g2)
    pp = z.PackedSharingParams.new(l)
    chunks = 6
    n = chunks * l
    stride, words = (136, 24) if g2 else (72, 12)
    bases = np.zeros((n, stride), dtype=np.uint8)
    (o.zko_g2_fixed_base if g2 else o.zko_g1_fixed_base)(_p(ol.rand_fr(rng, n)), n, bases.ctypes.data, stride)
    bases[3] = 0
    bases[3, stride - 8] = 1                                   # infinity
    got = api.crs_det_pack(bases, pp, g2)
    half = (stride - 8) // 2
    for j in range(chunks):
        sec = np.zeros((l, words), dtype=np.uint64)
        for k in range(l):
            img = bases[j * l + k]
            if img[stride - 8]:
                pt = None
            elif g2:
                f = lambda b: int.from_bytes(bytes(b), "little") * pow(1 << 256, -1, pyref.Q_MOD) % pyref.Q_MOD
                pt = (pyref.Fq2(f(img[0:32]), f(img[32:64])), pyref.Fq2(f(img[64:96]), f(img[96:128])))
            else:
                f = lambda b: int.from_bytes(bytes(b), "little") * pow(1 << 256, -1, pyref.Q_MOD) % pyref.Q_MOD
                pt = (f(img[0:32]), f(img[32:64]))
            sec[k] = ol.g2_point_to_xyz(pt) if g2 else ol.g1_point_to_xyz(pt)
        exp = np.zeros(pp.n * words, dtype=np.uint64)
        (o.zko_pss_pack_g2 if g2 else o.zko_pss_pack_g1)(l, _p(sec.reshape(-1)), None, _p(exp))
        for i in range(pp.n):
            e = exp[i * words:(i + 1) * words]
            want = api._xyz_to_affine_images([e], g2)[0]
            assert (got[i][j] == want).all(), (j, i)


# ------------------------------------------------------------------------------------------------
# offline packing (pack_from_witness, QAP::pss) and MsmMask::sample
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("l,length", [(2, 13), (2, 1), (4, 1000), (8, 64), (2, 0)])
def test_pack_from_witness_vs_oracle(z, l, length):
    """groth16/examples/sha256.rs:131-156: ragged last chunk is zero-padded."""
    import random
    rng = random.Random(l * 1000 + length)
    pp = z.PackedSharingParams.new(l)
    ref = pyref.PackedSharingParams(l)
    w = [rng.randrange(R) for _ in range(length)]
    cols = (length + l - 1) // l
    rand = [[rng.randrange(R) for _ in range(ref.t)] for _ in range(cols)]
    got = z.pack_from_witness(pp, ol.fr_np(w).reshape(-1, 4), ol.fr_np([x for r in rand for x in r]).reshape(-1, 4))
    padded = w + [0] * (cols * l - length)
    exp = pyref.transpose([ref.pack(padded[c * l:(c + 1) * l], rand[c]) for c in range(cols)]) if cols else [[] for _ in range(ref.n)]
    assert len(got) == ref.n
    for p in range(ref.n):
        assert ol.np_fr(got[p]) == exp[p]


@pytest.mark.parametrize("l,m", [(2, 32), (4, 256), (2, 1 << 12)])
def test_qap_pss_pack_vs_oracle(z, l, m):
    """groth16/src/qap.rs:99-133 with a = (0..m) as in ext_wit.rs:415-421."""
    import random
    rng = random.Random(m + l)
    pp = z.PackedSharingParams.new(l)
    ref = pyref.PackedSharingParams(l)
    a = [(i * i + 7) % R for i in range(m)]
    mbyl = m // l
    rand = [[rng.randrange(R) for _ in range(ref.t)] for _ in range(mbyl)]
    got = z.qap_pss_pack(ol.fr_np(a), pp, ol.fr_np([x for r in rand for x in r]))
    xr = pyref.fft_in_place_rearrange(list(a))
    exp = pyref.transpose([ref.pack([xr[i + j * mbyl] for j in range(l)], rand[i]) for i in range(mbyl)])
    for p in range(ref.n):
        assert ol.np_fr(got[p]) == exp[p]
    # and through the three-vector wrapper
    tri = z.qap_pss(ol.fr_np(a), ol.fr_np(a), ol.fr_np(a), pp, *([ol.fr_np([x for r in rand for x in r])] * 3))
    assert all((tri[p][k] == got[p]).all() for p in range(ref.n) for k in range(3))


@pytest.mark.parametrize("g2", [False, True])
def test_msm_mask_sample_vs_oracle(z, g2):
    """dmsm/mod.rs:21-48: in-mask shares pack gen*x_i, out-mask shares pack -(sum) repeated l times;
    unpacking the shares (over the group) returns the mask values, which cancel."""
    import random
    rng = random.Random(99 + g2)
    l = 2
    pp = z.PackedSharingParams.new(l)
    ref = pyref.PackedSharingParams(l)
    curve, gen = (pyref.G2, pyref.G2_GEN_PT) if g2 else (pyref.G1, pyref.G1_GEN)
    to_xyz = ol.g2_point_to_xyz if g2 else ol.g1_point_to_xyz
    to_pt = ol.g2_xyz_to_point if g2 else ol.g1_xyz_to_point
    xs = [rng.randrange(R) for _ in range(l)]
    rin = [curve.mul(gen, rng.randrange(R)) for _ in range(ref.t)]
    rout = [curve.mul(gen, rng.randrange(R)) for _ in range(ref.t)]
    masks = z.MsmMask.sample(pp, ol.fr_np(xs), [to_xyz(p) for p in rin], [to_xyz(p) for p in rout], g2=g2)
    assert len(masks) == ref.n
    ops = pyref.group_ops(curve)
    values = [curve.mul(gen, x) for x in xs]
    total = None
    for v in values:
        total = curve.add(total, v)
    out_value = curve.neg(total)
    exp_in = ref.pack(values, rin, ops)
    exp_out = ref.pack([out_value] * l, rout, ops)
    assert [to_pt(mk.in_mask) for mk in masks] == exp_in
    assert [to_pt(mk.out_mask) for mk in masks] == exp_out
    assert z.group_generator(g2).tolist() == to_xyz(gen).tolist()
