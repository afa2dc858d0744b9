"""GPU parity for the rows added in round 2, all through the C ABI:
  * king side of d_msm over group elements incl. the dropout (Lagrange) path   dmsm/mod.rs:85-87, pss.rs:141-221
  * compressed G1 / G2 wire format                                              ser_net.rs:25,40,119
  * fused h = (a+ma)(b+mb) - (c+mc) [* 1/Z]                                     ext_wit.rs:82-86,173-177
  * FftMask::sample / DegRedMask::sample device-resident                        dfft/mod.rs:30-85, deg_red.rs:40-66
  * the opt-in kernels (batched-affine accumulation, single-thread Horner tail) on every special case
  * registered-bases handle lifetime (release racing an MSM, stale handles, slot reuse)
  * the sizes bench.py reports that had no parity test: G2 at 2^19, d_fft at m = 2^24
"""
import ctypes as C
import random
import threading

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import _p, pyref

pytestmark = pytest.mark.gpu
R, Q = pyref.R_MOD, pyref.Q_MOD


@pytest.fixture(scope="module")
def z():
    import zksaas_b200
    return zksaas_b200


def _curve(g2):
    if g2:
        return pyref.G2, pyref.G2_GEN_PT, ol.g2_point_to_xyz, ol.g2_xyz_to_point
    return pyref.G1, pyref.G1_GEN, ol.g1_point_to_xyz, ol.g1_xyz_to_point


def _scale_jacobian(xyz, lam, g2):
    """(X, Y, Z) -> (lam^2 X, lam^3 Y, lam Z): another representative of the same point (lam in Fq)."""
    w = 8 if g2 else 4
    a = np.asarray(xyz, dtype=np.uint64).reshape(3, w // 4, 4)
    out = np.zeros_like(a)
    for comp, e in ((0, 2), (1, 3), (2, 1)):
        for k in range(w // 4):
            v = pyref.from_mont_limbs(a[comp, k], Q) * pow(lam, e, Q) % Q
            out[comp, k] = pyref.to_mont_limbs(v, Q)
    return out.reshape(-1)


# ---------------------------------------------------------------------------------------------------
# king side of d_msm
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("g2,l,dropouts", [(False, 2, ()), (False, 2, (7,)), (False, 2, (3,)), (True, 2, ()), (True, 2, (0,)),
                                           (False, 4, (15,)), (True, 4, ())])       # l = 4: 16 shares, slow in the big-int model
def test_pss_unpack2_group_vs_oracle(z, g2, l, dropouts):
    """unpack_missing_shares over points == the packed secrets; the sum is what d_msm's king replicates."""
    rng = random.Random(100 * l + len(dropouts) + g2)
    curve, gen, to_xyz, to_pt = _curve(g2)
    pp, ref = z.PackedSharingParams.new(l), pyref.PackedSharingParams(l)
    ops = pyref.group_ops(curve)
    secrets = [curve.mul(gen, rng.randrange(R)) for _ in range(l)]
    if l == 4:
        secrets[1] = None                                             # an identity among the secrets
    # unpack2 must also hold for degree-2(l+t)-2 sharings; a fresh packing has degree l+t-1
    shares = ref.pack(secrets, [curve.mul(gen, rng.randrange(R)) for _ in range(ref.t)], ops)
    parties = [p for p in range(ref.n) if p not in dropouts]
    imgs = [to_xyz(shares[p]) for p in parties]
    imgs[0] = _scale_jacobian(imgs[0], rng.randrange(2, Q), g2)        # a non-normalised Projective input
    rows, total = z.pss_unpack2_group(pp, imgs, parties, g2)
    assert [to_pt(r) for r in rows] == secrets
    exp = None
    for s_ in secrets:
        exp = curve.add(exp, s_)
    assert to_pt(total) == exp
    _, total2 = z.pss_unpack2_group(pp, imgs, parties, g2, want_unpacked=False)   # the king's form: column sums, one row
    assert (total2 == total).all()
    # the oracle's own dropout path agrees (pss.rs:210-221)
    got_ref = ref.unpack_missing_shares([shares[p] for p in parties], parties, ops)
    assert got_ref == secrets


def test_pss_unpack2_group_too_few_shares(z):
    from zksaas_b200 import capi
    pp = z.PackedSharingParams.new(2)
    pts = [ol.g1_point_to_xyz(pyref.G1_GEN)] * 6
    with pytest.raises(capi.ZkgError) as e:
        z.pss_unpack2_group(pp, pts, [0, 1, 2, 3, 4, 5])
    assert e.value.code == capi.ZKG_ERR_BAD_ARG


@pytest.mark.parametrize("dropouts", [(), (7,), (2,)])
def test_d_msm_with_dropouts(z, dropouts):
    """dmsm/mod.rs:59-102 under simulate_lossy_network_round (mpc-net/src/multi.rs:330-363): the king reconstructs
    from the parties that answered; every party still ends with the sharing of the plain MSM."""
    o = ol.oracle()
    l, M = 2, 1 << 8
    rng = np.random.default_rng(77)
    pp = z.PackedSharingParams.new(l)
    y_pub, dl = ol.rand_fr(rng, M), ol.rand_fr(rng, M)
    x_pub = np.zeros((M, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(dl), M, x_pub.ctypes.data, 72)
    should_be = ol.o_g1_msm(x_pub, y_pub, threads=4)
    dl_sh = z.transpose(z.pack_vec(dl, pp, ol.rand_fr(rng, M // l * pp.t)))
    x_shares = []
    for p in range(pp.n):
        aff = np.zeros((M // l, 72), dtype=np.uint8)
        o.zko_g1_fixed_base(_p(dl_sh[p]), M // l, aff.ctypes.data, 72)
        x_shares.append(aff)
    y_shares = z.transpose(z.pack_vec(y_pub, pp, ol.rand_fr(rng, M // l * pp.t)))
    masks = [z.MsmMask.zero() for _ in range(pp.n)]
    out = z.d_msm(x_shares, y_shares, masks, pp, z.LocalTestNet(pp.n, dropouts=dropouts))       # shares cross as compressed points
    assert all((o_ == should_be).all() for o_ in out)       # the king replicates the clear output (dmsm/mod.rs:87)
    out2 = z.d_msm(x_shares, y_shares, masks, pp, z.LocalTestNet(pp.n, dropouts=dropouts), wire=False)
    assert all((o_ == should_be).all() for o_ in out2)


# ---------------------------------------------------------------------------------------------------
# compressed wire format
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("g2", [False, True])
def test_group_wire_roundtrip_vs_oracle(z, g2):
    rng = random.Random(31 + g2)
    curve, gen, to_xyz, to_pt = _curve(g2)
    ser = pyref.g2_serialize_compressed if g2 else pyref.g1_serialize_compressed
    de = pyref.g2_deserialize_compressed if g2 else pyref.g1_deserialize_compressed
    pts = [curve.mul(gen, rng.randrange(R)) for _ in range(24)] + [None]
    pts += [curve.neg(p) for p in pts[:8]]                                   # both signs of the same abscissa
    imgs = [to_xyz(p) for p in pts]
    imgs[3] = _scale_jacobian(imgs[3], rng.randrange(2, Q), g2)              # to_wire takes any Projective representative
    wire = z.group_to_wire(np.stack(imgs), g2)
    exp = np.stack([np.frombuffer(ser(p), dtype=np.uint8) for p in pts])
    assert (wire == exp).all()
    assert {bool(w[-1] & 0x80) for w in exp[:8]} | {bool(w[-1] & 0x80) for w in exp[25:]} == {True, False}
    back = z.group_from_wire(wire, g2)
    assert [to_pt(b) for b in back] == pts == [de(bytes(w)) for w in exp]


@pytest.mark.parametrize("g2", [False, True])
def test_group_wire_rejects_invalid_encodings(z, g2):
    from zksaas_b200 import capi
    nb = 64 if g2 else 32
    curve, gen, to_xyz, _ = _curve(g2)
    ser = pyref.g2_serialize_compressed if g2 else pyref.g1_serialize_compressed
    de = pyref.g2_deserialize_compressed if g2 else pyref.g1_deserialize_compressed
    good = bytearray(ser(curve.mul(gen, 12345)))
    bad = []
    b = bytearray(good); b[-1] |= 0xC0; bad.append(bytes(b))                 # both flags
    b = bytearray(Q.to_bytes(32, "little")) if not g2 else bytearray((5).to_bytes(32, "little") + Q.to_bytes(32, "little"))
    bad.append(bytes(b))                                                     # x (or x.c1) == q: not below the modulus
    x = 1
    while True:                                                              # an abscissa with no curve point
        cand = bytearray(x.to_bytes(32, "little") + (bytes(32) if g2 else b""))
        try:
            de(bytes(cand)); x += 1
        except pyref.WireError as e:
            if "curve" not in str(e):                                        # on the twist but outside the subgroup: next x
                x += 1
                continue
            bad.append(bytes(cand)); break
    if g2:                                                                   # on the twist, outside the r-torsion subgroup
        xx = pyref.Fq2(1, 0)
        while True:
            y = pyref._fq2_sqrt(xx * xx * xx + pyref.G2_B)
            if y is not None and pyref.G2.add(pyref.G2.mul_raw((xx, y), R - 1), (xx, y)) is not None:
                bad.append(pyref.g2_serialize_compressed((xx, y))); break
            xx = xx + pyref.Fq2(1, 0)
    for enc in bad:
        with pytest.raises(pyref.WireError):
            de(enc)
        with pytest.raises(capi.ZkgError) as e:
            z.group_from_wire(np.frombuffer(enc, dtype=np.uint8).reshape(1, nb), g2)
        assert e.value.code == capi.ZKG_ERR_BAD_ARG
    ok = z.group_from_wire(np.frombuffer(bytes(good), dtype=np.uint8).reshape(1, nb), g2)
    assert (ok[0] == to_xyz(curve.mul(gen, 12345))).all()


# ---------------------------------------------------------------------------------------------------
# fused h
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 255, 4096])
def test_qap_h_fused(z, n):
    rng = np.random.default_rng(n)
    a, b, c, ma, mb, mc = (ol.rand_fr(rng, n) for _ in range(6))
    f = ol.rand_fr(rng, 1)[0]
    ai, bi, ci, mai, mbi, mci = (ol.np_fr(v) for v in (a, b, c, ma, mb, mc))
    fi = ol.np_fr(f.reshape(1, 4))[0]
    exp = [((x + mx) * (y + my) - (w + mw)) % R for x, y, w, mx, my, mw in zip(ai, bi, ci, mai, mbi, mci)]
    assert (z.qap_h(a, b, c, ma, mb, mc) == ol.fr_np(exp)).all()
    assert (z.qap_h(a, b, c, ma, mb, mc, factor=f) == ol.fr_np([e * fi % R for e in exp])).all()
    assert (z.qap_h(a, b, c) == ol.fr_np([(x * y - w) % R for x, y, w in zip(ai, bi, ci)])).all()
    # edge: a*b == c gives exactly zero; c == 0
    zero = np.zeros_like(a)
    from zksaas_b200 import api
    assert not z.qap_h(a, b, api.fr_mul(a, b)).any()
    assert (z.qap_h(a, b, zero) == api.fr_mul(a, b)).all()
    with pytest.raises(ValueError):
        z.qap_h(a, b[:-1] if n > 1 else np.zeros((2, 4), dtype=np.uint64), c)


# ---------------------------------------------------------------------------------------------------
# offline masks
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("l,m", [(2, 8), (2, 256), (4, 64), (8, 64)])
@pytest.mark.parametrize("rearrange", [False, True])
@pytest.mark.parametrize("coset", [False, True])
def test_fft_mask_sample_vs_literal_steps(z, l, m, rearrange, coset):
    """dfft/mod.rs:30-85 step by step in the big-integer model, same draws."""
    rng = np.random.default_rng(m + l + rearrange)
    pp, ref = z.PackedSharingParams.new(l), pyref.PackedSharingParams(l)
    mbyl = m // l
    dom = pyref.Radix2Domain(m)
    gen = dom.group_gen
    g = pyref.Radix2Domain(2 * m).element(1) if coset else 1
    mask, rin, rout = ol.rand_fr(rng, m), ol.rand_fr(rng, mbyl * ref.t), ol.rand_fr(rng, mbyl * ref.t)
    got = z.FftMask.sample(rearrange, ol.fr_np([g])[0], ol.fr_np([gen])[0], m, pp, mask, rin, rout)
    mv, ri, ro = ol.np_fr(mask), ol.np_fr(rin), ol.np_fr(rout)
    chunks = lambda r_: [r_[i * ref.t:(i + 1) * ref.t] for i in range(mbyl)]
    in_sh = pyref.transpose(pyref.pack_vec(list(mv), ref, chunks(ri)))               # :41-42
    s = pyref.fft2_in_place(list(mv), ref, gen)                                      # :44
    if g != 1:
        s = pyref.distribute_powers(s, g)                                            # :46-48
    s = [(-v) % R for v in s]                                                        # :51
    if rearrange:                                                                    # :55-72
        s = pyref.fft_in_place_rearrange(s)
        cols = [ref.pack([s[i + j * mbyl] for j in range(l)], ro[i * ref.t:(i + 1) * ref.t]) for i in range(mbyl)]
        out_sh = pyref.transpose(cols)
    else:
        out_sh = pyref.transpose(pyref.pack_vec(s, ref, chunks(ro)))
    for p in range(ref.n):
        assert ol.np_fr(got[p].in_mask) == in_sh[p]
        assert ol.np_fr(got[p].out_mask) == out_sh[p]


def test_deg_red_mask_sample_vs_literal_steps(z):
    l, num = 2, 33
    rng = np.random.default_rng(5)
    pp, ref = z.PackedSharingParams.new(l), pyref.PackedSharingParams(l)
    mask, rin, rout = ol.rand_fr(rng, num * l), ol.rand_fr(rng, num * ref.t), ol.rand_fr(rng, num * ref.t)
    got = z.DegRedMask.sample(pp, num, mask, rin, rout)
    mv = ol.np_fr(mask)
    chunks = lambda r_: [r_[i * ref.t:(i + 1) * ref.t] for i in range(num)]
    ins = pyref.transpose(pyref.pack_vec(list(mv), ref, chunks(ol.np_fr(rin))))
    outs = pyref.transpose(pyref.pack_vec([(-v) % R for v in mv], ref, chunks(ol.np_fr(rout))))
    for p in range(ref.n):
        assert ol.np_fr(got[p].in_mask) == ins[p] and ol.np_fr(got[p].out_mask) == outs[p]
    with pytest.raises(ValueError):
        z.DegRedMask.sample(pp, num, mask[:-1], rin, rout)


# ---------------------------------------------------------------------------------------------------
# opt-in kernels on every special case (k_accumulate_ba / single-thread tail had no committed test)
# ---------------------------------------------------------------------------------------------------
def _special_case_inputs(o, g2, n, seed):
    """points with repeats, negations, infinities and a few heavy buckets; returns (affine images, scalars)."""
    rng = np.random.default_rng(seed)
    stride = 136 if g2 else 72
    dl = ol.rand_fr(rng, n)
    dl[1::7] = dl[0]                                       # many copies of one point: P + P inside a bucket
    k = min(len(dl[2::11]), len(dl[3::11]))
    dl[3::11][:k] = ol.fr_np([(R - v) % R for v in ol.np_fr(np.ascontiguousarray(dl[2::11][:k]))])     # P and -P pairs
    bases = np.zeros((n, stride), dtype=np.uint8)
    (o.zko_g2_fixed_base if g2 else o.zko_g1_fixed_base)(_p(np.ascontiguousarray(dl)), n, bases.ctypes.data, stride)
    bases[5::13] = 0
    bases[5::13, stride - 8] = 1                           # infinity flag
    sc = ol.rand_fr(rng, n)
    sc[0::3] = sc[0]                                       # equal scalars: equal digits, long buckets
    sc[4::9] = ol.fr_np([1])[0]
    sc[8::17] = 0
    return bases, sc


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("envs", ["ZKG_MSM_BA=1", "ZKG_MSM_COOP_TAIL=0", "ZKG_MSM_G2_PAIR=1", "ZKG_MSM_G2_PAIR=0",
                                  # round 2, second session: sort pipeline (window groups on the side stream / on one stream, three
                                  # groups), heavy-bucket threshold (nearly every bucket heavy / the default rule), both reduction
                                  # shapes with and without the four-warp levels
                                  "ZKG_MSM_GROUP0=2", "ZKG_MSM_GROUP0=2;ZKG_MSM_SIDE=0", "ZKG_MSM_GROUP0=1;ZKG_MSM_GROUPS=3",
                                  "ZKG_MSM_HEAVY_KEY=2", "ZKG_MSM_HEAVY_KEY=24",
                                  "ZKG_MSM_REDUCE_L=1", "ZKG_MSM_REDUCE_L=8", "ZKG_MSM_COOP_REDUCE=0;ZKG_MSM_REDUCE_L=1",
                                  "ZKG_MSM_COOP_REDUCE=0;ZKG_MSM_REDUCE_L=8"])
@pytest.mark.parametrize("n", [700, 1 << 13])
def test_optin_msm_kernels_special_cases(z, monkeypatch, g2, envs, n):
    o = ol.oracle()
    bases, sc = _special_case_inputs(o, g2, n, 17 + g2)
    msm, omsm = (z.msm_g2, ol.o_g2_msm) if g2 else (z.msm_g1, ol.o_g1_msm)
    exp = omsm(bases, sc, threads=8)
    assert (msm(bases, sc) == exp).all()
    for kv in envs.split(";"):
        monkeypatch.setenv(*kv.split("="))
    monkeypatch.setenv("ZKG_MSM_CHUNKS", "3")              # the chunked accumulate_into path as well
    assert (msm(bases, sc) == exp).all()
    monkeypatch.setenv("ZKG_MSM_C", "6")                   # long buckets: the batched-affine tree runs several rounds
    assert (msm(bases, sc) == exp).all()
    if n <= 1024:                                          # registered (prepared-table) path with the same switch
        from zksaas_b200 import capi
        h = C.c_uint64(0)
        capi.check(z.lib().zkg_bases_register(0, 2 if g2 else 1, bases.ctypes.data, bases.shape[1], n, C.byref(h)))
        out = np.zeros(24 if g2 else 12, dtype=np.uint64)
        capi.check(z.lib().zkg_msm_bn254_registered(h.value, _p(sc), n, _p(out)))
        capi.check(z.lib().zkg_bases_release(h.value))
        assert (out == exp).all()


# ---------------------------------------------------------------------------------------------------
# registered-bases handles
# ---------------------------------------------------------------------------------------------------
def test_bases_handle_lifetime(z):
    from zksaas_b200 import capi
    lib = z.lib()
    o = ol.oracle()
    rng = np.random.default_rng(8)
    n = 1 << 14
    bases = np.zeros((n, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(ol.rand_fr(rng, n)), n, bases.ctypes.data, 72)
    sc = ol.rand_fr(rng, n)
    exp = ol.o_g1_msm(bases, sc, threads=8)
    h = C.c_uint64(0)
    capi.check(lib.zkg_bases_register(0, 1, bases.ctypes.data, 72, n, C.byref(h)))
    # release racing MSMs on other threads: every MSM either completes with the right point or is refused with
    # BAD_ARG (handle already released) -- never a wrong point, never a fault
    results, errs = [], []

    def worker():
        for _ in range(6):
            out = np.zeros(12, dtype=np.uint64)
            rc = lib.zkg_msm_bn254_registered(h.value, _p(sc), n, _p(out))
            (results if rc == 0 else errs).append((rc, out))
    ths = [threading.Thread(target=worker) for _ in range(3)]
    for t in ths:
        t.start()
    rc_rel = lib.zkg_bases_release(h.value)
    for t in ths:
        t.join()
    assert rc_rel == 0
    assert all((out == exp).all() for _, out in results)
    assert all(rc == capi.ZKG_ERR_BAD_ARG for rc, _ in errs)
    assert lib.zkg_bases_release(h.value) == capi.ZKG_ERR_BAD_ARG                  # double release
    # the slot is reused under a new generation: the stale handle stays invalid
    h2 = C.c_uint64(0)
    capi.check(lib.zkg_bases_register(0, 1, bases.ctypes.data, 72, n, C.byref(h2)))
    assert h2.value != h.value and (h2.value & 0xffffffff) == (h.value & 0xffffffff)
    out = np.zeros(12, dtype=np.uint64)
    assert lib.zkg_msm_bn254_registered(h.value, _p(sc), n, _p(out)) == capi.ZKG_ERR_BAD_ARG
    capi.check(lib.zkg_msm_bn254_registered(h2.value, _p(sc), n, _p(out)))
    assert (out == exp).all()
    capi.check(lib.zkg_bases_release(h2.value))


def test_python_mirror_validates_lengths(z):
    pp = z.PackedSharingParams.new(2)
    rng = np.random.default_rng(1)
    with pytest.raises(ValueError):
        pp.pack(ol.rand_fr(rng, 3), ol.rand_fr(rng, 2))
    with pytest.raises(ValueError):
        pp.unpack2(ol.rand_fr(rng, 9))
    shares = [ol.rand_fr(rng, 8) for _ in range(8)]
    gen = z.Radix2EvaluationDomain.new(16).group_gen()
    with pytest.raises(ValueError):
        z.king_fft2(shares[:7] + [ol.rand_fr(rng, 7)], list(range(8)), pp, gen, gen, False, ol.rand_fr(rng, 16))
    with pytest.raises(ValueError):
        z.king_fft2(shares, list(range(8)), pp, gen, gen, False, ol.rand_fr(rng, 15))
    with pytest.raises(ValueError):
        z.fft1_in_place(ol.rand_fr(rng, 8), pp, gen, in_mask=ol.rand_fr(rng, 7))


# ---------------------------------------------------------------------------------------------------
# the peer-store exchange kernels, with all "ranks" on one device (the same launches the multi-GPU paths make;
# tests/test_gpu_multi.py runs them across real GPUs)
# ---------------------------------------------------------------------------------------------------
def _dev_ctx(z):
    from zksaas_b200 import capi
    ctx = capi.ctx_p()
    capi.check(z.lib().zkg_ctx_create(0, C.c_void_p(1), C.byref(ctx)))
    return ctx


def _rand_dev(torch, k, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 61) - 1
    return t


@pytest.mark.parametrize("l,mbyl,world,rearrange,coset", [(2, 16, 2, 1, 1), (2, 1 << 10, 8, 1, 1), (2, 1 << 13, 4, 0, 0), (4, 1 << 11, 4, 1, 1),
                                                           (2, 1 << 16, 8, 1, 1), (8, 1 << 8, 2, 1, 0)])
def test_king_scatter_stages_on_one_gpu(z, l, mbyl, world, rearrange, coset):
    """zkg_king_stage1_scatter_bn254_dev writes each value into the segment of the rank that owns its output column; with
    the `world` segments on one device the G launches + G stage-2 launches must equal the single-GPU king closure."""
    import torch
    from zksaas_b200 import capi
    lib = z.lib()
    ctx = _dev_ctx(z)
    try:
        n, t, m = 4 * l, l, mbyl * l
        dom = z.Radix2EvaluationDomain.new(m)
        gen = dom.group_gen()
        g = z.Radix2EvaluationDomain.new(2 * m).element(1) if coset else dom.element(0)
        shares = _rand_dev(torch, n * mbyl, 3).reshape(n, mbyl, 4)
        rnd = _rand_dev(torch, mbyl * t, 4)
        full = torch.empty((n, mbyl, 4), dtype=torch.int64, device="cuda")
        capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, n, mbyl, l, gen.ctypes.data, g.ctypes.data,
                                               rearrange, C.c_void_p(rnd.data_ptr()), C.c_void_p(full.data_ptr())))
        cols = mbyl // world
        segs = [torch.full((cols * l, 4), -1, dtype=torch.int64, device="cuda") for _ in range(world)]      # poisoned: every slot must be written
        arr = (C.c_void_p * world)(*[s_.data_ptr() for s_ in segs])
        for r in range(world):
            loc = shares[:, r * cols:(r + 1) * cols, :].contiguous()
            capi.check(lib.zkg_king_stage1_scatter_bn254_dev(ctx, C.c_void_p(loc.data_ptr()), None, n, r * cols, cols, mbyl, l,
                                                             gen.ctypes.data, g.ctypes.data, rearrange, arr, world))
        for r in range(world):
            out = torch.empty((n, cols, 4), dtype=torch.int64, device="cuda")
            rloc = rnd[r * cols * t:(r + 1) * cols * t].contiguous()
            capi.check(lib.zkg_king_stage2_bn254_dev(ctx, C.c_void_p(segs[r].data_ptr()), C.c_void_p(rloc.data_ptr()), cols, l,
                                                     C.c_void_p(out.data_ptr())))
            capi.check(lib.zkg_ctx_sync(ctx))
            assert bool((out == full[:, r * cols:(r + 1) * cols, :]).all()), r
    finally:
        lib.zkg_ctx_destroy(ctx)


@pytest.mark.parametrize("l,mbyl,world", [(2, 1 << 6, 2), (2, 1 << 14, 4), (2, 1 << 18, 8), (8, 1 << 10, 8), (4, 1 << 20, 2)])
def test_fft1_scatter_steps_on_one_gpu(z, l, mbyl, world):
    """zkg_fft1_shard_local_scatter_bn254_dev (inner transform + twiddles + all-to-all by direct stores) followed by the
    outer step reproduces the single-GPU fft1 (pinned against the literal loops in test_gpu_core)."""
    import torch
    from zksaas_b200 import capi, sharding
    lib = z.lib()
    ctx = _dev_ctx(z)
    try:
        px = _rand_dev(torch, mbyl, mbyl + world)
        gen = z.Radix2EvaluationDomain.new(mbyl * l).group_gen()
        full = px.clone()
        capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(full.data_ptr()), mbyl, l, gen.ctypes.data, None, None))
        n2, cnt = mbyl // world, mbyl // world // world
        recvs = [torch.full((n2, 4), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
        arr = (C.c_void_p * world)(*[r_.data_ptr() for r_ in recvs])
        for r in range(world):
            blk = px[r * n2:(r + 1) * n2].clone()
            capi.check(lib.zkg_fft1_shard_local_scatter_bn254_dev(ctx, C.c_void_p(blk.data_ptr()), n2, l, world, r, gen.ctypes.data, None, arr))
        for r in range(world):
            out = torch.empty((world * cnt, 4), dtype=torch.int64, device="cuda")
            capi.check(lib.zkg_fft1_shard_outer_bn254_dev(ctx, C.c_void_p(recvs[r].data_ptr()), cnt, n2, l, world, gen.ctypes.data,
                                                          C.c_void_p(out.data_ptr())))
            capi.check(lib.zkg_ctx_sync(ctx))
            idx = torch.from_numpy(sharding.fft1_sharded_index(mbyl, world, r)).cuda()
            assert bool((out == full[idx]).all()), r
    finally:
        lib.zkg_ctx_destroy(ctx)


def test_sharded_entry_points_with_one_device_match(z):
    """The device-list entry points degrade to the single-GPU ones for a one-element list (what a 1-GPU host gets)."""
    from zksaas_b200 import capi
    lib = z.lib()
    o = ol.oracle()
    rng = np.random.default_rng(2)
    devs = (C.c_int32 * 1)(0)
    n = 1 << 12
    bases = np.zeros((n, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(ol.rand_fr(rng, n)), n, bases.ctypes.data, 72)
    sc = ol.rand_fr(rng, n)
    out = np.zeros(12, dtype=np.uint64)
    capi.check(lib.zkg_msm_bn254_g1_sharded(devs, 1, bases.ctypes.data, 72, n, _p(sc), n, _p(out)))
    assert (out == z.msm_g1(bases, sc)).all()
    assert lib.zkg_msm_bn254_g1_sharded(devs, 1, bases.ctypes.data, 72, n, _p(sc), n - 1, _p(out)) == capi.ZKG_ERR_LEN_MISMATCH
    h = C.c_uint64(0)
    capi.check(lib.zkg_bases_register_sharded(devs, 1, 1, bases.ctypes.data, 72, n, C.byref(h)))
    out2 = np.zeros(12, dtype=np.uint64)
    capi.check(lib.zkg_msm_bn254_registered(h.value, _p(sc), n, _p(out2)))
    capi.check(lib.zkg_bases_release(h.value))
    assert (out2 == out).all()
    l, mbyl = 2, 1 << 10
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    px = ol.rand_fr(rng, mbyl)
    e = z.fft1_in_place(px.copy(), pp, dom.group_gen())
    g_ = px.copy()
    capi.check(lib.zkg_fft1_bn254_sharded(devs, 1, _p(g_), mbyl, l, _p(dom.group_gen()), None, None))
    assert (g_ == e).all()
    bad = (C.c_int32 * 2)(0, 0)
    assert lib.zkg_fft1_bn254_sharded(bad, 2, _p(g_), mbyl, l, _p(dom.group_gen()), None, None) == capi.ZKG_ERR_BAD_ARG


# ------------------------------------------------------------------------------------------------
# Opt-in transparent registration for the unchanged d_msm caller (ZKG_AUTO_REGISTER=1, msm_api.cu): the same host
# pointer seen again is served from a prepared table, the shipped bases being compared on the device with the registered copy
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("g2", [False, True])
def test_auto_register_serves_the_same_results_and_detects_changed_bases(z, monkeypatch, g2):
    o = ol.oracle()
    n = 1 << 16
    rng = np.random.default_rng(77 + g2)
    stride = 136 if g2 else 72
    fixed = o.zko_g2_fixed_base if g2 else o.zko_g1_fixed_base
    msm = z.msm_g2 if g2 else z.msm_g1
    bases = np.zeros((n, stride), dtype=np.uint8)
    fixed(_p(ol.rand_fr(rng, n)), n, bases.ctypes.data, stride)
    bases[5] = 0
    bases[5, stride - 8] = 1                                             # an identity among the bases
    sc = [ol.rand_fr(rng, n) for _ in range(3)]
    monkeypatch.delenv("ZKG_AUTO_REGISTER", raising=False)
    want = [msm(bases, s) for s in sc]                                   # ordinary path
    monkeypatch.setenv("ZKG_AUTO_REGISTER", "1")
    # call 1: first sighting; call 2: pays the preparation; calls 3..: served from the table (different scalars each time)
    got = [msm(bases, sc[k % 3]) for k in range(6)]
    for k in range(6):
        assert (got[k] == want[k % 3]).all(), k
    # the memory behind the SAME pointer changes: one coordinate byte, one whole point, the infinity flag
    other = np.zeros((3, stride), dtype=np.uint8)
    fixed(_p(ol.rand_fr(rng, 3)), 3, other.ctypes.data, stride)
    for mutate in ("byte", "point", "flag"):
        saved = bases[n - 7].copy()
        if mutate == "byte":
            bases[n - 7] = other[0]
        elif mutate == "point":
            bases[n - 7] = other[1]
        else:
            bases[n - 7, stride - 8] = 1
        monkeypatch.delenv("ZKG_AUTO_REGISTER", raising=False)
        ref = msm(bases.copy(), sc[0])                                   # a fresh buffer: never cached
        monkeypatch.setenv("ZKG_AUTO_REGISTER", "1")
        for _ in range(4):                                               # detect + drop, re-sight, re-prepare, serve
            assert (msm(bases, sc[0]) == ref).all(), mutate
        bases[n - 7] = saved
        for _ in range(4):
            assert (msm(bases, sc[0]) == want[0]).all(), mutate


def test_auto_register_budget_and_many_pointers(z, monkeypatch):
    """The transparent registrations are bounded: a byte budget (ZKG_AUTO_REGISTER_MAX_MB; least recently used sets are dropped,
    a set that cannot fit is never registered) and a bounded list of sighted pointers.  Results stay the ordinary path's."""
    o = ol.oracle()
    n = 1 << 16
    rng = np.random.default_rng(99)
    sets = []
    for _ in range(3):
        b = np.zeros((n, 72), dtype=np.uint8)
        o.zko_g1_fixed_base(_p(ol.rand_fr(rng, n)), n, b.ctypes.data, 72)
        sets.append(b)
    sc = ol.rand_fr(rng, n)
    monkeypatch.delenv("ZKG_AUTO_REGISTER", raising=False)
    want = [z.msm_g1(b, sc) for b in sets]
    monkeypatch.setenv("ZKG_AUTO_REGISTER", "1")
    monkeypatch.setenv("ZKG_AUTO_REGISTER_MAX_MB", "1")                  # nothing fits: every call takes the ordinary path
    for _ in range(3):
        for b, w in zip(sets, want):
            assert (z.msm_g1(b, sc) == w).all()
    # room for ONE set of 2^16 points (table 13 x 4 MiB + copies): the three sets keep evicting each other
    monkeypatch.setenv("ZKG_AUTO_REGISTER_MAX_MB", "80")
    for _ in range(4):
        for b, w in zip(sets, want):
            assert (z.msm_g1(b, sc) == w).all()
    # more distinct pointers than the sighting list holds
    monkeypatch.setenv("ZKG_AUTO_REGISTER_MAX_MB", "32768")
    many = [sets[0][: n - 64 * k].copy() for k in range(1, 20)]
    for b in many:
        got = z.msm_g1(b, sc[: b.shape[0]])
    monkeypatch.delenv("ZKG_AUTO_REGISTER", raising=False)
    assert (got == z.msm_g1(many[-1].copy(), sc[: many[-1].shape[0]])).all()
