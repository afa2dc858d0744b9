"""GPU parity tests (B200): every CUDA entry point of the hot path, through the C ABI, against the
CPU oracle on the same seeded inputs -- bit-exact (integer arithmetic; no tolerance)."""
import ctypes as C
import random

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


@pytest.fixture(scope="module")
def o():
    return ol.oracle()


def raw_np(vals):
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = [(v >> (64 * k)) & (2**64 - 1) for k in range(4)]
    return out


def unraw(arr):
    return [sum(int(x) << (64 * k) for k, x in enumerate(r)) for r in arr]


# ------------------------------------------------------------------------------------------------
# field arithmetic (K1)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field,p", [(0, R), (1, Q)])
def test_field_ops_bit_exact(z, field, p):
    from zksaas_b200 import api
    rng = random.Random(11 + field)
    edge = [0, 1, 2, p - 1, p - 2, 1 << 253, (1 << 32) - 1, (1 << 64) - 1, 1 << 224, p >> 1, (p >> 1) + 1]
    a = [rng.randrange(p) for _ in range(5000)] + [x for x in edge for _ in edge]
    b = [rng.randrange(p) for _ in range(5000)] + [y for _ in edge for y in edge]
    A, B = raw_np(a), raw_np(b)
    rinv = pow(1 << 256, -1, p)
    assert unraw(api._field_op(0, A, B, field)) == [x * y * rinv % p for x, y in zip(a, b)]
    assert unraw(api._field_op(1, A, B, field)) == [(x + y) % p for x, y in zip(a, b)]
    assert unraw(api._field_op(2, A, B, field)) == [(x - y) % p for x, y in zip(a, b)]
    # the dedicated squaring (values with saturated limbs and every limb's top bit set exercise the folded doubling),
    # the two-term inner product (one reduction for a*b - (a+1)*b = -b) and the binary-Euclid inverse
    sq = a + [p - 1 - (1 << (32 * k)) for k in range(8)] + [((1 << 254) - 1) % p, int("7fffffff" * 8, 16) % p,
              int("80000000" * 8, 16) % p, int("ffffffff" * 7, 16), (1 << 253) | (1 << 31) | 1]
    S = raw_np(sq)
    assert unraw(api._field_op(3, S, S, field)) == [x * x * rinv % p for x in sq]
    assert unraw(api._field_op(4, A, B, field)) == [(-y) % p for y in b]
    R1 = pow(1 << 256, 1, p)
    assert unraw(api._field_op(5, S, S, field)) == [pow(x * rinv % p, -1, p) * R1 % p if x else 0 for x in sq]


# ------------------------------------------------------------------------------------------------
# MSM (K2 + K3)
# ------------------------------------------------------------------------------------------------
def _g1_points(rng, n):
    dl = [rng.randrange(R) for _ in range(n)]
    lib = ol.oracle()
    out = np.zeros((n, 72), dtype=np.uint8)
    if n:
        lib.zko_g1_fixed_base(_p(ol.fr_np(dl)), n, out.ctypes.data, 72)
    return dl, out


def _g2_points(rng, n):
    dl = [rng.randrange(R) for _ in range(n)]
    lib = ol.oracle()
    out = np.zeros((n, 136), dtype=np.uint8)
    if n:
        lib.zko_g2_fixed_base(_p(ol.fr_np(dl)), n, out.ctypes.data, 136)
    return dl, out


def test_msm_g1_known_answer(z):
    """SURVEY 8c derived KAT: MSM([G,2G,3G,4G],[5, r-1, 0, 2^253])."""
    G = pyref.G1_GEN
    bases = ol.g1_aff_np([pyref.G1.mul(G, k) for k in (1, 2, 3, 4)])
    got = ol.g1_xyz_to_point(z.msm_g1(bases, ol.fr_np([5, R - 1, 0, 1 << 253])))
    assert got == (13029254136549853374917870083460735038808062714769386524861890685576727298910,
                   8493657113625770995739651764829897578810448832951094624325556898429206321085)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 100, 512, 1000])
def test_msm_g1_vs_oracle(z, n):
    rng = random.Random(100 + n)
    _, bases = _g1_points(rng, n)
    scalars = ol.fr_np([rng.randrange(R) for _ in range(n)])
    got = z.msm_g1(bases, scalars)
    exp = ol.o_g1_msm(bases, scalars, threads=8)
    assert (got == exp).all()


def test_msm_g1_edge_cases(z):
    """zero / one / r-1 scalars, infinity bases, all-equal bases (dmsm/mod.rs:144-147), P and -P."""
    rng = random.Random(5)
    dl, bases = _g1_points(rng, 64)
    pts = [pyref.G1.mul(pyref.G1_GEN, d) for d in dl[:4]]
    sc = [rng.randrange(R) for _ in range(64)]
    sc[0], sc[1], sc[2] = 0, 1, R - 1
    bases[3] = np.frombuffer(pyref.g1_affine_image(None), dtype=np.uint8)          # infinity base
    bases[5] = bases[4]; sc[5] = sc[4]                                             # P + P inside a bucket
    bases[7] = np.frombuffer(pyref.g1_affine_image(pyref.G1.neg(pts[0])), dtype=np.uint8)
    bases[6] = np.frombuffer(pyref.g1_affine_image(pts[0]), dtype=np.uint8); sc[7] = sc[6]   # P and -P
    S = ol.fr_np(sc)
    assert (z.msm_g1(bases, S) == ol.o_g1_msm(bases, S)).all()
    # all bases identical, all scalars one  (pack_unpack2_test)
    same = np.repeat(bases[8:9], 256, axis=0)
    ones = ol.fr_np([1] * 256)
    assert (z.msm_g1(same, ones) == ol.o_g1_msm(same, ones)).all()
    # everything cancels -> identity image (1, 1, 0)
    two = np.stack([bases[6], bases[7]])
    ident = z.msm_g1(two, ol.fr_np([7, 7]))
    assert ol.g1_xyz_to_point(ident) is None
    assert (ident == ol.g1_point_to_xyz(None)).all()


def test_msm_length_mismatch(z):
    rng = random.Random(1)
    _, bases = _g1_points(rng, 4)
    with pytest.raises(z.MsmLengthMismatch) as ei:
        z.msm_g1(bases, ol.fr_np([1, 2, 3]))
    assert ei.value.min_len == 3


@pytest.mark.parametrize("n", [0, 1, 5, 33, 200])
def test_msm_g2_vs_oracle(z, n):
    rng = random.Random(200 + n)
    _, bases = _g2_points(rng, n)
    sc = [rng.randrange(R) for _ in range(n)]
    if n >= 5:
        sc[0], sc[1] = 0, R - 1
        bases[2] = np.frombuffer(pyref.g2_affine_image(None), dtype=np.uint8)
        bases[4] = bases[3]; sc[4] = sc[3]
    S = ol.fr_np(sc)
    assert (z.msm_g2(bases, S) == ol.o_g2_msm(bases, S, threads=8)).all()


@pytest.mark.parametrize("c", [5, 8, 11, 13])
def test_msm_window_sizes_agree(z, c, monkeypatch):
    """The window size changes the schedule, never the (normalised) result."""
    rng = random.Random(77)
    _, bases = _g1_points(rng, 300)
    S = ol.fr_np([rng.randrange(R) for _ in range(300)])
    exp = ol.o_g1_msm(bases, S, threads=8)
    monkeypatch.setenv("ZKG_MSM_C", str(c))
    assert (z.msm_g1(bases, S) == exp).all()


def test_msm_registered_bases(z):
    from zksaas_b200 import capi
    rng = random.Random(9)
    _, bases = _g1_points(rng, 128)
    sc = [rng.randrange(R) for _ in range(128)]
    sc[0], sc[1], sc[2] = 0, 1, R - 1
    bases[3] = np.frombuffer(pyref.g1_affine_image(None), dtype=np.uint8)
    bases[5] = bases[4]; sc[5] = sc[4]
    S = ol.fr_np(sc)
    h = C.c_uint64(0)
    capi.check(z.lib().zkg_bases_register(0, 1, bases.ctypes.data, 72, 128, C.byref(h)))
    out = np.zeros(12, dtype=np.uint64)
    capi.check(z.lib().zkg_msm_bn254_registered(h.value, S.ctypes.data, 128, out.ctypes.data))
    assert (out == ol.o_g1_msm(bases, S)).all()
    assert z.lib().zkg_msm_bn254_registered(h.value, S.ctypes.data, 127, out.ctypes.data) == capi.ZKG_ERR_LEN_MISMATCH
    capi.check(z.lib().zkg_bases_release(h.value))


def _skewed_scalars(rng, n, kind):
    if kind == "all_equal":
        return [rng.randrange(R)] * n
    if kind == "all_one":
        return [1] * n
    if kind == "all_minus_one":
        return [R - 1] * n
    if kind == "tiny":                       # witness-like: booleans and small integers, a few full-size values
        return [rng.choice([0, 1, 1, 2, 3, R - 1, rng.randrange(1 << 16)]) if rng.random() < 0.95 else rng.randrange(R)
                for _ in range(n)]
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["all_equal", "all_one", "all_minus_one", "tiny"])
@pytest.mark.parametrize("chunks", [0, 3])
def test_msm_g1_skewed_scalars(z, kind, chunks, monkeypatch):
    """Skewed digit distributions put thousands of points into one bucket (block-cooperative
    k_accumulate_heavy); also through the chunked path, where heavy buckets accumulate across chunks."""
    rng = random.Random(hash(kind) & 0xffff)
    n = 1 << 14
    _, bases = _g1_points(rng, n)
    S = ol.fr_np(_skewed_scalars(rng, n, kind))
    if chunks:
        monkeypatch.setenv("ZKG_MSM_CHUNKS", str(chunks))
    assert (z.msm_g1(bases, S) == ol.o_g1_msm(bases, S, threads=8)).all()


@pytest.mark.parametrize("reps", [41, 100, 150, 192])
def test_msm_g1_repeated_points_in_one_bucket(z, reps):
    """Buckets of 40..192 points made of identical and opposite points: the pairwise (batched-affine,
    ZKG_MSM_BA=1) accumulation meets P + P, P + (-P) and infinity operands in every tree round; the
    chain path sees the same input.  Compared with the oracle."""
    rng = random.Random(reps)
    dl, bases = _g1_points(rng, 8)
    pts = [pyref.G1.mul(pyref.G1_GEN, d) for d in dl]
    neg0 = np.frombuffer(pyref.g1_affine_image(pyref.G1.neg(pts[0])), dtype=np.uint8)
    inf = np.frombuffer(pyref.g1_affine_image(None), dtype=np.uint8)
    s0, s1 = rng.randrange(R), rng.randrange(R)
    rows, sc = [], []
    for i in range(reps):                                   # one scalar -> the same bucket in every window
        kind = i % 7
        rows.append(bases[0] if kind in (0, 1, 2) else neg0 if kind == 3 else inf if kind == 4 else bases[1 + kind % 3])
        sc.append(s0)
    for i in range(reps):                                   # all-equal bases and scalars (dmsm/mod.rs:144-147)
        rows.append(bases[5]); sc.append(s1)
    for i in range(reps + 1):                               # exact cancellation inside one bucket
        rows.append(bases[6] if i % 2 == 0 else np.frombuffer(pyref.g1_affine_image(pyref.G1.neg(pts[6])), dtype=np.uint8))
        sc.append((s1 * 3 + 1) % R)
    B = np.stack(rows)
    S = ol.fr_np(sc)
    assert (z.msm_g1(B, S) == ol.o_g1_msm(B, S, threads=8)).all()


def test_msm_skewed_registered_and_g2(z):
    from zksaas_b200 import capi
    rng = random.Random(4242)
    n = 1 << 13
    _, bases = _g1_points(rng, n)
    h = C.c_uint64(0)
    capi.check(z.lib().zkg_bases_register(0, 1, bases.ctypes.data, 72, n, C.byref(h)))
    out = np.zeros(12, dtype=np.uint64)
    for kind in ("all_equal", "tiny"):
        S = ol.fr_np(_skewed_scalars(rng, n, kind))
        capi.check(z.lib().zkg_msm_bn254_registered(h.value, S.ctypes.data, n, out.ctypes.data))
        assert (out == ol.o_g1_msm(bases, S, threads=8)).all()
    capi.check(z.lib().zkg_bases_release(h.value))
    n2 = 1 << 12
    _, b2 = _g2_points(rng, n2)
    for kind in ("all_equal", "tiny"):
        S = ol.fr_np(_skewed_scalars(rng, n2, kind))
        assert (z.msm_g2(b2, S) == ol.o_g2_msm(b2, S, threads=8)).all()


# ------------------------------------------------------------------------------------------------
# fft1 (K4)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("l,mbyl,world", [(2, 16, 1), (2, 16, 2), (2, 64, 4), (2, 1 << 13, 8), (4, 1 << 16, 4), (2, 1 << 20, 8)])
def test_fft1_sharded_steps_on_one_gpu(z, l, mbyl, world):
    """The two CUDA steps of the rank-sharded fft1 with the all-to-all emulated by slicing on one device:
    together they must reproduce the single-GPU fft1 (which the next test pins against the literal loops)."""
    import torch
    from zksaas_b200 import capi, sharding
    lib = z.lib()
    ctx = capi.ctx_p()
    capi.check(lib.zkg_ctx_create(0, C.c_void_p(1), C.byref(ctx)))
    try:
        g = torch.Generator(device="cuda"); g.manual_seed(mbyl + world)
        px = torch.randint(-2**63, 2**63 - 1, (mbyl, 4), dtype=torch.int64, device="cuda", generator=g)
        px[:, 3] &= (1 << 61) - 1
        gen = z.Radix2EvaluationDomain.new(mbyl * l).group_gen()
        full = px.clone()
        capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(full.data_ptr()), mbyl, l, gen.ctypes.data, None, None))
        n2, cnt = mbyl // world, mbyl // world // world
        sends = []
        for r in range(world):
            blk = px[r * n2:(r + 1) * n2].clone()
            capi.check(lib.zkg_fft1_shard_local_bn254_dev(ctx, C.c_void_p(blk.data_ptr()), n2, l, world, r, gen.ctypes.data, None))
            sends.append(blk)
        for r in range(world):
            recv = torch.cat([sends[src][r * cnt:(r + 1) * cnt] for src in range(world)]).contiguous()
            out = torch.empty((world * cnt, 4), dtype=torch.int64, device="cuda")
            capi.check(lib.zkg_fft1_shard_outer_bn254_dev(ctx, C.c_void_p(recv.data_ptr()), cnt, n2, l, world, gen.ctypes.data,
                                                          C.c_void_p(out.data_ptr())))
            capi.check(lib.zkg_ctx_sync(ctx))
            idx = torch.from_numpy(sharding.fft1_sharded_index(mbyl, world, r)).cuda()
            assert bool((out == full[idx]).all()), r
    finally:
        lib.zkg_ctx_destroy(ctx)


@pytest.mark.parametrize("l,mbyl,world,rearrange,coset", [(2, 16, 2, 1, 1), (2, 1 << 10, 4, 0, 0), (2, 1 << 13, 8, 1, 1),
                                                           (2, 1 << 10, 8, 1, 1), (2, 64, 4, 0, 1), (4, 1 << 11, 4, 1, 1), (2, 1 << 16, 2, 1, 1)])
def test_king_sharded_stages_on_one_gpu(z, l, mbyl, world, rearrange, coset):
    """The two CUDA stages of the column-sharded king pipeline, with the reduce-scatter emulated on one device
    (sum of the ranks' full-size pack-order buffers, sliced): together they must reproduce the single-GPU king
    closure bit for bit.  Exercises column ranges that do not start at 0 (and, for 2^10 columns over 8 ranks or 64
    over 4, ranges that are not multiples of the 256-aligned exponent tiles of the l = 2 kernel)."""
    import torch
    from zksaas_b200 import capi
    lib = z.lib()
    ctx = capi.ctx_p()
    capi.check(lib.zkg_ctx_create(0, C.c_void_p(1), C.byref(ctx)))
    try:
        n, t = 4 * l, l
        m = mbyl * l
        g_ = torch.Generator(device="cuda"); g_.manual_seed(mbyl * 31 + world)

        def rnd(k):
            x = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g_)
            x[:, 3] &= (1 << 61) - 1
            return x
        shares = rnd(n * mbyl).reshape(n, mbyl, 4)
        rand = rnd(mbyl * t)
        gen = z.Radix2EvaluationDomain.new(m).group_gen()
        g = z.Radix2EvaluationDomain.new(2 * m).element(1) if coset else z.Radix2EvaluationDomain.new(m).element(0)
        full = torch.empty((n, mbyl, 4), dtype=torch.int64, device="cuda")
        capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, n, mbyl, l, gen.ctypes.data, g.ctypes.data,
                                               rearrange, C.c_void_p(rand.data_ptr()), C.c_void_p(full.data_ptr())))
        cols = mbyl // world
        S_sum = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
        for r in range(world):
            loc = shares[:, r * cols:(r + 1) * cols].contiguous()
            S = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
            capi.check(lib.zkg_king_stage1_bn254_dev(ctx, C.c_void_p(loc.data_ptr()), None, n, r * cols, cols, mbyl, l,
                                                     gen.ctypes.data, g.ctypes.data, rearrange, C.c_void_p(S.data_ptr())))
            capi.check(lib.zkg_ctx_sync(ctx))
            S_sum += S                                   # every slot is written by exactly one rank: lane sums never carry
        for r in range(world):
            S_r = S_sum[r * cols * l:(r + 1) * cols * l].contiguous()
            rd = rand[r * cols * t:(r + 1) * cols * t].contiguous()
            out = torch.empty((n, cols, 4), dtype=torch.int64, device="cuda")
            capi.check(lib.zkg_king_stage2_bn254_dev(ctx, C.c_void_p(S_r.data_ptr()), C.c_void_p(rd.data_ptr()), cols, l,
                                                     C.c_void_p(out.data_ptr())))
            capi.check(lib.zkg_ctx_sync(ctx))
            assert bool((out == full[:, r * cols:(r + 1) * cols]).all()), r
    finally:
        lib.zkg_ctx_destroy(ctx)


@pytest.mark.parametrize("l,mbyl", [(2, 1), (2, 2), (2, 4), (2, 8), (2, 512), (2, 1024), (2, 2048), (2, 1 << 15),
                                    (4, 16), (4, 1 << 11), (8, 1 << 12), (2, 1 << 17)])
def test_fft1_vs_literal_oracle(z, o, l, mbyl):
    rng = np.random.default_rng(mbyl * 7 + l)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    px = ol.rand_fr(rng, mbyl)
    exp = px.copy()
    gen = dom.group_gen()
    o.zko_fft1_in_place(_p(exp), mbyl, l, _p(gen))
    got = px.copy()
    z.fft1_in_place(got, pp, gen)
    assert (got == exp).all()


@pytest.mark.parametrize("mbyl", [1, 2, 4, 8, 16, 32, 128, 512, 1 << 10, 1 << 11, 1 << 12, 1 << 14, 1 << 16, 1 << 18])
@pytest.mark.parametrize("elog,tables,ept_log", [(11, 1, 3), (9, 1, 3), (11, 0, 3), (10, 1, 2), (8, 0, 2)])
def test_fft1_radix8_kernel_all_shapes(z, o, monkeypatch, mbyl, elog, tables, ept_log):
    """The register-blocked radix-8 pass kernel (production for N >= 2^18) forced onto every transform size, so that
    all its group shapes (1, 2 and 3 stages per group; fewer than 8 elements; one, two and three passes; narrow
    tiles) and both twiddle sources (per-pass table / on-the-fly powers) are pinned against the literal loops of
    fft1_in_place."""
    monkeypatch.setenv("ZKG_NTT_R8_MIN", "0")
    monkeypatch.setenv("ZKG_NTT_ELOG", str(elog))
    monkeypatch.setenv("ZKG_NTT_EPT_LOG", str(ept_log))          # 8 or 4 elements per thread (radix-8 / radix-4 groups)
    if not tables:
        monkeypatch.setenv("ZKG_NTT_ONTHEFLY", "1")
    l = 2
    rng = np.random.default_rng(mbyl * 13 + elog)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    px = ol.rand_fr(rng, mbyl)
    exp = px.copy()
    gen = dom.group_gen()
    o.zko_fft1_in_place(_p(exp), mbyl, l, _p(gen))
    got = px.copy()
    z.fft1_in_place(got, pp, gen)
    assert (got == exp).all()


def test_parameter_cache_flush_mid_call(z, monkeypatch):
    """The per-context parameter cache (power tables, per-pass twiddle tables, scaled unpack matrices) is bounded in
    bytes; with a 1 MiB budget it is flushed in the middle of calls.  Results must not change, and pointers handed out
    before a flush must stay valid until the call returns (deferred frees)."""
    import torch
    from zksaas_b200 import capi
    lib = z.lib()
    l, mbyl = 2, 1 << 14
    m = mbyl * l
    g_ = torch.Generator(device="cuda"); g_.manual_seed(77)

    def rnd(k):
        x = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g_)
        x[:, 3] &= (1 << 61) - 1
        return x
    shares, rand, px = rnd(8 * mbyl), rnd(2 * mbyl), rnd(mbyl)
    gen = z.Radix2EvaluationDomain.new(m).group_gen()
    cosets = [z.Radix2EvaluationDomain.new(2 * m).element(k) for k in (1, 3, 5)]

    def run_all():
        ctx = capi.ctx_p()
        capi.check(lib.zkg_ctx_create(0, C.c_void_p(1), C.byref(ctx)))
        outs = []
        try:
            for rep in range(2):
                for g in cosets:
                    out = torch.empty((8 * mbyl, 4), dtype=torch.int64, device="cuda")
                    capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, 8, mbyl, l, gen.ctypes.data,
                                                           g.ctypes.data, 1, C.c_void_p(rand.data_ptr()), C.c_void_p(out.data_ptr())))
                    f = px.clone()
                    capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(f.data_ptr()), mbyl, l, g.ctypes.data, None, None))
                    capi.check(lib.zkg_ctx_sync(ctx))
                    outs.append((out, f))
        finally:
            lib.zkg_ctx_destroy(ctx)
        return outs
    base = run_all()
    monkeypatch.setenv("ZKG_CACHE_MAX_MB", "1")
    small = run_all()
    for (a0, b0), (a1, b1) in zip(base, small):
        assert bool((a0 == a1).all()) and bool((b0 == b1).all())


def test_fft1_fused_scale_and_mask(z, o):
    l, mbyl = 2, 1 << 12
    rng = np.random.default_rng(3)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    px, mask = ol.rand_fr(rng, mbyl), ol.rand_fr(rng, mbyl)
    sinv = dom.size_inv()
    exp = px.copy()
    o.zko_fr_mul(_p(exp), _p(np.repeat(sinv[None], mbyl, axis=0)), _p(exp), mbyl)       # dfft/mod.rs:159
    o.zko_fft1_in_place(_p(exp), mbyl, l, _p(dom.group_gen_inv()))                       # :162
    o.zko_fr_add(_p(exp), _p(mask), _p(exp), mbyl)                                       # :254-258
    got = px.copy()
    z.fft1_in_place(got, pp, dom.group_gen_inv(), pre_scale=sinv, in_mask=mask)
    assert (got == exp).all()


# ------------------------------------------------------------------------------------------------
# PSS (K6)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("l", [2, 4, 8])
def test_pss_pack_unpack_vs_oracle(z, o, l):
    rng = np.random.default_rng(40 + l)
    pp = z.PackedSharingParams.new(l)
    cols = 333
    sec, rnd = ol.rand_fr(rng, cols * l), ol.rand_fr(rng, cols * l)
    exp = np.zeros((cols * pp.n, 4), dtype=np.uint64)
    o.zko_pss_pack_fr(l, _p(sec), _p(rnd), _p(exp), cols)
    shares = pp.pack(sec, rnd)
    assert (shares == exp).all()
    o.zko_pss_pack_fr(l, _p(sec), None, _p(exp), cols)
    assert (pp.det_pack(sec) == exp).all()
    assert (pp.unpack(shares) == sec).all()                       # pss.rs:270-271
    assert (pp.unpack2(shares) == sec).all()
    assert (pp.unpack(pp.det_pack(sec)) == sec).all()             # pss.rs:287
    # products of shares unpack2 to products of secrets (pss.rs:299-310)
    from zksaas_b200 import api
    sq = api.fr_mul(shares, shares)
    e2 = np.zeros((cols * l, 4), dtype=np.uint64)
    o.zko_pss_unpack2_fr(l, _p(sq), _p(e2), cols)
    assert (pp.unpack2(sq) == e2).all()
    assert (e2 == api.fr_mul(sec, sec)).all()


# ------------------------------------------------------------------------------------------------
# king pipeline (K5) and deg_red
# ------------------------------------------------------------------------------------------------
def _valid_shares(o, rng, l, mbyl):
    """party-major degree-2(l+t)-2 sharings: squares of random packed sharings."""
    n = 4 * l
    sec, rnd = ol.rand_fr(rng, mbyl * l), ol.rand_fr(rng, mbyl * l)
    sh = np.zeros((mbyl * n, 4), dtype=np.uint64)
    o.zko_pss_pack_fr(l, _p(sec), _p(rnd), _p(sh), mbyl)
    o.zko_fr_mul(_p(sh), _p(sh), _p(sh), mbyl * n)
    cols = sh.reshape(mbyl, n, 4)
    return [np.ascontiguousarray(cols[:, p, :]) for p in range(n)]


def _oracle_king(o, shares, parties, mbyl, l, gen, g, rearrange, rand):
    n = 4 * l
    outs = [np.zeros((mbyl, 4), dtype=np.uint64) for _ in range(n)]
    par = (C.c_uint32 * len(parties))(*parties)
    rc = o.zko_king_fft2(ol.ptr_array(shares), par, len(shares), mbyl, l, _p(gen), _p(g), rearrange, _p(rand),
                         ol.ptr_array(outs))
    assert rc == 0
    return outs


@pytest.mark.parametrize("l,mbyl", [(2, 1), (2, 4), (2, 64), (2, 1 << 10), (2, 1 << 15), (4, 8), (4, 1 << 9), (8, 1 << 6)])
@pytest.mark.parametrize("rearrange", [0, 1])
def test_king_fft2_vs_literal_oracle(z, o, l, mbyl, rearrange):
    rng = np.random.default_rng(l * 1000 + mbyl + rearrange)
    pp = z.PackedSharingParams.new(l)
    m = mbyl * l
    dom = z.Radix2EvaluationDomain.new(m)
    shares = _valid_shares(o, rng, l, mbyl)
    rand = ol.rand_fr(rng, mbyl * l)
    zeta = z.Radix2EvaluationDomain.new(2 * m).element(1)
    from zksaas_b200 import api
    for gen in (dom.group_gen(), dom.group_gen_inv()):
        for g in (api.fr_image(1), zeta, api.fr_image(5)):
            exp = _oracle_king(o, shares, list(range(pp.n)), mbyl, l, gen, g, rearrange, rand)
            got = z.king_fft2(shares, list(range(pp.n)), pp, gen, g, rearrange, rand)
            for p in range(pp.n):
                assert (got[p] == exp[p]).all(), (p,)


def test_king_fft2_dropout_uses_lagrange(z, o):
    """One party missing: unpack_missing_shares falls back to lagrange_unpack (pss.rs:210-221)."""
    l, mbyl = 2, 256
    rng = np.random.default_rng(8)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    shares = _valid_shares(o, rng, l, mbyl)
    rand = ol.rand_fr(rng, mbyl * l)
    from zksaas_b200 import api
    for missing in (7, 0, 3):
        parties = [p for p in range(pp.n) if p != missing]
        sub = [shares[p] for p in parties]
        exp = _oracle_king(o, sub, parties, mbyl, l, dom.group_gen(), api.fr_image(1), 1, rand)
        full = _oracle_king(o, shares, list(range(pp.n)), mbyl, l, dom.group_gen(), api.fr_image(1), 1, rand)
        got = z.king_fft2(sub, parties, pp, dom.group_gen(), api.fr_image(1), True, rand)
        for p in range(pp.n):
            assert (got[p] == exp[p]).all() and (got[p] == full[p]).all()
    with pytest.raises(z.ZkgError):       # two parties missing: not enough shares for degree 2(l+t)-2
        z.king_fft2(shares[:6], list(range(6)), pp, dom.group_gen(), api.fr_image(1), True, rand)


@pytest.mark.parametrize("l,cols", [(2, 1000), (4, 37)])
def test_deg_red_king_vs_oracle(z, o, l, cols):
    rng = np.random.default_rng(cols)
    pp = z.PackedSharingParams.new(l)
    shares = _valid_shares(o, rng, l, cols)
    rand = ol.rand_fr(rng, cols * l)
    outs = [np.zeros((cols, 4), dtype=np.uint64) for _ in range(pp.n)]
    par = (C.c_uint32 * pp.n)(*range(pp.n))
    assert o.zko_deg_red_king(ol.ptr_array(shares), par, pp.n, cols, l, _p(rand), ol.ptr_array(outs)) == 0
    got = z.deg_red_king(shares, list(range(pp.n)), pp, rand)
    for p in range(pp.n):
        assert (got[p] == outs[p]).all()


# ------------------------------------------------------------------------------------------------
# stand-alone pieces
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("l,m", [(2, 8), (2, 1 << 12), (4, 1 << 10), (8, 64)])
def test_fft2_standalone(z, o, l, m):
    rng = np.random.default_rng(m + l)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    s1 = ol.rand_fr(rng, m)
    exp = s1.copy()
    o.zko_fft2_in_place(_p(exp), m, l, _p(dom.group_gen()))
    assert (z.fft2_in_place(s1.copy(), pp, dom.group_gen()) == exp).all()


@pytest.mark.parametrize("n", [1, 2, 8, 1 << 13, 1 << 16])
def test_bitrev_and_powers(z, o, n):
    rng = np.random.default_rng(n)
    v = ol.rand_fr(rng, n)
    exp = v.copy()
    o.zko_fr_rearrange(_p(exp), n)
    assert (z.fft_in_place_rearrange(v.copy()) == exp).all()
    g = ol.rand_fr(rng, 1)[0]
    exp = v.copy()
    o.zko_fr_distribute_powers(_p(exp), n, _p(g))
    assert (z.distribute_powers(v.copy(), g) == exp).all()


@pytest.mark.parametrize("n", [1, 2, 16, 1 << 10, 1 << 11, 1 << 16, 1 << 21])
def test_domain_fft_ifft_coset(z, o, n):
    rng = np.random.default_rng(n + 1)
    dom = z.Radix2EvaluationDomain.new(n)
    from zksaas_b200 import api
    v = ol.rand_fr(rng, n)
    for off in (None, api.fr_image(5)):
        for inv in (0, 1):
            exp = v.copy()
            o.zko_fr_fft(_p(exp), n, _p(off) if off is not None else None, inv)
            got = dom.ifft(v, off) if inv else dom.fft(v, off)
            assert (got == exp).all(), (off is not None, inv)
    assert (dom.ifft(dom.fft(v)) == v).all()


def test_fr_wire_format_roundtrip(z):
    """ark-serialize compressed Fr (32 LE bytes, canonical) <-> Montgomery images; >= r is rejected."""
    from zksaas_b200 import capi
    rng = random.Random(6)
    vals = [0, 1, R - 1, 1 << 253] + [rng.randrange(R) for _ in range(1000)]
    wire = raw_np(vals)                                   # canonical little-endian limbs == the 32 wire bytes
    mont = np.zeros_like(wire)
    capi.check(z.lib().zkg_fr_from_wire_bn254(0, wire.ctypes.data, mont.ctypes.data, len(vals)))
    assert (mont == ol.fr_np(vals)).all()
    back = np.zeros_like(wire)
    capi.check(z.lib().zkg_fr_to_wire_bn254(0, mont.ctypes.data, back.ctypes.data, len(vals)))
    assert (back == wire).all()
    bad = raw_np([5, R, 7])
    assert z.lib().zkg_fr_from_wire_bn254(0, bad.ctypes.data, mont.ctypes.data, 3) == capi.ZKG_ERR_BAD_ARG
