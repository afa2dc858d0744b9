"""GPU parity at BASELINE sizes through size-independent properties (the oracle cannot run 2^22-point
MSMs in seconds): bases are s_i * G with known discrete logs, so MSM(bases, a) must equal
((sum_i a_i s_i) mod r) * G -- an O(k) field computation done by the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import _p

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import zksaas_b200 as z
    from zksaas_b200 import capi
    ctx = capi.ctx_p()
    # cudaStreamLegacy (handle 1): the library's kernels are then ordered with torch's default-stream
    # work (random fills, slice assignments) instead of racing it on a private non-blocking stream
    capi.check(z.lib().zkg_ctx_create(0, C.c_void_p(1), C.byref(ctx)))
    yield z, capi, torch, ctx
    z.lib().zkg_ctx_destroy(ctx)


def _rand_dev(torch, k, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 61) - 1
    return t


def _closed_form(o, a_img, s_img, g2=False):
    n = a_img.shape[0]
    prod = np.zeros_like(a_img)
    o.zko_fr_mul(_p(a_img), _p(s_img), _p(prod), n)
    tot = np.zeros(4, dtype=np.uint64)
    o.zko_fr_sum(_p(prod), n, _p(tot))
    if g2:
        aff = np.zeros((1, 136), dtype=np.uint8)
        o.zko_g2_fixed_base(_p(tot), 1, aff.ctypes.data, 136)
        out = np.zeros(24, dtype=np.uint64)
        one = ol.fr_np([1])
        o.zko_g2_msm(aff.ctypes.data, 136, _p(one), 1, _p(out), 1, 0)
        return out
    aff = np.zeros((1, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(tot), 1, aff.ctypes.data, 72)
    return ol.o_g1_msm(aff, ol.fr_np([1]))


@pytest.mark.parametrize("log2n", [14, 18, 20, 22, 24])
def test_msm_g1_closed_form(env, log2n):
    z, capi, torch, ctx = env
    o = ol.oracle()
    n = 1 << log2n
    a, s = _rand_dev(torch, n, 1000 + log2n), _rand_dev(torch, n, 2000 + log2n)
    bases = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    out = torch.zeros(12, dtype=torch.int64, device="cuda")
    lib = z.lib()
    capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(out.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    got = out.cpu().numpy().view(np.uint64)
    exp = _closed_form(o, a.cpu().numpy().view(np.uint64), s.cpu().numpy().view(np.uint64))
    assert (got == exp).all()
    if log2n == 14:
        # the device generator itself vs the oracle's scalar multiplication, and the host-pointer path
        hb = np.zeros((n, 72), dtype=np.uint8)
        hb[:, :64] = bases.cpu().numpy()
        ref = np.zeros((64, 72), dtype=np.uint8)
        o.zko_g1_fixed_base(_p(np.ascontiguousarray(s.cpu().numpy().view(np.uint64)[:64])), 64, ref.ctypes.data, 72)
        assert (hb[:64] == ref).all()
        assert (z.msm_g1(hb, a.cpu().numpy().view(np.uint64)) == exp).all()


@pytest.mark.parametrize("kind", ["witness_like", "all_equal"])
def test_msm_g1_skewed_closed_form_2p20(env, kind):
    """Groth16 witnesses are mostly booleans / small values: 2^20 points whose scalars pile into a
    handful of buckets must stay exact AND must not serialise on one thread per bucket."""
    import time
    z, capi, torch, ctx = env
    o = ol.oracle()
    n = 1 << 20
    a, s = _rand_dev(torch, n, 91), _rand_dev(torch, n, 92)
    one = torch.from_numpy(ol.fr_np([1]).view(np.int64)).cuda()
    if kind == "witness_like":
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        u = torch.rand(n, device="cuda", generator=g)
        a[u < 0.6] = one
        a[u < 0.3] = 0
    else:
        a[:] = a[0].clone()
    bases = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    out = torch.zeros(12, dtype=torch.int64, device="cuda")
    lib = z.lib()
    capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    t0 = time.perf_counter()
    capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(out.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    dt = time.perf_counter() - t0
    got = out.cpu().numpy().view(np.uint64)
    exp = _closed_form(o, a.cpu().numpy().view(np.uint64), s.cpu().numpy().view(np.uint64))
    assert (got == exp).all()
    print(f"skewed {kind}: {dt * 1e3:.1f} ms")
    assert dt < 1.0, f"skewed MSM took {dt:.2f} s"


@pytest.mark.parametrize("log2n", [12, 17])
def test_msm_g2_closed_form(env, log2n):
    z, capi, torch, ctx = env
    o = ol.oracle()
    n = 1 << log2n
    a, s = _rand_dev(torch, n, 3000 + log2n), _rand_dev(torch, n, 4000 + log2n)
    bases = torch.empty((n, 128), dtype=torch.uint8, device="cuda")
    out = torch.zeros(24, dtype=torch.int64, device="cuda")
    lib = z.lib()
    capi.check(lib.zkg_fixed_base_dev(ctx, 2, C.c_void_p(s.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    capi.check(lib.zkg_msm_bn254_g2_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(out.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    got = out.cpu().numpy().view(np.uint64)
    exp = _closed_form(o, a.cpu().numpy().view(np.uint64), s.cpu().numpy().view(np.uint64), g2=True)
    assert (got == exp).all()


def test_msm_partials_combine_like_one_msm(env):
    """Point-range sharding (multi-GPU path): partial XYZZ sums of two halves combine to the full MSM."""
    z, capi, torch, ctx = env
    n = 1 << 16
    a, s = _rand_dev(torch, n, 7), _rand_dev(torch, n, 8)
    bases = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    lib = z.lib()
    capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    full = torch.zeros(12, dtype=torch.int64, device="cuda")
    capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(full.data_ptr())))
    parts = torch.zeros((3, 16), dtype=torch.int64, device="cuda")
    h = n // 2
    capi.check(lib.zkg_msm_bn254_partial_dev(ctx, 1, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), h, C.c_void_p(parts[0].data_ptr())))
    capi.check(lib.zkg_msm_bn254_partial_dev(ctx, 1, C.c_void_p(bases[h:].data_ptr()), C.c_void_p(a[h:].data_ptr()), h, C.c_void_p(parts[1].data_ptr())))
    capi.check(lib.zkg_msm_bn254_partial_dev(ctx, 1, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), 0, C.c_void_p(parts[2].data_ptr())))
    comb = torch.zeros(12, dtype=torch.int64, device="cuda")
    capi.check(lib.zkg_msm_combine_dev(ctx, 1, C.c_void_p(parts.data_ptr()), 3, C.c_void_p(comb.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    assert bool((comb == full).all())


@pytest.mark.parametrize("log2m", [16, 20, 22])
def test_d_fft_round_reconstructs_plain_fft(env, log2m):
    """dfft/tests.rs:78-139 at BASELINE sizes: client fft1 on 8 GPUs' worth of shares + king pipeline,
    then unpack == Radix2EvaluationDomain::fft of the clear vector (computed by the device plain FFT,
    which test_gpu_core pins against the oracle)."""
    z, capi, torch, ctx = env
    l, m = 2, 1 << log2m
    mbyl = m // l
    rng = np.random.default_rng(log2m)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    x = ol.rand_fr(rng, m)
    expect = dom.fft(x)
    xr = z.fft_in_place_rearrange(x.copy())
    # column i packs (xr[i], xr[i + m/l]) : dfft/tests.rs:29-39
    secrets = np.ascontiguousarray(xr.reshape(l, mbyl, 4).transpose(1, 0, 2)).reshape(-1, 4)
    shares = z.transpose(z.pack_vec(secrets, pp, ol.rand_fr(rng, mbyl * pp.t)))
    masks = [z.FftMask.zero(mbyl) for _ in range(pp.n)]
    out = z.d_fft(shares, masks, False, dom, pp, z.LocalTestNet(pp.n), ol.rand_fr(rng, mbyl * pp.t))
    cols = np.stack(out, axis=1).reshape(-1, 4)           # column-major (mbyl x n)
    got = pp.unpack(cols)
    assert (got == expect).all()


@pytest.mark.parametrize("group,log2n", [(1, 10), (1, 15), (1, 19), (1, 22), (1, 24), (2, 13), (2, 19)])
def test_msm_registered_dev_closed_form(env, group, log2n):
    """Prepared bases (window-shifted table, merged buckets, no Horner) give the same group element."""
    z, capi, torch, ctx = env
    o = ol.oracle()
    n = 1 << log2n
    a, s = _rand_dev(torch, n, 5000 + log2n), _rand_dev(torch, n, 6000 + log2n)
    a[0] = 0                                              # a zero scalar
    a[1, :] = torch.tensor([-1, -1, -1, (1 << 61) - 1])   # all window bits set (carries ripple to the top window)
    pk = 64 if group == 1 else 128
    bases = torch.empty((n, pk), dtype=torch.uint8, device="cuda")
    lib = z.lib()
    capi.check(lib.zkg_fixed_base_dev(ctx, group, C.c_void_p(s.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    bases[5] = 0                                          # an infinity base
    s_h = s.cpu().numpy().view(np.uint64).copy()
    s_h[5] = 0
    h = C.c_uint64(0)
    capi.check(lib.zkg_bases_register_dev(ctx, group, C.c_void_p(bases.data_ptr()), n, C.byref(h)))
    out = torch.zeros(12 * group, dtype=torch.int64, device="cuda")
    capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h.value, C.c_void_p(a.data_ptr()), n, C.c_void_p(out.data_ptr()), 0))
    plain = torch.zeros(12 * group, dtype=torch.int64, device="cuda")
    fn = lib.zkg_msm_bn254_g1_dev if group == 1 else lib.zkg_msm_bn254_g2_dev
    capi.check(fn(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(plain.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    exp = _closed_form(o, a.cpu().numpy().view(np.uint64).copy(), s_h, g2=(group == 2))
    assert (out.cpu().numpy().view(np.uint64) == exp).all()
    assert bool((plain == out).all())
    # partial (XYZZ) output combines to the same point
    part = torch.zeros(16 * group, dtype=torch.int64, device="cuda")
    comb = torch.zeros(12 * group, dtype=torch.int64, device="cuda")
    capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h.value, C.c_void_p(a.data_ptr()), n, C.c_void_p(part.data_ptr()), 1))
    capi.check(lib.zkg_msm_combine_dev(ctx, group, C.c_void_p(part.data_ptr()), 1, C.c_void_p(comb.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    assert bool((comb == out).all())
    assert lib.zkg_msm_bn254_registered_dev(ctx, h.value, C.c_void_p(a.data_ptr()), n - 1, C.c_void_p(out.data_ptr()), 0) == capi.ZKG_ERR_LEN_MISMATCH
    capi.check(lib.zkg_bases_release(h.value))


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("envs", ["ZKG_MSM_GROUP0=2", "ZKG_MSM_GROUP0=3;ZKG_MSM_SIDE=0", "ZKG_MSM_GROUP0=1;ZKG_MSM_GROUPS=3",
                                  "ZKG_MSM_GROUP0=2;ZKG_MSM_SORT_BPS_HIDDEN=4;ZKG_MSM_SCATTER_ILP_HIDDEN=1;ZKG_MSM_DIGITS_TB_HIDDEN=128",
                                  "ZKG_MSM_GROUP0=2;ZKG_MSM_SCATTER_ILP_HIDDEN=8;ZKG_MSM_HEAVY_KEY=3",
                                  "ZKG_MSM_HEAVY_KEY=2", "ZKG_MSM_REDUCE_L=1", "ZKG_MSM_REDUCE_L=8;ZKG_MSM_COOP_REDUCE=0"])
def test_msm_device_resident_switches(env, monkeypatch, group, envs):
    """The sort pipeline (window groups, side stream, thin hidden sorts), the heavy-bucket threshold and the reduction shapes
    change the schedule of a device-resident MSM, never the group element: registered and generic path, closed form."""
    z, capi, torch, ctx = env
    o = ol.oracle()
    lib = z.lib()
    n = (1 << 13) + 77
    a, s = _rand_dev(torch, n, 7100 + group), _rand_dev(torch, n, 7200 + group)
    a[0::5] = a[0]                                        # long buckets
    a[3] = 0
    pk = 64 if group == 1 else 128
    bases = torch.empty((n, pk), dtype=torch.uint8, device="cuda")
    capi.check(lib.zkg_fixed_base_dev(ctx, group, C.c_void_p(s.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    exp = _closed_form(o, a.cpu().numpy().view(np.uint64).copy(), s.cpu().numpy().view(np.uint64).copy(), g2=(group == 2))
    h = C.c_uint64(0)
    capi.check(lib.zkg_bases_register_dev(ctx, group, C.c_void_p(bases.data_ptr()), n, C.byref(h)))
    fn = lib.zkg_msm_bn254_g1_dev if group == 1 else lib.zkg_msm_bn254_g2_dev
    for kv in envs.split(";"):
        monkeypatch.setenv(*kv.split("="))
    for _ in range(2):                                    # twice: the second call reuses the sort sets and events
        out = torch.zeros(12 * group, dtype=torch.int64, device="cuda")
        plain = torch.zeros(12 * group, dtype=torch.int64, device="cuda")
        capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h.value, C.c_void_p(a.data_ptr()), n, C.c_void_p(out.data_ptr()), 0))
        capi.check(fn(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(plain.data_ptr())))
        capi.check(lib.zkg_ctx_sync(ctx))
        assert (out.cpu().numpy().view(np.uint64) == exp).all()
        assert (plain.cpu().numpy().view(np.uint64) == exp).all()
    capi.check(lib.zkg_bases_release(h.value))


def test_d_fft_2p24_sampled_against_the_dft_definition(env):
    """The top of the north_star range (bench.py reports d_fft at m = 2^24): one d_fft round through the host-pointer
    C ABI -- QAP-style packing (zkg_pss_pack_vec layout 1), client fft1 per party, king pipeline -- then sampled output
    positions are unpacked and compared with X[k] = sum_j x_j w^(jk) evaluated by the CPU oracle (Horner, O(m) each).
    Nothing here leans on the device's own plain FFT."""
    z, capi, torch, ctx = env
    o = ol.oracle()
    lib = z.lib()
    l, log2m = 2, 24
    m = 1 << log2m
    mbyl = m // l
    rng = np.random.default_rng(24)
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(m)
    x = rng.integers(0, 2**64, size=(m, 4), dtype=np.uint64)
    x[:, 3] &= np.uint64((1 << 61) - 1)                               # any value below 2^253 is a valid Montgomery image
    shares = z.qap_pss_pack(x, pp, ol.rand_fr(rng, 64)[np.arange(mbyl * pp.t) % 64])     # 64 draws, tiled
    for p in range(pp.n):                                             # clients: dfft/mod.rs:121
        z.fft1_in_place(shares[p], pp, dom.group_gen())
    rnd = ol.rand_fr(rng, 64)[np.arange(mbyl * pp.t) % 64]
    out = z.king_fft2(shares, list(range(pp.n)), pp, dom.group_gen(), dom.element(0),
                      False, rnd)                                     # king: dfft/mod.rs:264-304, consecutive packing
    del shares
    ks = [0, 1, 2, m // 2 - 1, m // 2, m - 1] + [int(v) for v in rng.integers(0, m, size=6)]
    cols = sorted({k // l for k in ks})
    col_shares = np.ascontiguousarray(np.stack([np.stack([out[p][c] for p in range(pp.n)]) for c in cols])).reshape(-1, 4)
    secrets = pp.unpack(col_shares).reshape(len(cols), l, 4)
    from zksaas_b200.api import fr_image
    w = pyref_root(m)
    for k in ks:
        got = secrets[cols.index(k // l), k % l]
        pt = fr_image(pow(w, k, ol.pyref.R_MOD))
        exp = np.zeros(4, dtype=np.uint64)
        o.zko_fr_eval_poly(_p(x), m, _p(pt), _p(exp), 16)
        assert (got == exp).all(), k


def pyref_root(m):
    return ol.pyref.Radix2Domain(m).group_gen
