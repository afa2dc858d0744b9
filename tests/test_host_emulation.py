"""CPU: the product's device arithmetic headers (fp.cuh, ec.cuh, msm_common.cuh) compiled for the
HOST with the PTX carry-chain primitives emulated (ZKG_HOST_EMU), checked against the oracle.
This exercises the exact limb schedules and group formulas the CUDA kernels run, without a GPU."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import _p, pyref, u64p

R, Q = pyref.R_MOD, pyref.Q_MOD
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "host_emu", "emu.cpp")
    out_dir = os.path.join(HERE, "host_emu", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libemu.so")
    csrc = os.path.join(HERE, "..", "zk-saas_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("fp.cuh", "ec.cuh", "msm_common.cuh", "bn254_consts.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    return lib


def raw(vals):
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = [(v >> (64 * k)) & (2**64 - 1) for k in range(4)]
    return out


def unraw(arr):
    return [sum(int(x) << (64 * k) for k, x in enumerate(r)) for r in np.asarray(arr).reshape(-1, 4)]


@pytest.mark.parametrize("fld,p", [("fr", R), ("fq", Q)])
def test_montgomery_limb_schedule(emu, fld, p):
    rng = random.Random(7)
    edge = [0, 1, 2, p - 1, p - 2, 1 << 253, (1 << 32) - 1, (1 << 64) - 1, 1 << 224, p >> 1, 0xffffffff00000000ffffffff]
    a = [rng.randrange(p) for _ in range(20000)] + [x for x in edge for _ in edge]
    b = [rng.randrange(p) for _ in range(20000)] + [y for _ in edge for y in edge]
    A, B = raw(a), raw(b)
    O = np.zeros_like(A)
    rinv = pow(1 << 256, -1, p)
    for op, f in (("mul", lambda x, y: x * y * rinv % p), ("mul_r29", lambda x, y: x * y * rinv % p), ("add", lambda x, y: (x + y) % p), ("sub", lambda x, y: (x - y) % p)):
        fn = getattr(emu, f"emu_{fld}_{op}")
        fn.argtypes = [u64p, u64p, u64p, C.c_size_t]
        fn(_p(A), _p(B), _p(O), len(a))
        assert unraw(O) == [f(x, y) for x, y in zip(a, b)], op
    fn = getattr(emu, f"emu_{fld}_neg"); fn.argtypes = [u64p, u64p, C.c_size_t]
    fn(_p(A), _p(O), len(a))
    assert unraw(O) == [(-x) % p for x in a]
    # dedicated squaring (triangular product + word-by-word reduction): random values, the edge set, and
    # values with saturated limbs / top bits of every limb set (the doubling folded into the multiplicand)
    sq = a + [p - 1 - (1 << (32 * k)) for k in range(8)] + [((1 << 254) - 1) % p, int("7fffffff" * 8, 16) % p,
              int("80000000" * 8, 16) % p, int("ffffffff" * 7, 16), ((1 << 253) | (1 << 31) | 1)]
    SA = raw(sq); SO = np.zeros_like(SA)
    fn = getattr(emu, f"emu_{fld}_sqr"); fn.argtypes = [u64p, u64p, C.c_size_t]
    fn(_p(SA), _p(SO), len(sq))
    assert unraw(SO) == [x * x * rinv % p for x in sq]
    # binary-Euclid inversion (production) and the Fermat ladder, on random values and the edge set
    inv_in = a[:3000] + edge + [3, 4, p - 3, (p + 1) // 2, 1 << 255 & (p - 1), pow(1 << 256, 1, p), pow(1 << 256, 2, p)]
    IA = raw(inv_in)
    IO = np.zeros_like(IA)
    expect = [pow(x * rinv % p, -1, p) * (1 << 256) % p if x else 0 for x in inv_in]
    fn = getattr(emu, f"emu_{fld}_inv"); fn.argtypes = [u64p, u64p, C.c_size_t]
    fn(_p(IA), _p(IO), len(inv_in))
    assert unraw(IO) == expect
    fn = getattr(emu, f"emu_{fld}_inv_fermat"); fn.argtypes = [u64p, u64p, C.c_size_t]
    fn(_p(IA[:60]), _p(IO), 60)
    assert unraw(IO[:60]) == expect[:60]
    fn = getattr(emu, f"emu_{fld}_from_mont"); fn.argtypes = [u64p, u64p, C.c_size_t]
    fn(_p(A), _p(O), len(a))
    assert unraw(O) == [x * rinv % p for x in a]


@pytest.mark.parametrize("fld,p", [("fr", R), ("fq", Q)])
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_montgomery_dot_product(emu, fld, p, k):
    """fp_dot<K>: one reduction row per K accumulated product rows; must equal the sum of K products,
    including the extremes that drive the running value to its (K+1) p bound."""
    rng = random.Random(31 * k + len(fld))
    rinv = pow(1 << 256, -1, p)
    n = 4000
    a = [rng.randrange(p) for _ in range(n * k)]
    b = [rng.randrange(p) for _ in range(n * k)]
    ext = [p - 1, p - 2, (1 << 253) + 12345, (1 << 32) - 1, 0, 1, p >> 1, (p - 1) ^ ((1 << 192) - 1)]
    for x in ext:
        for y in ext:
            a += [x] * k
            b += [y] * k
    # mixed extremes inside one dot product
    for _ in range(200):
        a += [rng.choice(ext) for _ in range(k)]
        b += [rng.choice(ext) for _ in range(k)]
    cnt = len(a) // k
    A, B = raw(a), raw(b)
    O = np.zeros((cnt, 4), dtype=np.uint64)
    fn = getattr(emu, f"emu_{fld}_dot")
    fn.argtypes = [u64p, u64p, u64p, C.c_size_t, C.c_int]
    fn(_p(A), _p(B), _p(O), cnt, k)
    exp = [sum(a[i * k + j] * b[i * k + j] for j in range(k)) * rinv % p for i in range(cnt)]
    assert unraw(O) == exp


def test_fq2_vs_oracle(emu):
    o = ol.oracle()
    rng = np.random.default_rng(5)
    n = 500
    a = np.concatenate([ol.rand_fr(rng, 2 * n)]).reshape(n, 8) % np.uint64(2**64 - 1)
    # reduce into Fq by construction: reuse canonical Fr samples (r < q) as Fq images
    b = ol.rand_fr(rng, 2 * n).reshape(n, 8)
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    got, exp = np.zeros_like(a), np.zeros_like(a)
    for name, args in (("mul", (a, b)), ("sqr", (a,)), ("inv", (a,))):
        fe = getattr(emu, f"emu_fq2_{name}"); fo = getattr(o, f"zko_fq2_{name}")
        fe.argtypes = [u64p] * (len(args) + 1) + [C.c_size_t]
        fe(*[_p(x) for x in args], _p(got), n)
        fo(*[_p(x) for x in args], _p(exp), n)
        assert (got == exp).all(), name


def test_signed_digits_reconstruct(emu):
    rng = random.Random(9)
    emu.emu_digits.argtypes = [u64p, C.c_int, C.POINTER(C.c_int32)]
    emu.emu_num_windows.restype = C.c_int
    for c in (5, 8, 13, 16, 17, 20):
        W = emu.emu_num_windows(c)
        assert W * c >= 255
        half = 1 << (c - 1)
        for s in [0, 1, R - 1, (1 << 253), half, half - 1, (1 << c) - 1] + [rng.randrange(R) for _ in range(300)]:
            out = (C.c_int32 * W)()
            emu.emu_digits(_p(raw([s])), c, out)
            d = list(out)
            assert sum(v << (c * i) for i, v in enumerate(d)) == s
            assert all(-half <= v <= half for v in d)


def _packed_g1(aff72):
    return np.ascontiguousarray(aff72[:, :64]).view(np.uint64).reshape(-1, 8)


@pytest.mark.parametrize("n,c", [(1, 5), (40, 5), (200, 7), (64, 3)])
def test_emulated_msm_g1_vs_oracle(emu, n, c):
    o = ol.oracle()
    rng = random.Random(n * 31 + c)
    dl = [rng.randrange(R) for _ in range(n)]
    sc = [rng.randrange(R) for _ in range(n)]
    if n >= 40:
        dl[1] = dl[0]; sc[1] = sc[0]; dl[3] = R - dl[2]; sc[3] = sc[2]; sc[4] = 0; sc[5] = R - 1; dl[6] = 0
    bases = np.zeros((n, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(ol.fr_np(dl)), n, bases.ctypes.data, 72)
    packed = _packed_g1(bases)
    packed[[i for i in range(n) if bases[i, 64]]] = 0          # infinity -> (0, 0)
    S = ol.fr_np(sc)
    out = np.zeros(12, dtype=np.uint64)
    emu.emu_msm_g1.argtypes = [u64p, u64p, C.c_size_t, C.c_int, u64p]
    emu.emu_msm_g1(_p(packed), _p(S), n, c, _p(out))
    assert (out == ol.o_g1_msm(bases, S)).all()


def test_emulated_msm_g2_vs_oracle(emu):
    o = ol.oracle()
    rng = random.Random(77)
    n = 30
    dl = [rng.randrange(R) for _ in range(n)]
    sc = [rng.randrange(R) for _ in range(n)]
    dl[1] = dl[0]; sc[1] = sc[0]; sc[2] = 0; dl[3] = 0
    bases = np.zeros((n, 136), dtype=np.uint8)
    o.zko_g2_fixed_base(_p(ol.fr_np(dl)), n, bases.ctypes.data, 136)
    packed = np.ascontiguousarray(bases[:, :128]).view(np.uint64).reshape(-1, 16)
    packed[3] = 0
    S = ol.fr_np(sc)
    out = np.zeros(24, dtype=np.uint64)
    emu.emu_msm_g2.argtypes = [u64p, u64p, C.c_size_t, C.c_int, u64p]
    emu.emu_msm_g2(_p(packed), _p(S), n, 4, _p(out))
    assert (out == ol.o_g2_msm(bases, S)).all()
