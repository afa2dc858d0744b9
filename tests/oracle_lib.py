"""ctypes access to the CPU oracle (oracle/libzkoracle.so) + image conversion helpers.

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use this.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref  # noqa: E402

_LIB = None
u64p = C.POINTER(C.c_uint64)


def _p(a):
    return a.ctypes.data_as(u64p)


def oracle():
    """Load (building if needed and a compiler is present) the C oracle."""
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "libzkoracle.so")
        src = os.path.join(ROOT, "oracle", "zkoracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        lib = C.CDLL(so)
        lib.zko_max_threads.restype = C.c_int
        for f in ("zko_g1_msm", "zko_g2_msm"):
            getattr(lib, f).argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_size_t, u64p, C.c_int, C.c_int]
            getattr(lib, f).restype = C.c_int
        for f in ("zko_g1_fixed_base", "zko_g2_fixed_base"):
            getattr(lib, f).argtypes = [u64p, C.c_size_t, C.c_void_p, C.c_size_t]
        for fld in ("fr", "fq"):
            for op in ("mul", "add", "sub"):
                getattr(lib, f"zko_{fld}_{op}").argtypes = [u64p, u64p, u64p, C.c_size_t]
            for op in ("inv", "to_mont", "from_mont"):
                getattr(lib, f"zko_{fld}_{op}").argtypes = [u64p, u64p, C.c_size_t]
        lib.zko_g1_sequence.argtypes = [u64p, u64p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.zko_fr_sum.argtypes = [u64p, C.c_size_t, u64p]
        lib.zko_fr_eval_poly.argtypes = [u64p, C.c_size_t, u64p, u64p, C.c_int]
        lib.zko_fq2_mul.argtypes = [u64p, u64p, u64p, C.c_size_t]
        lib.zko_fq2_sqr.argtypes = [u64p, u64p, C.c_size_t]
        lib.zko_fq2_inv.argtypes = [u64p, u64p, C.c_size_t]
        lib.zko_fr_root_of_unity.argtypes = [C.c_size_t, u64p]
        lib.zko_fr_fft.argtypes = [u64p, C.c_size_t, u64p, C.c_int]
        lib.zko_fr_distribute_powers.argtypes = [u64p, C.c_size_t, u64p]
        lib.zko_fr_rearrange.argtypes = [u64p, C.c_size_t]
        lib.zko_fft1_in_place.argtypes = [u64p, C.c_size_t, C.c_uint32, u64p]
        lib.zko_fft2_in_place.argtypes = [u64p, C.c_size_t, C.c_uint32, u64p]
        lib.zko_pss_pack_fr.argtypes = [C.c_uint32, u64p, u64p, u64p, C.c_size_t]
        lib.zko_pss_unpack_fr.argtypes = [C.c_uint32, u64p, u64p, C.c_size_t]
        lib.zko_pss_unpack2_fr.argtypes = [C.c_uint32, u64p, u64p, C.c_size_t]
        lib.zko_pss_lagrange_unpack_fr.argtypes = [C.c_uint32, u64p, C.POINTER(C.c_uint32), C.c_uint32, u64p, C.c_size_t]
        lib.zko_pss_lagrange_unpack_fr.restype = C.c_int
        for g in ("g1", "g2"):
            getattr(lib, f"zko_pss_pack_{g}").argtypes = [C.c_uint32, u64p, u64p, u64p]
            getattr(lib, f"zko_pss_unpack2_{g}").argtypes = [C.c_uint32, u64p, u64p]
            getattr(lib, f"zko_pss_unpack_{g}").argtypes = [C.c_uint32, u64p, u64p]
            getattr(lib, f"zko_{g}_add").argtypes = [u64p, u64p, u64p]
            getattr(lib, f"zko_{g}_mul").argtypes = [u64p, u64p, u64p]
            getattr(lib, f"zko_{g}_normalize").argtypes = [u64p]
            getattr(lib, f"zko_{g}_on_curve").argtypes = [C.c_void_p]
            getattr(lib, f"zko_{g}_on_curve").restype = C.c_int
        pp = C.POINTER(u64p)
        lib.zko_king_fft2.argtypes = [pp, C.POINTER(C.c_uint32), C.c_uint32, C.c_size_t, C.c_uint32, u64p, u64p,
                                      C.c_int, u64p, pp]
        lib.zko_king_fft2.restype = C.c_int
        lib.zko_deg_red_king.argtypes = [pp, C.POINTER(C.c_uint32), C.c_uint32, C.c_size_t, C.c_uint32, u64p, pp]
        lib.zko_deg_red_king.restype = C.c_int
        lib.zko_dpp_king.argtypes = [pp, C.POINTER(C.c_uint32), C.c_uint32, C.c_size_t, C.c_uint32, u64p, pp]
        lib.zko_dpp_king.restype = C.c_int
        _LIB = lib
    return _LIB


# ---------------------------------------------------------------------------------------------
# int <-> arkworks memory images
# ---------------------------------------------------------------------------------------------
def mont_np(vals, p):
    """list of ints -> (n,4) uint64 Montgomery images."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = pyref.to_mont_limbs(v, p)
    return out


def fr_np(vals):
    return mont_np(vals, pyref.R_MOD)


def fq_np(vals):
    return mont_np(vals, pyref.Q_MOD)


def np_ints(arr, p):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [pyref.from_mont_limbs(row, p) for row in arr]


def np_fr(arr):
    return np_ints(arr, pyref.R_MOD)


def g1_aff_np(points):
    out = np.zeros((len(points), 72), dtype=np.uint8)
    for i, P in enumerate(points):
        out[i] = np.frombuffer(pyref.g1_affine_image(P), dtype=np.uint8)
    return out


def g2_aff_np(points):
    out = np.zeros((len(points), 136), dtype=np.uint8)
    for i, P in enumerate(points):
        out[i] = np.frombuffer(pyref.g2_affine_image(P), dtype=np.uint8)
    return out


def g1_xyz_to_point(xyz):
    """normalised Jacobian image (12 u64) -> affine int tuple or None (identity)."""
    xyz = np.asarray(xyz, dtype=np.uint64).reshape(3, 4)
    z = pyref.from_mont_limbs(xyz[2], pyref.Q_MOD)
    if z == 0:
        return None
    assert z == 1, "result not normalised"
    return (pyref.from_mont_limbs(xyz[0], pyref.Q_MOD), pyref.from_mont_limbs(xyz[1], pyref.Q_MOD))


def g2_xyz_to_point(xyz):
    xyz = np.asarray(xyz, dtype=np.uint64).reshape(3, 2, 4)
    z = (pyref.from_mont_limbs(xyz[2, 0], pyref.Q_MOD), pyref.from_mont_limbs(xyz[2, 1], pyref.Q_MOD))
    if z == (0, 0):
        return None
    assert z == (1, 0), "result not normalised"
    f = lambda r: pyref.Fq2(pyref.from_mont_limbs(r[0], pyref.Q_MOD), pyref.from_mont_limbs(r[1], pyref.Q_MOD))
    return (f(xyz[0]), f(xyz[1]))


def g1_point_to_xyz(P):
    out = np.zeros((3, 4), dtype=np.uint64)
    if P is None:
        out[0] = pyref.to_mont_limbs(1, pyref.Q_MOD)
        out[1] = pyref.to_mont_limbs(1, pyref.Q_MOD)
    else:
        out[0] = pyref.to_mont_limbs(P[0], pyref.Q_MOD)
        out[1] = pyref.to_mont_limbs(P[1], pyref.Q_MOD)
        out[2] = pyref.to_mont_limbs(1, pyref.Q_MOD)
    return out.reshape(12)


def g2_point_to_xyz(P):
    out = np.zeros((3, 2, 4), dtype=np.uint64)
    one = pyref.to_mont_limbs(1, pyref.Q_MOD)
    if P is None:
        out[0, 0] = one
        out[1, 0] = one
    else:
        out[0, 0] = pyref.to_mont_limbs(P[0].c0, pyref.Q_MOD)
        out[0, 1] = pyref.to_mont_limbs(P[0].c1, pyref.Q_MOD)
        out[1, 0] = pyref.to_mont_limbs(P[1].c0, pyref.Q_MOD)
        out[1, 1] = pyref.to_mont_limbs(P[1].c1, pyref.Q_MOD)
        out[2, 0] = one
    return out.reshape(24)


def ptr_array(arrs):
    """list of (k,4) uint64 arrays -> C array of uint64_t* (keeps references alive via return)."""
    T = u64p * len(arrs)
    return T(*[_p(a) for a in arrs])


# ---------------------------------------------------------------------------------------------
# convenience wrappers used by several tests
# ---------------------------------------------------------------------------------------------
def o_g1_msm(bases_img, scalars, threads=1, c_override=0):
    lib = oracle()
    out = np.zeros(12, dtype=np.uint64)
    bases_img = np.ascontiguousarray(bases_img)
    scalars = np.ascontiguousarray(scalars)
    n = scalars.reshape(-1, 4).shape[0]
    lib.zko_g1_msm(bases_img.ctypes.data, bases_img.strides[0] if bases_img.ndim == 2 else 72, _p(scalars), n,
                   _p(out), threads, c_override)
    return out


def o_g2_msm(bases_img, scalars, threads=1, c_override=0):
    lib = oracle()
    out = np.zeros(24, dtype=np.uint64)
    bases_img = np.ascontiguousarray(bases_img)
    scalars = np.ascontiguousarray(scalars)
    n = scalars.reshape(-1, 4).shape[0]
    lib.zko_g2_msm(bases_img.ctypes.data, bases_img.strides[0] if bases_img.ndim == 2 else 136, _p(scalars), n,
                   _p(out), threads, c_override)
    return out


def rand_fr(rng, n):
    """uniform Fr elements as Montgomery images, via the arkworks Fp::rand recipe
    (sample 256 bits, clear the top 2, reject >= r).  Montgomery image of a uniform element is
    uniform, so the sampled canonical value is used directly as the image."""
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    mod = np.array(pyref.to_mont_limbs(0, pyref.R_MOD), dtype=np.uint64)  # placeholder
    r_limbs = [(pyref.R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    while filled < n:
        k = n - filled
        cand = rng.integers(0, 2**64, size=(k + 8, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 62) - 1)
        # lexicographic compare with r from the top limb
        lt = np.zeros(len(cand), dtype=bool)
        eq = np.ones(len(cand), dtype=bool)
        for i in (3, 2, 1, 0):
            lt |= eq & (cand[:, i] < np.uint64(r_limbs[i]))
            eq &= cand[:, i] == np.uint64(r_limbs[i])
        good = cand[lt][:k]
        out[filled:filled + len(good)] = good
        filled += len(good)
    return out
