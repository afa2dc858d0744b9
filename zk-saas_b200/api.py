"""Host-side mirror of the reference's Rust interface for the hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference so that the parity tests read
like the reference's own tests (paths relative to the reference root):

    PackedSharingParams.{new,pack,det_pack,unpack,unpack2,unpack_missing_shares}  secret-sharing/src/pss.rs:39-221
    pack_vec, transpose                                    dist-primitives/src/utils/pack.rs:8-35
    fft1_in_place / fft2_in_place / fft_in_place_rearrange dist-primitives/src/dfft/mod.rs:178-237,322-335
    d_fft / d_ifft / FftMask                               dist-primitives/src/dfft/mod.rs:16-175
    d_msm / MsmMask                                        dist-primitives/src/dmsm/mod.rs:10-102
    deg_red / DegRedMask                                   dist-primitives/src/utils/deg_red.rs:14-126
    msm_g1 / msm_g2  (= G::msm)                            call site dist-primitives/src/dmsm/mod.rs:73

Data are arkworks memory images held in numpy arrays: Fr vectors are (k, 4) uint64 Montgomery
limbs, G1/G2 affine bases are (k, 72) / (k, 136) uint8, group results are normalised Jacobian images
of 12 / 24 uint64.  All arithmetic on those arrays happens in libzksaas_gpu.so; the big-integer
code below only derives *parameters* (roots of unity, domain constants).  mpc-net is out of scope:
`LocalTestNet` here is an in-process stand-in for `LocalTestNet::simulate_network_round`
(mpc-net/src/multi.rs:301-328) that hands the king every party's message directly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, u64p

# ---- parameters (big-int, host) ---------------------------------------------------------------
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_GENERATOR = 5
_TWO_ADICITY = 28
_TWO_ADIC_ROOT = pow(FR_GENERATOR, (R_MOD - 1) >> _TWO_ADICITY, R_MOD)
_MASK64 = (1 << 64) - 1


def fr_image(v: int) -> np.ndarray:
    """int -> (4,) uint64 Montgomery image of an Fr element."""
    m = (v % R_MOD) * (1 << 256) % R_MOD
    return np.array([(m >> (64 * i)) & _MASK64 for i in range(4)], dtype=np.uint64)


def fr_value(img) -> int:
    v = sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(img, dtype=np.uint64).reshape(4)))
    return v * pow(1 << 256, -1, R_MOD) % R_MOD


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _require(cond, msg):
    """Argument checks that must survive `python -O` (the C side reads the buffers it is told about)."""
    if not cond:
        raise ValueError(msg)


def _fr_vec(a, name="vector"):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError(f"{name}: expected (k, 4) uint64 Fr images, got shape {a.shape}")
    return a


class Radix2EvaluationDomain:
    """ark-poly Radix2EvaluationDomain<Fr>: the constants d_fft / d_ifft take from `dom`."""

    def __init__(self, num_coeffs: int):
        size = 1
        while size < num_coeffs:
            size <<= 1
        log = size.bit_length() - 1
        if log > _TWO_ADICITY:
            raise ValueError("domain larger than 2^28")
        self.size_ = size
        self.log_size_of_group = log
        self._gen = pow(_TWO_ADIC_ROOT, 1 << (_TWO_ADICITY - log), R_MOD)

    @classmethod
    def new(cls, num_coeffs: int):
        return cls(num_coeffs)

    def size(self):
        return self.size_

    def group_gen(self):
        return fr_image(self._gen)

    def group_gen_inv(self):
        return fr_image(pow(self._gen, -1, R_MOD))

    def size_inv(self):
        return fr_image(pow(self.size_, -1, R_MOD))

    def element(self, i: int):
        return fr_image(pow(self._gen, i, R_MOD))

    # plain transforms (device): Radix2EvaluationDomain::{fft,ifft}, optional coset offset image
    def fft(self, v, offset=None, device=0):
        out = _fr_vec(v).copy()
        if out.shape[0] != self.size_:
            raise ValueError("fft: vector length must equal the domain size")
        check(lib().zkg_fr_fft_bn254(device, _ptr(out), out.shape[0], _ptr(offset), 0))
        return out

    def ifft(self, v, offset=None, device=0):
        out = _fr_vec(v).copy()
        if out.shape[0] != self.size_:
            raise ValueError("ifft: vector length must equal the domain size")
        check(lib().zkg_fr_fft_bn254(device, _ptr(out), out.shape[0], _ptr(offset), 1))
        return out


# ---- secret-sharing/src/pss.rs ----------------------------------------------------------------
class PackedSharingParams:
    """pss.rs:19-66: n = 4l parties, t = l."""

    def __init__(self, l: int, device: int = 0):
        if l not in (2, 4, 8):
            raise ValueError("packing factor l must be 2, 4 or 8")
        self.l, self.t, self.n, self.device = l, l, 4 * l, device

    @classmethod
    def new(cls, l: int, device: int = 0):
        return cls(l, device)

    def pack(self, secrets, rand_points):
        """pss.rs:90-122; secrets (cols*l, 4), rand_points (cols*t, 4) -> shares (cols*n, 4).
        The t random points per column are an input (the host RNG stays with the caller)."""
        secrets, rand_points = _fr_vec(secrets, "secrets"), _fr_vec(rand_points, "rand_points")
        _require(secrets.shape[0] % self.l == 0, "Secrets length mismatch")
        cols = secrets.shape[0] // self.l
        _require(rand_points.shape[0] == cols * self.t, "rand_points length mismatch")
        out = np.empty((cols * self.n, 4), dtype=np.uint64)
        check(lib().zkg_pss_pack_bn254_fr(self.device, self.l, _ptr(secrets), _ptr(rand_points), _ptr(out), cols))
        return out

    def det_pack(self, secrets):
        """pss.rs:69-87."""
        secrets = _fr_vec(secrets, "secrets")
        _require(secrets.shape[0] % self.l == 0, "Secrets length mismatch")
        cols = secrets.shape[0] // self.l
        out = np.empty((cols * self.n, 4), dtype=np.uint64)
        check(lib().zkg_pss_pack_bn254_fr(self.device, self.l, _ptr(secrets), None, _ptr(out), cols))
        return out

    def unpack(self, shares):
        """pss.rs:125-138."""
        shares = _fr_vec(shares, "shares")
        _require(shares.shape[0] % self.n == 0, "shares length is not a multiple of n")
        cols = shares.shape[0] // self.n
        out = np.empty((cols * self.l, 4), dtype=np.uint64)
        check(lib().zkg_pss_unpack_bn254_fr(self.device, self.l, _ptr(shares), _ptr(out), cols))
        return out

    def unpack2(self, shares):
        """pss.rs:141-166."""
        shares = _fr_vec(shares, "shares")
        _require(shares.shape[0] % self.n == 0, "shares length is not a multiple of n")
        cols = shares.shape[0] // self.n
        out = np.empty((cols * self.l, 4), dtype=np.uint64)
        check(lib().zkg_pss_unpack2_bn254_fr(self.device, self.l, _ptr(shares), _ptr(out), cols))
        return out


def _unpack2_matrix(pp):
    """l x n matrix of Fr images with secrets = M * shares: unpack2 applied to the n unit vectors."""
    eye = np.zeros((pp.n * pp.n, 4), dtype=np.uint64)
    one = fr_image(1)
    for j in range(pp.n):
        eye[j * pp.n + j] = one
    cols = pp.unpack2(eye).reshape(pp.n, pp.l, 4)            # cols[j][i] = M[i][j]
    return [[cols[j, i] for j in range(pp.n)] for i in range(pp.l)]


PackedSharingParams.unpack2_matrix = _unpack2_matrix


def transpose(matrix):
    """utils/pack.rs:22-35 for a list of equal-length (k,4) vectors -> list of (len,4) vectors."""
    _require(len(matrix) > 0, "transpose of an empty matrix")
    a = np.stack([_fr_vec(r) for r in matrix])            # rows x cols x 4
    return [np.ascontiguousarray(a[:, c, :]) for c in range(a.shape[1])]


def pack_vec(secrets, pp: PackedSharingParams, rand_points):
    """utils/pack.rs:8-20: chunk by l and pack each chunk -> list of per-column share vectors."""
    shares = pp.pack(secrets, rand_points)
    return [shares[c * pp.n:(c + 1) * pp.n] for c in range(shares.shape[0] // pp.n)]


def _pack_vec_by_party(x, pp: "PackedSharingParams", rand_points, layout):
    x, rand_points = _fr_vec(x, "x"), _fr_vec(rand_points, "rand_points")
    cols = (x.shape[0] + pp.l - 1) // pp.l
    _require(rand_points.shape[0] == cols * pp.t, "rand_points length mismatch")
    outs = [np.zeros((cols, 4), dtype=np.uint64) for _ in range(pp.n)]
    arr = (C.POINTER(C.c_uint64) * pp.n)(*[o.ctypes.data_as(C.POINTER(C.c_uint64)) for o in outs])
    check(lib().zkg_pss_pack_vec_bn254_fr(pp.device, pp.l, layout, _ptr(x), x.shape[0], _ptr(rand_points), arr))
    return outs


def pack_from_witness(pp: "PackedSharingParams", full_assignment, rand_points):
    """groth16/examples/sha256.rs:131-156: l-chunks of the assignment (the last one zero-padded) packed,
    returned as the n parties' share vectors.  rand_points: ceil(len/l)*t host-RNG values."""
    return _pack_vec_by_party(full_assignment, pp, rand_points, 0)


def qap_pss_pack(x, pp: "PackedSharingParams", rand_points):
    """The `pack` closure of QAP::pss (groth16/src/qap.rs:99-112): fft_in_place_rearrange(x), column i packs
    x[i], x[i + m/l], ...; returned party-major (what the closing loop at :118-133 builds)."""
    return _pack_vec_by_party(x, pp, rand_points, 1)


def qap_pss(a, b, c, pp: "PackedSharingParams", rand_a, rand_b, rand_c):
    """QAP::pss (groth16/src/qap.rs:92-134): per party the (a, b, c) share vectors of a PackedQAPShare."""
    pa, pb, pc = qap_pss_pack(a, pp, rand_a), qap_pss_pack(b, pp, rand_b), qap_pss_pack(c, pp, rand_c)
    return [(pa[i], pb[i], pc[i]) for i in range(pp.n)]


# ---- G::msm -------------------------------------------------------------------------------------
class MsmLengthMismatch(ValueError):
    """`G::msm` returns Err(min(bases.len(), scalars.len())) on a length mismatch."""

    def __init__(self, min_len):
        super().__init__(f"msm length mismatch, min len {min_len}")
        self.min_len = min_len


def _msm(fn, words, stride, bases, scalars, device):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    scalars = _fr_vec(scalars, "scalars")
    if bases.ndim != 2 or bases.shape[1] != stride:
        raise ValueError(f"bases: expected (k, {stride}) uint8 affine images")
    out = np.zeros(words, dtype=np.uint64)
    rc = fn(device, _ptr(bases), stride, bases.shape[0], _ptr(scalars), scalars.shape[0], _ptr(out))
    if rc == capi.ZKG_ERR_LEN_MISMATCH:
        raise MsmLengthMismatch(min(bases.shape[0], scalars.shape[0]))
    check(rc)
    return out


def msm_g1(bases, scalars, device=0):
    """ark_bn254::G1Projective::msm(bases, scalars) (dmsm/mod.rs:73) -> normalised Jacobian image."""
    return _msm(lib().zkg_msm_bn254_g1, 12, 72, bases, scalars, device)


def msm_g2(bases, scalars, device=0):
    return _msm(lib().zkg_msm_bn254_g2, 24, 136, bases, scalars, device)


def crs_det_pack(bases, pp: "PackedSharingParams", g2=False):
    """pack_from_arkworks_proving_key's inner step (groth16/src/proving_key.rs:72-104): det_pack every l-chunk of a
    CRS query over group elements; returns the n parties' affine share vectors ((chunks, 72|136) uint8 each)."""
    stride = 136 if g2 else 72
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    _require(bases.ndim == 2 and bases.shape[1] == stride and bases.shape[0] % pp.l == 0,
             f"bases: expected (chunks*l, {stride}) uint8 affine images")
    chunks = bases.shape[0] // pp.l
    outs = [np.zeros((chunks, stride), dtype=np.uint8) for _ in range(pp.n)]
    arr = (C.c_void_p * pp.n)(*[o.ctypes.data for o in outs])
    check(lib().zkg_crs_det_pack_bn254(pp.device, 2 if g2 else 1, _ptr(bases), stride, bases.shape[0], pp.l, arr, stride))
    return outs


def _xyz_to_affine_images(points, g2=False):
    """normalised Jacobian images -> arkworks Affine images (for feeding results back as bases)."""
    w = 8 if g2 else 4
    pts = np.asarray(points, dtype=np.uint64).reshape(-1, 3, w)
    out = np.zeros((pts.shape[0], 136 if g2 else 72), dtype=np.uint8)
    for i, p in enumerate(pts):
        if not p[2].any():
            out[i, 2 * w * 8] = 1
        else:
            out[i, : 2 * w * 8] = np.frombuffer(p[:2].tobytes(), dtype=np.uint8)
    return out


_ONE = None


def _one_img():
    global _ONE
    if _ONE is None:
        _ONE = fr_image(1)
    return _ONE


def group_lincomb(points, coeffs, g2=False, device=0):
    """sum_i coeffs[i] * points[i] over normalised Jacobian images (a tiny MSM on the device)."""
    aff = _xyz_to_affine_images(points, g2)
    return (msm_g2 if g2 else msm_g1)(aff, np.stack(coeffs), device)


def group_add(a, b, g2=False, device=0):
    return group_lincomb([a, b], [_one_img(), _one_img()], g2, device)


def pss_unpack2_group(pp: "PackedSharingParams", shares, parties=None, g2=False, want_unpacked=True):
    """pp.unpack_missing_shares(&shares, &parties) over group elements + the sum of the results
    (dmsm/mod.rs:85-86; sha256.rs:375-377).  shares: Projective images (any Z), one per entry of `parties`.
    Returns (list of the l unpacked points or None, their sum), normalised Jacobian images."""
    w = 24 if g2 else 12
    pts = np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.uint64).reshape(w) for x in shares]))
    parties = list(range(len(shares))) if parties is None else list(parties)
    _require(len(parties) == pts.shape[0], "one share per received party expected")
    par = (C.c_uint32 * len(parties))(*parties)
    rows = np.zeros((pp.l, w), dtype=np.uint64) if want_unpacked else None
    total = np.zeros(w, dtype=np.uint64)
    fn = lib().zkg_pss_unpack2_bn254_g2 if g2 else lib().zkg_pss_unpack2_bn254_g1
    check(fn(pp.device, pp.l, _ptr(pts), par, len(parties), _ptr(rows), _ptr(total)))
    return (list(rows) if want_unpacked else None), total


def group_to_wire(points, g2=False, device=0):
    """ark-serialize compressed form of Projective images (ser_net.rs:25): (k, 32 | 64) uint8."""
    w, nb = (24, 64) if g2 else (12, 32)
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.uint64).reshape(-1, w))
    out = np.zeros((pts.shape[0], nb), dtype=np.uint8)
    fn = lib().zkg_g2_to_wire_bn254 if g2 else lib().zkg_g1_to_wire_bn254
    check(fn(device, _ptr(pts), _ptr(out), pts.shape[0]))
    return out


def group_from_wire(wire, g2=False, device=0):
    """deserialize_compressed (ser_net.rs:40,119; Validate::Yes) -> normalised Jacobian images; raises ZkgError
    (ZKG_ERR_BAD_ARG) on an invalid encoding."""
    w, nb = (24, 64) if g2 else (12, 32)
    wire = np.ascontiguousarray(wire, dtype=np.uint8)
    _require(wire.ndim == 2 and wire.shape[1] == nb, f"wire: (k, {nb}) uint8 expected")
    out = np.zeros((wire.shape[0], w), dtype=np.uint64)
    fn = lib().zkg_g2_from_wire_bn254 if g2 else lib().zkg_g1_from_wire_bn254
    check(fn(device, _ptr(wire), _ptr(out), wire.shape[0]))
    return out


def qap_h(a, b, c, mask_a=None, mask_b=None, mask_c=None, factor=None, device=0):
    """h = (a + mask_a)(b + mask_b) - (c + mask_c) [* factor] on share vectors: ext_wit.rs:173-177 / :82-86 fused with
    the out-mask additions of dfft/mod.rs:313-317."""
    a, b, c = _fr_vec(a, "a"), _fr_vec(b, "b"), _fr_vec(c, "c")
    _require(a.shape == b.shape == c.shape, "a, b, c of different lengths")
    ms = [None if m is None else _fr_vec(m, "mask") for m in (mask_a, mask_b, mask_c)]
    _require(all(m is None or m.shape == a.shape for m in ms), "mask length differs from the share vectors")
    out = np.empty_like(a)
    check(lib().zkg_qap_h_bn254(device, _ptr(a), _ptr(b), _ptr(c), _ptr(ms[0]), _ptr(ms[1]), _ptr(ms[2]), _ptr(factor), _ptr(out),
                                a.shape[0]))
    return out


# ---- dist-primitives/src/dfft ---------------------------------------------------------------------
def fft1_in_place(px, pp: PackedSharingParams, gen, pre_scale=None, in_mask=None, device=None):
    """dfft/mod.rs:178-208, in place on the (m/l, 4) share vector.  pre_scale / in_mask are the fused
    forms of dfft/mod.rs:159 and :254-258."""
    _require(px.dtype == np.uint64 and px.flags.c_contiguous and px.ndim == 2 and px.shape[1] == 4, "px: (m/l, 4) uint64 expected")
    dev = pp.device if device is None else device
    if in_mask is not None:
        in_mask = _fr_vec(in_mask, "in_mask")
        _require(in_mask.shape == px.shape, "in_mask length differs from the share vector")
    check(lib().zkg_fft1_bn254(dev, _ptr(px), px.shape[0], pp.l, _ptr(gen), _ptr(pre_scale), _ptr(in_mask)))
    return px


def fft2_in_place(s1, pp: PackedSharingParams, gen):
    """dfft/mod.rs:210-237."""
    _require(s1.dtype == np.uint64 and s1.flags.c_contiguous and s1.ndim == 2 and s1.shape[1] == 4, "(k, 4) uint64 expected")
    check(lib().zkg_fft2_bn254(pp.device, _ptr(s1), s1.shape[0], pp.l, _ptr(gen)))
    return s1


def fft_in_place_rearrange(data, device=0):
    """dfft/mod.rs:322-335."""
    _require(data.dtype == np.uint64 and data.flags.c_contiguous and data.ndim == 2 and data.shape[1] == 4, "(k, 4) uint64 expected")
    check(lib().zkg_bitrev_bn254(device, _ptr(data), data.shape[0]))
    return data


def distribute_powers(v, g, device=0):
    """Radix2EvaluationDomain::distribute_powers (dfft/mod.rs:49,279)."""
    _require(v.dtype == np.uint64 and v.flags.c_contiguous and v.ndim == 2 and v.shape[1] == 4, "(k, 4) uint64 expected")
    check(lib().zkg_distribute_powers_bn254(device, _ptr(v), v.shape[0], _ptr(g)))
    return v


def _ptr_array(arrs):
    T = u64p * len(arrs)
    return T(*[a.ctypes.data_as(u64p) for a in arrs])


def _share_vectors(shares, parties, pp):
    """The vectors the king received: one per entry of `parties`, all of the same length."""
    shares = [_fr_vec(s, "shares") for s in shares]
    _require(len(shares) >= 1 and len(shares) == len(parties), "one share vector per received party expected")
    _require(len(shares) <= pp.n, "more share vectors than parties")
    length = shares[0].shape[0]
    _require(all(s.shape[0] == length for s in shares), "share vectors of different lengths")
    return shares, length


def king_fft2(shares, parties, pp: PackedSharingParams, gen, g, rearrange, rand_points):
    """King closure of fft2_with_rearrange, dfft/mod.rs:264-304.  shares[r] is the vector received
    from parties[r]; returns the n per-party output vectors."""
    shares, mbyl = _share_vectors(shares, parties, pp)
    rand_points = _fr_vec(rand_points, "rand_points")
    _require(rand_points.shape[0] == mbyl * pp.t, "rand_points: m/l * t draws expected")
    outs = [np.empty((mbyl, 4), dtype=np.uint64) for _ in range(pp.n)]
    par = (C.c_uint32 * len(parties))(*parties)
    check(lib().zkg_king_fft2_bn254(pp.device, _ptr_array(shares), par, len(shares), mbyl, pp.l, _ptr(gen), _ptr(g),
                                    1 if rearrange else 0, _ptr(rand_points), _ptr_array(outs)))
    return outs


def deg_red_king(shares, parties, pp: PackedSharingParams, rand_points):
    """King closure of deg_red, utils/deg_red.rs:103-111."""
    shares, cols = _share_vectors(shares, parties, pp)
    rand_points = _fr_vec(rand_points, "rand_points")
    _require(rand_points.shape[0] == cols * pp.t, "rand_points: cols * t draws expected")
    outs = [np.empty((cols, 4), dtype=np.uint64) for _ in range(pp.n)]
    par = (C.c_uint32 * len(parties))(*parties)
    check(lib().zkg_deg_red_king_bn254(pp.device, _ptr_array(shares), par, len(shares), cols, pp.l,
                                       _ptr(rand_points), _ptr_array(outs)))
    return outs


def dpp_king(shares, parties, pp: PackedSharingParams, rand_points):
    """King closure of d_pp, dpp/mod.rs:41-76.  shares[r]: (2*cols, 4) = num shares then den shares."""
    shares, two_cols = _share_vectors(shares, parties, pp)
    _require(two_cols % 2 == 0, "d_pp: each party sends num shares followed by den shares")
    cols = two_cols // 2
    rand_points = _fr_vec(rand_points, "rand_points")
    _require(rand_points.shape[0] == cols * pp.t, "rand_points: cols * t draws expected")
    outs = [np.empty((cols, 4), dtype=np.uint64) for _ in range(pp.n)]
    par = (C.c_uint32 * len(parties))(*parties)
    check(lib().zkg_dpp_king_bn254(pp.device, _ptr_array(shares), par, len(shares), cols, pp.l,
                                   _ptr(rand_points), _ptr_array(outs)))
    return outs


def _field_op(op, a, b, field=0, device=0):
    a, b = _fr_vec(a), _fr_vec(b)
    _require(a.shape == b.shape, "element-wise operands of different lengths")
    out = np.empty_like(a)
    check(lib().zkg_field_op(device, field, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
    return out


def fr_mul(a, b, device=0):
    return _field_op(0, a, b, 0, device)


def fr_add(a, b, device=0):
    return _field_op(1, a, b, 0, device)


def fr_sub(a, b, device=0):
    return _field_op(2, a, b, 0, device)


# ---- masks -----------------------------------------------------------------------------------------
class FftMask:
    """dfft/mod.rs:16-95 (one party's share of the masks)."""

    def __init__(self, in_mask, out_mask):
        self.in_mask, self.out_mask = in_mask, out_mask

    @staticmethod
    def zero(mbyl):
        return FftMask(np.zeros((mbyl, 4), dtype=np.uint64), np.zeros((mbyl, 4), dtype=np.uint64))

    @staticmethod
    def sample(rearrange, g, gen, m, pp: PackedSharingParams, mask_values, rand_in, rand_out):
        """dfft/mod.rs:30-85 with the random draws passed in: mask_values (m,4), rand_in / rand_out
        (m/l*t, 4) packing randomness.  Returns the n parties' FftMask shares."""
        mask_values, rand_in, rand_out = _fr_vec(mask_values, "mask_values"), _fr_vec(rand_in, "rand_in"), _fr_vec(rand_out, "rand_out")
        _require(mask_values.shape[0] == m and m % pp.l == 0, "mask_values: m draws expected")
        mbyl = m // pp.l
        _require(rand_in.shape[0] == mbyl * pp.t and rand_out.shape[0] == mbyl * pp.t, "rand_in / rand_out: m/l * t draws expected")
        in_shares = [np.empty((mbyl, 4), dtype=np.uint64) for _ in range(pp.n)]
        out_shares = [np.empty((mbyl, 4), dtype=np.uint64) for _ in range(pp.n)]
        check(lib().zkg_fft_mask_sample_bn254(pp.device, 1 if rearrange else 0, _ptr(g), _ptr(gen), m, pp.l, _ptr(mask_values),
                                              _ptr(rand_in), _ptr(rand_out), _ptr_array(in_shares), _ptr_array(out_shares)))
        return [FftMask(i, o) for i, o in zip(in_shares, out_shares)]


class MsmMask:
    """dmsm/mod.rs:10-57 (normalised Jacobian images)."""

    def __init__(self, in_mask, out_mask):
        self.in_mask, self.out_mask = in_mask, out_mask

    @staticmethod
    def zero(g2=False):
        w = 8 if g2 else 4
        z = np.zeros(3 * w, dtype=np.uint64)
        z[:4] = _one_img_fq()
        z[w:w + 4] = _one_img_fq()
        return MsmMask(z.copy(), z.copy())

    @staticmethod
    def sample(pp: "PackedSharingParams", mask_scalars, rand_in_points, rand_out_points, g2=False):
        """dmsm/mod.rs:21-48.  mask_scalars: the l field draws x_i (mask value i = gen * x_i);
        rand_in_points / rand_out_points: the t random GROUP elements each `pp.pack` call draws
        (normalised Jacobian images; the host RNG stays with the caller).  Returns the n parties' masks.
        Every group operation is a tiny device MSM; pack over group elements applies the rows of the
        pack matrix (pss.rs:90-122 with T = G)."""
        mask_scalars = _fr_vec(mask_scalars, "mask_scalars")
        _require(mask_scalars.shape[0] == pp.l and len(rand_in_points) == pp.t and len(rand_out_points) == pp.t,
                 "MsmMask.sample: l mask scalars and t + t random points expected")
        dev = pp.device
        gen = group_generator(g2)
        values = [group_lincomb([gen], [x], g2, dev) for x in mask_scalars]                 # gen * x_i
        minus_one = fr_sub(np.zeros((1, 4), dtype=np.uint64), _one_img().reshape(1, 4), dev)[0]
        out_value = group_lincomb(values, [minus_one] * pp.l, g2, dev)                     # -(sum of the mask values)
        M = _pack_matrix(pp)                                                               # (n, l + t) Fr images
        ins = [group_lincomb(values + list(rand_in_points), list(M[i]), g2, dev) for i in range(pp.n)]
        outs = [group_lincomb([out_value] * pp.l + list(rand_out_points), list(M[i]), g2, dev) for i in range(pp.n)]
        return [MsmMask(i, o) for i, o in zip(ins, outs)]


_G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634),
           (8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531))


def _fq_image(v: int) -> np.ndarray:
    m = (v % Q_MOD) * (1 << 256) % Q_MOD
    return np.array([(m >> (64 * i)) & _MASK64 for i in range(4)], dtype=np.uint64)


def group_generator(g2=False) -> np.ndarray:
    """G::generator() as a normalised Jacobian image: G1 (1, 2); G2 the standard BN254 generator
    (fixtures/verification_key.json:24-37, vk_gamma_2)."""
    if not g2:
        return np.concatenate([_fq_image(1), _fq_image(2), _fq_image(1)])
    (x0, x1), (y0, y1) = _G2_GEN
    return np.concatenate([_fq_image(x0), _fq_image(x1), _fq_image(y0), _fq_image(y1), _fq_image(1), _fq_image(0)])


_PACK_M = {}


def _pack_matrix(pp: "PackedSharingParams") -> np.ndarray:
    """(n, l + t, 4) Montgomery images of the pack matrix, read off the device by packing unit vectors."""
    if pp.l not in _PACK_M:
        k = pp.l + pp.t
        one = _one_img()
        sec = np.zeros((k * pp.l, 4), dtype=np.uint64)
        rnd = np.zeros((k * pp.t, 4), dtype=np.uint64)
        for c in range(k):
            if c < pp.l:
                sec[c * pp.l + c] = one
            else:
                rnd[c * pp.t + (c - pp.l)] = one
        shares = pp.pack(sec, rnd).reshape(k, pp.n, 4)          # column c = image of unit vector c
        _PACK_M[pp.l] = np.ascontiguousarray(shares.transpose(1, 0, 2))
    return _PACK_M[pp.l]


def _one_img_fq():
    m = (1 << 256) % Q_MOD
    return np.array([(m >> (64 * i)) & _MASK64 for i in range(4)], dtype=np.uint64)


class DegRedMask:
    """utils/deg_red.rs:14-77."""

    def __init__(self, in_mask, out_mask):
        _require(in_mask.shape == out_mask.shape, "in_mask / out_mask of different lengths")
        self.in_mask, self.out_mask = in_mask, out_mask

    @staticmethod
    def zero(num):
        return DegRedMask(np.zeros((num, 4), dtype=np.uint64), np.zeros((num, 4), dtype=np.uint64))

    @staticmethod
    def sample(pp: PackedSharingParams, num, mask_values, rand_in, rand_out):
        """utils/deg_red.rs:40-66 over Fr with gen = 1: mask_values (num*l, 4) are the random draws,
        rand_in / rand_out (num*t, 4) the packing randomness.  Returns the n parties' shares."""
        mask_values, rand_in, rand_out = _fr_vec(mask_values, "mask_values"), _fr_vec(rand_in, "rand_in"), _fr_vec(rand_out, "rand_out")
        _require(mask_values.shape[0] == num * pp.l, "mask_values: num * l draws expected")
        _require(rand_in.shape[0] == num * pp.t and rand_out.shape[0] == num * pp.t, "rand_in / rand_out: num * t draws expected")
        ins = [np.empty((num, 4), dtype=np.uint64) for _ in range(pp.n)]
        outs = [np.empty((num, 4), dtype=np.uint64) for _ in range(pp.n)]
        check(lib().zkg_deg_red_mask_sample_bn254(pp.device, num, pp.l, _ptr(mask_values), _ptr(rand_in), _ptr(rand_out),
                                                  _ptr_array(ins), _ptr_array(outs)))
        return [DegRedMask(i, o) for i, o in zip(ins, outs)]


# ---- in-process stand-in for mpc-net's LocalTestNet -----------------------------------------------
class LocalTestNet:
    """N parties in one process; party 0 is the king (mpc-net/src/lib.rs:65-67).  `dropouts` lists
    parties whose message the king does not receive (simulate_lossy_network_round discards the
    last party's result, mpc-net/src/multi.rs:330-363)."""

    def __init__(self, n_parties, dropouts=()):
        self.n = n_parties
        self.dropouts = tuple(dropouts)

    def n_parties(self):
        return self.n

    def received(self, messages):
        parties = [p for p in range(self.n) if p not in self.dropouts]
        return [messages[p] for p in parties], parties


def d_fft(pcoeff_shares, fft_masks, rearrange, dom, pp, net, rand_points, device_of_party=None):
    """dfft/mod.rs:99-134 for all parties at once: returns the n output share vectors."""
    return _d_fft_impl(pcoeff_shares, fft_masks, rearrange, dom, fr_image(1), pp, net, rand_points, False,
                       device_of_party)


def d_ifft(peval_shares, fft_masks, rearrange, dom, g, pp, net, rand_points, device_of_party=None):
    """dfft/mod.rs:137-175 for all parties at once."""
    return _d_fft_impl(peval_shares, fft_masks, rearrange, dom, g, pp, net, rand_points, True, device_of_party)


def _d_fft_impl(shares, masks, rearrange, dom, g, pp, net, rand_points, inverse, device_of_party):
    mbyl = shares[0].shape[0]
    _require(mbyl * pp.l == dom.size(), f"Mismatch of size in FFT, {mbyl * pp.l}, {dom.size()}.")
    gen = dom.group_gen_inv() if inverse else dom.group_gen()
    pre = dom.size_inv() if inverse else None
    sent = []
    for p in range(net.n_parties()):
        v = _fr_vec(shares[p]).copy()
        dev = device_of_party(p) if device_of_party else pp.device
        fft1_in_place(v, pp, gen, pre_scale=pre, in_mask=masks[p].in_mask, device=dev)     # :121/:159-162, :254-258
        sent.append(v)
    recv, parties = net.received(sent)
    out = king_fft2(recv, parties, pp, gen, g, rearrange, rand_points)                 # :264-304
    return [fr_add(out[p], masks[p].out_mask, pp.device) for p in range(net.n_parties())]  # :313-317


def deg_red(x_shares, masks, pp, net, rand_points):
    """utils/deg_red.rs:80-126 for all parties at once."""
    sent = [fr_add(x_shares[p], masks[p].in_mask, pp.device) for p in range(net.n_parties())]
    recv, parties = net.received(sent)
    out = deg_red_king(recv, parties, pp, rand_points)
    return [fr_add(out[p], masks[p].out_mask, pp.device) for p in range(net.n_parties())]


def d_pp(num_shares, den_shares, masks, pp, net, rand_king, rand_degred):
    """dpp/mod.rs:15-87 for all parties at once (the dummy randomness s = 1 of :24-25 included)."""
    sent = [np.concatenate([_fr_vec(num_shares[p]), _fr_vec(den_shares[p])]) for p in range(net.n_parties())]
    recv, parties = net.received(sent)
    out = dpp_king(recv, parties, pp, rand_king)
    return deg_red(out, masks, pp, net, rand_degred)        # :86


def d_msm(bases_by_party, scalars_by_party, masks, pp, net, g2=False, device_of_party=None, wire=True):
    """dmsm/mod.rs:59-102 for all parties at once.  The king's `unpack_missing_shares` over group elements
    (pss.rs:141-166; the Lagrange path :170-221 when `net` drops parties) and the sum run in
    zkg_pss_unpack2_bn254_g1/g2."""
    msm = msm_g2 if g2 else msm_g1
    c_shares = []
    for p in range(net.n_parties()):
        dev = device_of_party(p) if device_of_party else pp.device
        c = msm(bases_by_party[p], scalars_by_party[p], dev)                           # :73
        c_shares.append(group_add(c, masks[p].in_mask, g2, dev))                       # :74
    if wire:
        # client_send_or_king_receive_serialized (mpc-net/src/ser_net.rs:25,40): every share crosses as a compressed point
        frames = [group_to_wire(c, g2, pp.device)[0] for c in c_shares]
        recv_frames, parties = net.received(frames)
        recv = list(group_from_wire(np.stack(recv_frames), g2, pp.device))
    else:
        recv, parties = net.received(c_shares)
    _, output = pss_unpack2_group(pp, recv, parties, g2, want_unpacked=False)          # :85-86
    if wire:                                                                           # client_receive_or_king_send_serialized, :112-119
        output = group_from_wire(group_to_wire(output, g2, pp.device), g2, pp.device)[0]
    return [group_add(output, masks[p].out_mask, g2, pp.device) for p in range(net.n_parties())]  # :98
