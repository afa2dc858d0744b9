"""oracle/pyref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Pure-Python big-integer restatement of the zk-SaaS hot path (BN254), written to be
*obviously* equal to the mathematical definition so that it can pin the faster C
restatement in oracle/zkoracle.c and, through it, the CUDA kernels.  It is exact and
slow: use it for sizes up to a few thousand elements.

Parity status: the arithmetic of the reference lives in the un-vendored, un-pinned
arkworks 0.4 crates (ark-ff / ark-ec / ark-poly / ark-bn254; no Cargo.lock in the
reference tree) and there is no Rust toolchain in this image, so the reference cannot
be run here: **parity unpinned at the arkworks binary**.  What this file pins instead:
the BN254 constants and on-curve points that *are* in the reference tree
(fixtures/verifier.sol:26-37,52,216 and fixtures/verification_key.json:5-51) and the
equalities the reference's own tests assert (see tests/test_oracle_pins.py).

Every function cites the reference file:line it restates (paths relative to the
reference root).
"""
from __future__ import annotations

# ----------------------------------------------------------------------------------------
# BN254 constants.  R_MOD: fixtures/verifier.sol:216 ; Q_MOD: fixtures/verifier.sol:52.
# ----------------------------------------------------------------------------------------
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_GENERATOR = 5          # ark-bn254 FrConfig::GENERATOR (multiplicative generator of Fr*)
FR_TWO_ADICITY = 28
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)
MONT_R = 1 << 256         # ark-ff MontBackend<_,4>: residue a*2^256 mod p, 4 LE u64 limbs
G1_B = 3                  # y^2 = x^3 + 3
G1_GEN = (1, 2)           # fixtures/verifier.sol:26-27 (P1)
# Fq2 = Fq[u]/(u^2+1); G2: y^2 = x^3 + 3/(9+u)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)                         # fixtures/verification_key.json:24-37 (vk_gamma_2), (c0,c1) order


# ----------------------------------------------------------------------------------------
# Prime field helpers
# ----------------------------------------------------------------------------------------
def finv(a: int, p: int) -> int:
    return pow(a, p - 2, p)


def to_mont_limbs(a: int, p: int) -> list[int]:
    """arkworks in-memory image of Fp<MontBackend<_,4>,4>: (a*2^256 mod p) as 4 LE u64."""
    v = (a % p) * MONT_R % p
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_limbs(limbs, p: int) -> int:
    v = sum(int(x) << (64 * i) for i, x in enumerate(limbs))
    assert v < p, "non-canonical Montgomery residue"
    return v * finv(MONT_R % p, p) % p


def ceil_log2(x: int) -> int:
    """ark_std::log2 = ceil(log2 x) (0 for x<=1)."""
    return 0 if x <= 1 else (x - 1).bit_length()


# ----------------------------------------------------------------------------------------
# ark-poly Radix2EvaluationDomain (restated from the public arkworks 0.4 behaviour,
# SURVEY.md Appendix A).  Used by secret-sharing/src/pss.rs:44-52 and dfft tests.
# ----------------------------------------------------------------------------------------
class Radix2Domain:
    def __init__(self, num_coeffs: int, offset: int = 1, p: int = R_MOD):
        self.p = p
        self.size = 1 << ceil_log2(num_coeffs) if num_coeffs > 1 else 1
        self.log_size = ceil_log2(self.size)
        assert self.log_size <= FR_TWO_ADICITY
        # F::get_root_of_unity(size) = TWO_ADIC_ROOT ^ (2^(28 - log size))
        self.group_gen = pow(FR_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - self.log_size), p)
        self.group_gen_inv = finv(self.group_gen, p)
        self.size_inv = finv(self.size % p, p)
        self.offset = offset % p
        self.offset_inv = finv(self.offset, p)

    def get_coset(self, offset: int) -> "Radix2Domain":
        return Radix2Domain(self.size, offset, self.p)

    def element(self, i: int) -> int:
        return self.offset * pow(self.group_gen, i, self.p) % self.p

    def elements(self):
        return [self.element(i) for i in range(self.size)]

    # generic over DomainCoeff: `mul(coeff, scalar)` and `add` are passed in so that the
    # same code runs over field elements and over group elements (pss.rs is generic).
    def _dft(self, v, root, ops):
        n = self.size
        zero, add, mul = ops
        out = []
        for k in range(n):
            acc = zero
            for j in range(n):
                acc = add(acc, mul(v[j], pow(root, (j * k) % n, self.p)))
            out.append(acc)
        return out

    def fft(self, v, ops=None):
        """fft_in_place: resize (truncate / zero-pad) to domain size, coset-scale, in-order DFT."""
        ops = ops or field_ops(self.p)
        zero, add, mul = ops
        v = list(v[: self.size]) + [zero] * max(0, self.size - len(v))
        if self.offset != 1:
            v = [mul(x, pow(self.offset, i, self.p)) for i, x in enumerate(v)]
        if self.size > 64 and ops is _FIELD_OPS.get(self.p):
            return _fast_ntt(v, self.group_gen, self.p)
        return self._dft(v, self.group_gen, ops)

    def ifft(self, v, ops=None):
        ops = ops or field_ops(self.p)
        zero, add, mul = ops
        v = list(v[: self.size]) + [zero] * max(0, self.size - len(v))
        if self.size > 64 and ops is _FIELD_OPS.get(self.p):
            out = _fast_ntt(v, self.group_gen_inv, self.p)
        else:
            out = self._dft(v, self.group_gen_inv, ops)
        out = [mul(x, self.size_inv) for x in out]
        if self.offset != 1:
            out = [mul(x, pow(self.offset_inv, i, self.p)) for i, x in enumerate(out)]
        return out


_FIELD_OPS: dict = {}


def field_ops(p: int):
    if p not in _FIELD_OPS:
        _FIELD_OPS[p] = (0, lambda a, b: (a + b) % p, lambda a, s: a * s % p)
    return _FIELD_OPS[p]


def _fast_ntt(v, root, p):
    """Recursive radix-2 DFT, out[k] = sum v[j] root^(jk); only an accelerator for _dft."""
    n = len(v)
    if n == 1:
        return list(v)
    even = _fast_ntt(v[0::2], root * root % p, p)
    odd = _fast_ntt(v[1::2], root * root % p, p)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t = w * odd[k] % p
        out[k] = (even[k] + t) % p
        out[k + n // 2] = (even[k] - t) % p
        w = w * root % p
    return out


def distribute_powers(v, g, p=R_MOD):
    """Radix2EvaluationDomain::distribute_powers: v[i] *= g^i (dfft/mod.rs:49,279)."""
    out, pw = [], 1
    for x in v:
        out.append(x * pw % p)
        pw = pw * g % p
    return out


# ----------------------------------------------------------------------------------------
# secret-sharing/src/pss.rs
# ----------------------------------------------------------------------------------------
class PackedSharingParams:
    """secret-sharing/src/pss.rs:19-66."""

    def __init__(self, l: int, p: int = R_MOD):
        self.l, self.t, self.n, self.p = l, l, 4 * l, p          # pss.rs:40-42
        self.share = Radix2Domain(self.n, 1, p)                    # pss.rs:44
        self.secret = Radix2Domain(l + self.t, 1, p).get_coset(FR_GENERATOR)       # :45-48
        self.secret2 = Radix2Domain(2 * (l + self.t), 1, p).get_coset(FR_GENERATOR)  # :49-52

    def det_pack(self, secrets, ops=None):
        """pss.rs:69-87 (resize(t) then ifft's own resize to l+t ==> zero padding)."""
        ops = ops or field_ops(self.p)
        assert len(secrets) == self.l
        result = list(secrets)
        if len(result) < self.t:
            result += [ops[0]] * (self.t - len(result))
        else:
            result = result[: self.t]
        result = self.secret.ifft(result, ops)
        return self.share.fft(result, ops)

    def pack(self, secrets, rand_points, ops=None):
        """pss.rs:90-122; the t random points are an explicit input (host RNG stays outside)."""
        ops = ops or field_ops(self.p)
        assert len(secrets) == self.l and len(rand_points) == self.t
        result = list(secrets) + list(rand_points)
        result = self.secret.ifft(result, ops)
        return self.share.fft(result, ops)

    def unpack(self, shares, ops=None):
        """pss.rs:125-138."""
        ops = ops or field_ops(self.p)
        result = self.share.ifft(shares, ops)
        result = self.secret.fft(result, ops)      # truncates to l+t coefficients first
        return result[: self.l]

    def unpack2(self, shares, ops=None):
        """pss.rs:141-166."""
        ops = ops or field_ops(self.p)
        result = self.share.ifft(shares, ops)
        result = self.secret2.fft(result, ops)
        return result[0: 2 * self.l: 2]

    def lagrange_unpack(self, shares, parties, ops=None):
        """pss.rs:170-207 (O(k^2) interpolation, secret-sharing/src/utils.rs:78-116)."""
        ops = ops or field_ops(self.p)
        assert len(shares) == len(parties)
        assert len(parties) > 2 * (self.t + self.l - 1)
        elems = self.share.elements()
        xs = [elems[i] for i in parties]
        result = lagrange_interpolate(xs, shares, self.p, ops)
        result = self.secret2.fft(result, ops)
        return result[0: 2 * self.l: 2]

    def unpack_missing_shares(self, shares, parties, ops=None):
        """pss.rs:210-221."""
        if len(shares) == self.n:
            return self.unpack2(shares, ops)
        return self.lagrange_unpack(shares, parties, ops)

    # Closed forms of the linear maps (what the CUDA kernels apply as constant matrices).
    def pack_matrix(self):
        """n x (l+t): shares = M @ (secrets || rand)."""
        k = self.l + self.t
        cols = []
        for j in range(k):
            e = [0] * k
            e[j] = 1
            cols.append(self.share.fft(self.secret.ifft(e)))
        return [[cols[j][i] for j in range(k)] for i in range(self.n)]

    def unpack2_matrix(self):
        """l x n: secrets = M @ shares."""
        cols = []
        for j in range(self.n):
            e = [0] * self.n
            e[j] = 1
            cols.append(self.unpack2(e))
        return [[cols[j][i] for j in range(self.n)] for i in range(self.l)]

    def unpack_matrix(self):
        cols = []
        for j in range(self.n):
            e = [0] * self.n
            e[j] = 1
            cols.append(self.unpack(e))
        return [[cols[j][i] for j in range(self.n)] for i in range(self.l)]


def get_zero_roots(xs, p):
    """secret-sharing/src/utils.rs:120-135."""
    result = [0] * (len(xs) + 1)
    n = len(result) - 1
    result[n] = 1
    for i in range(len(xs)):
        n -= 1
        result[n] = 0
        for j in range(n, len(xs)):
            result[j] = (result[j] - result[j + 1] * xs[i]) % p
    return result


def syn_div_1(poly, b, p):
    """secret-sharing/src/utils.rs:37-55, a == 1 branch (divide by x - b)."""
    out = list(poly)
    c = 0
    for i in range(len(out) - 1, -1, -1):
        out[i] = (out[i] + b * c) % p
        out[i], c = c, out[i]
    return out


def lagrange_interpolate(xs, ys, p, ops):
    """secret-sharing/src/utils.rs:78-116."""
    zero, add, mul = ops
    roots = get_zero_roots(xs, p)
    numerators = [syn_div_1(roots, x, p) for x in xs]
    denominators = []
    for f, x in zip(numerators, xs):
        acc = 0
        for c in reversed(f):
            acc = (acc * x + c) % p
        denominators.append(finv(acc, p))
    result = [zero] * len(numerators)
    for i in range(len(ys)):
        y = mul(ys[i], denominators[i])
        for j in range(len(result)):
            result[j] = add(result[j], mul(y, numerators[i][j]))
    while result and result[-1] == zero:
        result.pop()
    return result


# ----------------------------------------------------------------------------------------
# dist-primitives/src/utils/pack.rs and dist-primitives/src/dfft/mod.rs
# ----------------------------------------------------------------------------------------
def transpose(matrix):
    """dist-primitives/src/utils/pack.rs:22-35."""
    assert matrix
    return [[row[c] for row in matrix] for c in range(len(matrix[0]))]


def pack_vec(secrets, pp: PackedSharingParams, rand):
    """pack.rs:8-20; rand[i] = the t random points of chunk i."""
    assert len(secrets) % pp.l == 0
    return [pp.pack(secrets[i * pp.l:(i + 1) * pp.l], rand[i]) for i in range(len(secrets) // pp.l)]


def fft_in_place_rearrange(data):
    """dist-primitives/src/dfft/mod.rs:322-335 (bit-reversal permutation), literal."""
    data = list(data)
    target = 0
    for pos in range(len(data)):
        if target > pos:
            data[target], data[pos] = data[pos], data[target]
        mask = len(data) >> 1
        while target & mask != 0:
            target &= ~mask
            mask >>= 1
        target |= mask
    return data


def fft1_in_place(px, pp: PackedSharingParams, gen, p=R_MOD):
    """dist-primitives/src/dfft/mod.rs:178-208, literal."""
    px = list(px)
    dom_size = len(px) * pp.l
    for i in range(ceil_log2(dom_size), ceil_log2(pp.l), -1):
        poly_size = dom_size // (1 << i)
        factor_stride = pow(gen, 1 << (i - 1), p)
        factor = factor_stride
        for k in range(poly_size):
            for j in range((1 << (i - 1)) // pp.l):
                x = px[(2 * j) * poly_size + k]
                y = px[(2 * j + 1) * poly_size + k] * factor % p
                px[j * (2 * poly_size) + k] = (x + y) % p
                px[j * (2 * poly_size) + k + poly_size] = (x - y) % p
            factor = factor * factor_stride % p
    return px


def fft2_in_place(s1, pp: PackedSharingParams, gen, p=R_MOD):
    """dist-primitives/src/dfft/mod.rs:210-237, literal."""
    s1 = list(s1)
    dom_size = len(s1)
    s2 = [0] * dom_size
    for i in range(ceil_log2(pp.l), 0, -1):
        poly_size = dom_size // (1 << i)
        factor_stride = pow(gen, 1 << (i - 1), p)
        factor = factor_stride
        for k in range(poly_size):
            for j in range(1 << (i - 1)):
                x = s1[k * (1 << i) + 2 * j]
                y = s1[k * (1 << i) + 2 * j + 1] * factor % p
                s2[k * (1 << (i - 1)) + j] = (x + y) % p
                s2[(k + poly_size) * (1 << (i - 1)) + j] = (x - y) % p
            factor = factor * factor_stride % p
        s1, s2 = s2, s1
    return s1[-1:] + s1[:-1]            # rotate_right(1), dfft/mod.rs:236


def king_fft2(shares_by_party, parties, pp: PackedSharingParams, gen, g, rearrange, rand, p=R_MOD):
    """King closure of fft2_with_rearrange, dist-primitives/src/dfft/mod.rs:264-304.

    shares_by_party[r] = masked share vector (length m/l) received from parties[r];
    rand[i] = the t random points used to re-pack output column i.  Returns the per-party
    output vectors (party-major, n x m/l).
    """
    mbyl = len(shares_by_party[0])
    all_shares = transpose(shares_by_party)                     # :265
    s1 = [0] * (mbyl * pp.l)
    for i in range(mbyl):                                        # :268-274
        tmp = pp.unpack_missing_shares(all_shares[i], parties)
        for j in range(pp.l):
            s1[i * pp.l + j] = tmp[j]
    s1 = fft2_in_place(s1, pp, gen, p)                           # :276
    if g % p != 1:                                               # :278-280
        s1 = distribute_powers(s1, g, p)
    if rearrange:                                                # :284-300
        s1 = fft_in_place_rearrange(s1)
        out_shares = [pp.pack(s1[i::len(s1) // pp.l][: pp.l], rand[i]) for i in range(len(s1) // pp.l)]
        return transpose(out_shares)
    return transpose(pack_vec(s1, pp, rand))                     # :302


def d_fft_round(pcoeff_shares, in_masks, out_masks, rearrange, m, pp, rand, inverse=False, g=1, p=R_MOD, parties=None):
    """In-process n-party emulation of d_fft (dfft/mod.rs:99-134) / d_ifft (:137-175).

    pcoeff_shares[party] = that party's share vector (length m/l); in_masks/out_masks likewise.
    parties: the parties whose message reaches the king (default all; a lossy round, mpc-net/src/multi.rs:330-363, drops some).
    """
    dom = Radix2Domain(m, 1, p)
    gen = dom.group_gen_inv if inverse else dom.group_gen
    sent = []
    for party in range(pp.n):
        v = list(pcoeff_shares[party])
        if inverse:
            v = [x * dom.size_inv % p for x in v]                # :159
        v = fft1_in_place(v, pp, gen, p)                         # :121 / :162
        v = [(x + mk) % p for x, mk in zip(v, in_masks[party])]  # :254-258
        sent.append(v)
    parties = list(range(pp.n)) if parties is None else list(parties)
    out = king_fft2([sent[q] for q in parties], parties, pp, gen, g, rearrange, rand, p)
    return [[(x + mk) % p for x, mk in zip(out[party], out_masks[party])] for party in range(pp.n)]  # :313-317


def fft_mask_sample(rearrange, g, gen, m, pp: PackedSharingParams, mask_values, rand_in, rand_out, p=R_MOD):
    """FftMask::sample, dist-primitives/src/dfft/mod.rs:30-85, with the draws passed in: mask_values (m), rand_in[i] /
    rand_out[i] = the t packing draws of column i.  Returns (in_mask_shares, out_mask_shares), party-major."""
    mbyl = m // pp.l
    in_shares = transpose(pack_vec(list(mask_values), pp, rand_in))                  # :41-42
    s = fft2_in_place(list(mask_values), pp, gen, p)                                 # :44
    if g % p != 1:
        s = distribute_powers(s, g, p)                                               # :46-48
    s = [(-v) % p for v in s]                                                        # :51
    if rearrange:                                                                    # :55-72
        s = fft_in_place_rearrange(s)
        out_shares = transpose([pp.pack(s[i::mbyl][: pp.l], rand_out[i]) for i in range(mbyl)])
    else:
        out_shares = transpose(pack_vec(s, pp, rand_out))                            # :74
    return in_shares, out_shares


def deg_red_mask_sample(pp: PackedSharingParams, num, mask_values, rand_in, rand_out, p=R_MOD):
    """DegRedMask::sample over F with gen = 1, dist-primitives/src/utils/deg_red.rs:40-66: in = pack(values), out = pack(-values)."""
    assert len(mask_values) == num * pp.l
    in_shares = transpose(pack_vec(list(mask_values), pp, rand_in))
    out_shares = transpose(pack_vec([(-v) % p for v in mask_values], pp, rand_out))
    return in_shares, out_shares


def deg_red_king(shares_by_party, parties, pp, rand):
    """King closure of deg_red, dist-primitives/src/utils/deg_red.rs:103-111."""
    cols = transpose(shares_by_party)
    out = []
    for i, col in enumerate(cols):
        xi = pp.unpack_missing_shares(col, parties)
        out.append(pp.pack(xi, rand[i]))
    return transpose(out)


def dpp_king(shares_by_party, parties, pp, rand, p=R_MOD):
    """King closure of d_pp, dist-primitives/src/dpp/mod.rs:41-76 (shares: num columns then den columns)."""
    cols2 = len(shares_by_party[0])
    numden = []
    for col in transpose(shares_by_party):
        numden += pp.unpack_missing_shares(col, parties)
    half = len(numden) // 2
    q = [numden[i] * finv(numden[i + half], p) % p for i in range(half)]
    for i in range(1, half):
        q[i] = q[i] * q[i - 1] % p
    assert cols2 % 2 == 0
    return transpose(pack_vec(q, pp, rand))


# ----------------------------------------------------------------------------------------
# Fq2 and curve arithmetic (affine, big-int; None = point at infinity)
# ----------------------------------------------------------------------------------------
class Fq2:
    """Fq[u]/(u^2+1) (ark-bn254 Fq2Config, NONRESIDUE = -1)."""
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1=0):
        self.c0, self.c1 = c0 % Q_MOD, c1 % Q_MOD

    def __add__(self, o): return Fq2(self.c0 + o.c0, self.c1 + o.c1)
    def __sub__(self, o): return Fq2(self.c0 - o.c0, self.c1 - o.c1)
    def __neg__(self): return Fq2(-self.c0, -self.c1)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq2(self.c0 * o, self.c1 * o)
        return Fq2(self.c0 * o.c0 - self.c1 * o.c1, self.c0 * o.c1 + self.c1 * o.c0)

    def __eq__(self, o): return isinstance(o, Fq2) and self.c0 == o.c0 and self.c1 == o.c1
    def __hash__(self): return hash((self.c0, self.c1))
    def is_zero(self): return self.c0 == 0 and self.c1 == 0

    def inv(self):
        d = finv((self.c0 * self.c0 + self.c1 * self.c1) % Q_MOD, Q_MOD)
        return Fq2(self.c0 * d, -self.c1 * d)

    def __repr__(self): return f"Fq2({self.c0},{self.c1})"


G2_B = Fq2(3) * Fq2(9, 1).inv()     # b' = 3/(9+u)


class Curve:
    """Short Weierstrass y^2 = x^3 + b, a = 0, over Fq (G1) or Fq2 (G2)."""

    def __init__(self, b, lift, inv):
        self.b, self.lift, self.inv = b, lift, inv

    def on_curve(self, P):
        if P is None:
            return True
        x, y = P
        lhs, rhs = y * y, x * x * x + self.b
        if isinstance(lhs, Fq2):
            return lhs == rhs
        return (lhs - rhs) % Q_MOD == 0

    def neg(self, P):
        return None if P is None else (P[0], -P[1] if isinstance(P[1], Fq2) else (-P[1]) % Q_MOD)

    def add(self, P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if y1 == y2 and not self._is_zero(y1):
                lam = (x1 * x1 * 3) * self.inv(y1 * 2)
            else:
                return None
        else:
            lam = (y2 - y1) * self.inv(x2 - x1)
        if not isinstance(lam, Fq2):
            lam %= Q_MOD
        x3 = lam * lam - x1 - x2
        y3 = lam * (x1 - x3) - y1
        if not isinstance(x3, Fq2):
            x3 %= Q_MOD
            y3 %= Q_MOD
        return (x3, y3)

    @staticmethod
    def _is_zero(v):
        return v.is_zero() if isinstance(v, Fq2) else v % Q_MOD == 0

    def mul(self, P, k: int):
        return self.mul_raw(P, k % R_MOD)

    def mul_raw(self, P, k: int):
        """double-and-add without reducing k (points outside the r-torsion subgroup, e.g. the wire-format checks)."""
        acc = None
        while k:
            if k & 1:
                acc = self.add(acc, P)
            P = self.add(P, P)
            k >>= 1
        return acc


G1 = Curve(G1_B, lambda v: v % Q_MOD, lambda v: finv(v % Q_MOD, Q_MOD))
G2 = Curve(G2_B, lambda v: v, lambda v: v.inv())
G2_GEN_PT = (Fq2(*G2_GEN[0]), Fq2(*G2_GEN[1]))


def group_ops(curve: Curve):
    """DomainCoeff ops over group elements: pack/unpack2 also run on points (dmsm/mod.rs:34,38,85)."""
    return (None, curve.add, curve.mul)


def msm_naive(curve: Curve, bases, scalars):
    """Mathematical definition of VariableBaseMSM::msm (call site dmsm/mod.rs:73)."""
    if len(bases) != len(scalars):
        raise ValueError(min(len(bases), len(scalars)))        # Err(min len)
    acc = None
    for P, s in zip(bases, scalars):
        acc = curve.add(acc, curve.mul(P, s))
    return acc


def ark_window_size(size: int) -> int:
    """ark-ec 0.4.2 msm_bigint_wnaf: c = size<32 ? 3 : ln_without_floats(size)+2."""
    if size < 32:
        return 3
    return ceil_log2(size) * 69 // 100 + 2


def make_digits(scalar: int, c: int, num_bits: int = 254):
    """ark-ec 0.4.2 make_digits (signed radix-2^c), SURVEY.md Appendix A."""
    digits_count = (num_bits + c - 1) // c
    radix, window_mask = 1 << c, (1 << c) - 1
    carry, out = 0, []
    for i in range(digits_count):
        coef = carry + ((scalar >> (i * c)) & window_mask)
        carry = (coef + radix // 2) >> c
        d = coef - (carry << c)
        if i == digits_count - 1:
            d += carry << c
        out.append(d)
    return out


def msm_pippenger(curve: Curve, bases, scalars):
    """ark-ec 0.4.2 msm_bigint_wnaf restated."""
    if len(bases) != len(scalars):
        raise ValueError(min(len(bases), len(scalars)))
    size = len(bases)
    if size == 0:
        return None
    c = ark_window_size(size)
    digs = [make_digits(s % R_MOD, c) for s in scalars]
    nwin = len(digs[0])
    window_sums = []
    for w in range(nwin):
        buckets = [None] * (1 << c)
        for d, base in zip(digs, bases):
            v = d[w]
            if v > 0:
                buckets[v - 1] = curve.add(buckets[v - 1], base)
            elif v < 0:
                buckets[-v - 1] = curve.add(buckets[-v - 1], curve.neg(base))
        running, res = None, None
        for b in reversed(buckets):
            running = curve.add(running, b)
            res = curve.add(res, running)
        window_sums.append(res)
    total = None
    for s in reversed(window_sums[1:]):
        total = curve.add(total, s)
        for _ in range(c):
            total = curve.add(total, total)
    return curve.add(window_sums[0], total)


# ----------------------------------------------------------------------------------------
# arkworks memory images at the FFI boundary
# ----------------------------------------------------------------------------------------
def g1_affine_image(P) -> bytes:
    """short_weierstrass::Affine<g1::Config>: x@0 (32B) y@32 (32B) infinity@64, size 72."""
    import struct
    if P is None:
        return struct.pack("<9Q", 0, 0, 0, 0, 0, 0, 0, 0, 1)
    return struct.pack("<9Q", *to_mont_limbs(P[0], Q_MOD), *to_mont_limbs(P[1], Q_MOD), 0)


def g2_affine_image(P) -> bytes:
    """Affine<g2::Config>: x.c0,x.c1,y.c0,y.c1 (4x32B) infinity@128, size 136."""
    import struct
    if P is None:
        return struct.pack("<17Q", *([0] * 16), 1)
    x, y = P
    limbs = to_mont_limbs(x.c0, Q_MOD) + to_mont_limbs(x.c1, Q_MOD) + \
        to_mont_limbs(y.c0, Q_MOD) + to_mont_limbs(y.c1, Q_MOD)
    return struct.pack("<17Q", *limbs, 0)


# ----------------------------------------------------------------------------------------
# ark-serialize 0.4 compressed form of short-Weierstrass points -- the wire form of the group element
# d_msm ships (mpc-net/src/ser_net.rs:25 serialize_compressed, :40,:119 deserialize_compressed; call sites
# dist-primitives/src/dmsm/mod.rs:79-81,90-92).  ark-serialize / ark-ec are un-vendored: the
# layout below restates their public 0.4.2 behaviour and is NOT pinned by any vector in the reference tree.
#   x canonical little-endian (G1: 32 B; G2: c0 then c1, 64 B); SWFlags in the two top bits of the LAST byte:
#   bit 7 = YIsNegative (y > -y), bit 6 = PointAtInfinity (x = 0), both clear = YIsPositive (y <= -y).
#   Fq2 ordering is lexicographic with c1 most significant (ark-ff QuadExtField::cmp).
#   deserialize_compressed validates: both flags set / x >= q / x^3 + b a non-residue / (G2) point outside the
#   r-torsion subgroup are errors.
# ----------------------------------------------------------------------------------------
class WireError(ValueError):
    pass


def _fq_sqrt(a: int):
    """q = 3 mod 4: a^((q+1)/4), or None for a non-residue."""
    a %= Q_MOD
    y = pow(a, (Q_MOD + 1) // 4, Q_MOD)
    return y if y * y % Q_MOD == a else None


def _fq2_sqrt(a: Fq2):
    """Square root in Fq[u]/(u^2+1) by the norm method; None for a non-residue.  Either root may be returned."""
    if a.is_zero():
        return Fq2(0)
    if a.c1 == 0:
        s = _fq_sqrt(a.c0)
        if s is not None:
            return Fq2(s, 0)
        return Fq2(0, _fq_sqrt(-a.c0))                 # -1 is a non-residue, so -a0 is a residue
    alpha = _fq_sqrt(a.c0 * a.c0 + a.c1 * a.c1)
    if alpha is None:
        return None
    half = finv(2, Q_MOD)
    delta = (a.c0 + alpha) * half % Q_MOD
    c0 = _fq_sqrt(delta)
    if c0 is None:
        c0 = _fq_sqrt((delta - alpha) % Q_MOD)
    c1 = a.c1 * finv(2 * c0 % Q_MOD, Q_MOD) % Q_MOD
    r = Fq2(c0, c1)
    assert r * r == a
    return r


def _fq2_gt(a: Fq2, b: Fq2) -> bool:
    return (a.c1, a.c0) > (b.c1, b.c0)


def g1_serialize_compressed(P) -> bytes:
    if P is None:
        return bytes(31) + bytes([0x40])
    x, y = P
    out = bytearray(x.to_bytes(32, "little"))
    if y > (-y) % Q_MOD:
        out[31] |= 0x80
    return bytes(out)


def g1_deserialize_compressed(b: bytes):
    if len(b) != 32:
        raise WireError("length")
    flags = b[31] >> 6
    if flags == 3:
        raise WireError("both flags set")
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x3F]), "little")
    if x >= Q_MOD:
        raise WireError("x not below the modulus")
    if flags == 1:
        return None
    y = _fq_sqrt(x * x * x + G1_B)
    if y is None:
        raise WireError("x is not on the curve")
    small, large = sorted((y, (-y) % Q_MOD))
    return (x, large if flags == 2 else small)


def g2_serialize_compressed(P) -> bytes:
    if P is None:
        return bytes(63) + bytes([0x40])
    x, y = P
    out = bytearray(x.c0.to_bytes(32, "little") + x.c1.to_bytes(32, "little"))
    if _fq2_gt(y, -y):
        out[63] |= 0x80
    return bytes(out)


def g2_deserialize_compressed(b: bytes):
    if len(b) != 64:
        raise WireError("length")
    flags = b[63] >> 6
    if flags == 3:
        raise WireError("both flags set")
    c0 = int.from_bytes(b[:32], "little")
    c1 = int.from_bytes(bytes(b[32:63]) + bytes([b[63] & 0x3F]), "little")
    if c0 >= Q_MOD or c1 >= Q_MOD:
        raise WireError("x not below the modulus")
    if flags == 1:
        return None
    x = Fq2(c0, c1)
    y = _fq2_sqrt(x * x * x + G2_B)
    if y is None:
        raise WireError("x is not on the curve")
    ny = -y
    small, large = (ny, y) if _fq2_gt(y, ny) else (y, ny)
    P = (x, large if flags == 2 else small)
    # Validate::Yes: r-torsion check (G1 has cofactor 1 and needs none)
    if G2.add(G2.mul_raw(P, R_MOD - 1), P) is not None:
        raise WireError("point is not in the prime-order subgroup")
    return P
