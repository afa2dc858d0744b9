"""oracle/pairing.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

BN254 optimal-ate pairing over the oracle's OWN field and curve arithmetic (oracle/pyref.py: Fq2, the
G1 / G2 `Curve` objects).  It exists for one purpose: the reference tree ships exactly one numerical
relation that involves this path's arithmetic,

    fixtures/verification_key.json:52-81   vk_alphabeta_12 = e(vk_alpha_1, vk_beta_2)   (:5-9, :10-23)

so evaluating the pairing with pyref's Fq / Fq2 / G1 / G2 code and comparing with that literal ties the
oracle's field tower and curve formulas to a value the REFERENCE holds (the Groth16 verifier of
groth16/examples/sha256.rs:400,415 consumes the same key).  Bilinearity then extends the pin to scalar
multiplication: e([a]alpha, [b]beta) = vk_alphabeta_12^(ab) checks `Curve.mul` on both groups.

Tower (ark-bn254 / snarkjs, identical): Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), xi = 9 + u,
Fq12 = Fq6[w]/(w^2 - v).  The JSON nests an Fq12 as [c0, c1] of Fq6 = [c0, c1, c2] of Fq2 = [c0, c1].
Every G2 step of the Miller loop is computed TWICE: by the slope formula here (needed for the line) and by
pyref.G2.add, and the two are asserted equal, so the G2 addition / doubling the MSM oracle uses is what the
fixture pins.
"""
from __future__ import annotations

import pyref            # oracle/ is put on sys.path by tests/oracle_lib.py
from pyref import Fq2, Q_MOD, R_MOD

XI = Fq2(9, 1)
BN_Z = 4965661367192848881                      # BN parameter of BN254 (x in the BN polynomial family)
ATE_LOOP = 6 * BN_Z + 2


def fq2_pow(a: Fq2, e: int) -> Fq2:
    r = Fq2(1)
    while e:
        if e & 1:
            r = r * a
        a = a * a
        e >>= 1
    return r


class Fq6:
    __slots__ = ("c",)

    def __init__(self, c0=None, c1=None, c2=None):
        self.c = (c0 or Fq2(0), c1 or Fq2(0), c2 or Fq2(0))

    def __add__(self, o): return Fq6(*(a + b for a, b in zip(self.c, o.c)))
    def __sub__(self, o): return Fq6(*(a - b for a, b in zip(self.c, o.c)))
    def __neg__(self): return Fq6(*(-a for a in self.c))
    def __eq__(self, o): return self.c == o.c
    def is_zero(self): return all(a.is_zero() for a in self.c)

    def __mul__(self, o):
        a0, a1, a2 = self.c
        b0, b1, b2 = o.c
        # schoolbook with v^3 = xi
        return Fq6(a0 * b0 + (a1 * b2 + a2 * b1) * XI,
                   a0 * b1 + a1 * b0 + (a2 * b2) * XI,
                   a0 * b2 + a1 * b1 + a2 * b0)

    def mul_by_v(self):
        a0, a1, a2 = self.c
        return Fq6(a2 * XI, a0, a1)

    def inv(self):
        a0, a1, a2 = self.c
        t0 = a0 * a0 - (a1 * a2) * XI
        t1 = (a2 * a2) * XI - a0 * a1
        t2 = a1 * a1 - a0 * a2
        d = (a0 * t0 + (a2 * t1 + a1 * t2) * XI).inv()
        return Fq6(t0 * d, t1 * d, t2 * d)


class Fq12:
    __slots__ = ("c0", "c1")

    def __init__(self, c0: Fq6, c1: Fq6):
        self.c0, self.c1 = c0, c1

    @staticmethod
    def one():
        return Fq12(Fq6(Fq2(1)), Fq6())

    def __mul__(self, o):
        # w^2 = v
        return Fq12(self.c0 * o.c0 + (self.c1 * o.c1).mul_by_v(), self.c0 * o.c1 + self.c1 * o.c0)

    def __eq__(self, o): return self.c0 == o.c0 and self.c1 == o.c1

    def inv(self):
        d = (self.c0 * self.c0 - (self.c1 * self.c1).mul_by_v()).inv()
        return Fq12(self.c0 * d, -(self.c1 * d))

    def pow(self, e: int):
        r, a = Fq12.one(), self
        while e:
            if e & 1:
                r = r * a
            a = a * a
            e >>= 1
        return r

    def to_json_ints(self):
        """[[[c0.c0.c0, c0.c0.c1], ...], [...]] as in fixtures/verification_key.json:52-81."""
        return [[[f.c0, f.c1] for f in h.c] for h in (self.c0, self.c1)]


def _line(lam: Fq2, T, P) -> Fq12:
    """Line through the untwisted T with (twist) slope lam, evaluated at P = (xP, yP) in G1.
    Untwist (D-type): (x', y') -> (x' w^2, y' w^3); the slope picks up one w.  l(P) = yP - lam xP w + (lam xT - yT) w^3,
    and w^3 = v w."""
    xP, yP = P
    xT, yT = T
    return Fq12(Fq6(Fq2(yP)), Fq6(-(lam * xP), lam * xT - yT))


def _dbl_step(T, P):
    x, y = T
    lam = (x * x * 3) * (y * 2).inv()
    x3 = lam * lam - x - x
    y3 = lam * (x - x3) - y
    assert (x3, y3) == pyref.G2.add(T, T), "pairing: doubling disagrees with pyref.G2.add"
    return (x3, y3), _line(lam, T, P)


def _add_step(T, Q, P):
    (x1, y1), (x2, y2) = T, Q
    lam = (y2 - y1) * (x2 - x1).inv()
    x3 = lam * lam - x1 - x2
    y3 = lam * (x1 - x3) - y1
    assert (x3, y3) == pyref.G2.add(T, Q), "pairing: addition disagrees with pyref.G2.add"
    return (x3, y3), _line(lam, T, P)


def frobenius_twist(Q, k: int = 1):
    """pi^k on the twist: (conj^k(x) xi^((p^k-1)/3), conj^k(y) xi^((p^k-1)/2))."""
    x, y = Q
    if k & 1:
        x, y = Fq2(x.c0, -x.c1), Fq2(y.c0, -y.c1)
    e = Q_MOD ** k - 1
    return (x * fq2_pow(XI, e // 3), y * fq2_pow(XI, e // 2))


def miller_loop(P, Q) -> Fq12:
    """f_{6z+2,Q}(P) * l_{[6z+2]Q, pi Q}(P) * l_{[6z+2]Q + pi Q, -pi^2 Q}(P)   (optimal ate, BN curves)."""
    assert pyref.G1.on_curve(P) and pyref.G2.on_curve(Q)
    f, T = Fq12.one(), Q
    for bit in bin(ATE_LOOP)[3:]:
        T, l = _dbl_step(T, P)
        f = f * f * l
        if bit == "1":
            T, l = _add_step(T, Q, P)
            f = f * l
    Q1 = frobenius_twist(Q, 1)
    Q2 = pyref.G2.neg(frobenius_twist(Q, 2))
    assert pyref.G2.on_curve(Q1) and pyref.G2.on_curve(Q2)
    T, l = _add_step(T, Q1, P)
    f = f * l
    T, l = _add_step(T, Q2, P)
    return f * l


FINAL_EXP = (Q_MOD ** 12 - 1) // R_MOD
# arkworks (ark-ec models/bn, after libff) and several other libraries return the pairing raised to this fixed
# extra power: their hard part follows Fuentes-Castaneda et al. and computes elt^(2z(6z^2+3z+1) (q^4-q^2+1)/r)
FUENTES_FACTOR = 2 * BN_Z * (6 * BN_Z * BN_Z + 3 * BN_Z + 1)


def pairing(P, Q, fuentes: bool = False) -> Fq12:
    """Reduced optimal-ate pairing e(P, Q) = miller_loop(P, Q)^((q^12-1)/r) (optionally the Fuentes-Castaneda multiple)."""
    if P is None or Q is None:
        return Fq12.one()
    return miller_loop(P, Q).pow(FINAL_EXP * (FUENTES_FACTOR if fuentes else 1))
