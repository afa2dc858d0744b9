"""tests/groth16_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Clear-text Groth16 over BN254 in the CircomReduction flavour the reference proves with
(groth16/examples/sha256.rs:170-200: `Groth16::<Bn254, CircomReduction>::circuit_specific_setup`,
`create_proof_with_reduction_and_matrices`; :400-415 `verify_with_processed_vk`), so that the end-to-end
test can do what that example does: deal a CRS to the distributed prover, compute the same proof in the
clear from the same (r, s), and run the verifier -- over oracle/pairing.py, whose pairing is pinned to the
reference's `vk_alphabeta_12` (fixtures/verification_key.json:52-81).

The setup / prover / verifier arithmetic lives in dependencies that are NOT under /root/reference
(ark-groth16 ^0.4, ark-relations ^0.4, ark-circom: an un-pinned git dependency, groth16/Cargo.toml:16);
restated from their public behaviour:
  ark-groth16  generator.rs    generate_parameters_with_qap  (queries a_i(tau) G1, b_i(tau) G1|G2,
                               (beta a_i + alpha b_i + c_i)/gamma|delta G1, random group generators)
               r1cs_to_qap.rs  instance_map_with_evaluation   (domain = constraints + instance variables; the
                               extra rows a[num_constraints + i] = x_i that groth16/src/qap.rs:73-77 mirrors)
               prover.rs       create_proof_with_assignment   (A, B, C from the queries, r and s)
               verifier.rs     e(A, B) = e(alpha, beta) e(sum x_i IC_i, gamma) e(C, delta)
  ark-circom   circom/qap.rs   CircomReduction: h_i = (ab - c)(w_2m^(2i+1)) with c = a*b on the constraint rows
                               (what groth16/src/qap.rs:79-86 and ext_wit.rs:239-285 `circom_ref` restate in-tree);
                               h_query_i = delta^-1 * [ifft_2m(tau^0 .. tau^(2m-2))]_(2i+1)
The in-tree halves (qap.rs, circom_ref) are followed literally; the out-of-tree halves are checked by the only
property that matters to the reference's own end-to-end test: the proof VERIFIES.

The circuit is synthetic (fixtures/sha256/sha256.r1cs is missing from the reference tree and there is no circom
toolchain here): random rank-1 constraints over a growing variable set, satisfied by construction.
"""
from __future__ import annotations

import random

import numpy as np

import oracle_lib as ol
from oracle_lib import _p, pyref

R = pyref.R_MOD
Q = pyref.Q_MOD


# ---------------------------------------------------------------------------------------------------
# circuit
# ---------------------------------------------------------------------------------------------------
class R1CS:
    """ark_relations ConstraintMatrices: rows of (coeff, variable index); variable 0 is the constant 1, the next
    num_instance - 1 are the public inputs, the rest the witness."""

    def __init__(self, num_instance, a, b, c, num_vars):
        self.num_instance, self.a, self.b, self.c, self.num_vars = num_instance, a, b, c, num_vars
        self.num_constraints = len(a)


def _dot(row, z):
    return sum(co * z[i] for co, i in row) % R


def synthetic_circuit(num_constraints, num_instance, seed):
    """A satisfiable R1CS and its full assignment: every constraint multiplies two random linear combinations of the
    variables defined so far and defines one new witness variable through its C row (one in four C rows has two
    terms, one in eight B rows is the constant 1, i.e. a linear constraint)."""
    rnd = random.Random(seed)
    z = [1] + [rnd.randrange(R) for _ in range(num_instance - 1)]
    A, B, C = [], [], []

    def lc():
        return [(rnd.randrange(1, R) if rnd.random() < 0.5 else rnd.randrange(1, 16), rnd.randrange(len(z)))
                for _ in range(rnd.randrange(1, 4))]

    for j in range(num_constraints):
        ra = lc()
        rb = [(1, 0)] if j % 8 == 5 else lc()
        prod = _dot(ra, z) * _dot(rb, z) % R
        k = len(z)
        if j % 4 == 3:
            c1, c2, k2 = rnd.randrange(1, R), rnd.randrange(1, R), rnd.randrange(1, k)
            z.append((prod - c2 * z[k2]) * pow(c1, -1, R) % R)
            C.append([(c1, k), (c2, k2)])
        else:
            z.append(prod)
            C.append([(1, k)])
        A.append(ra)
        B.append(rb)
    cs = R1CS(num_instance, A, B, C, len(z))
    assert all(_dot(cs.a[j], z) * _dot(cs.b[j], z) % R == _dot(cs.c[j], z) for j in range(num_constraints))
    return cs, z


def qap_witness(cs: R1CS, z):
    """groth16/src/qap.rs:43-90 `qap()`: a, b over the domain of num_constraints + num_inputs points (zero-padded),
    a[num_constraints + i] = z[i] for the instance variables, c = a*b on the constraint rows."""
    m = pyref.Radix2Domain(cs.num_constraints + cs.num_instance).size
    a, b, c = [0] * m, [0] * m, [0] * m
    for j in range(cs.num_constraints):
        a[j], b[j] = _dot(cs.a[j], z), _dot(cs.b[j], z)
        c[j] = a[j] * b[j] % R
    for i in range(cs.num_instance):
        a[cs.num_constraints + i] = z[i]
    return a, b, c


def circom_h(a, b, c):
    """groth16/src/ext_wit.rs:239-285 `circom_ref` (= ark-circom's witness map): (ab - c) on the coset w_2m * H."""
    m = len(a)
    dom = pyref.Radix2Domain(m)
    root = pyref.Radix2Domain(2 * m).element(1)
    ev = [dom.fft(pyref.distribute_powers(dom.ifft(v), root)) for v in (a, b, c)]
    return [(x * y - w) % R for x, y, w in zip(*ev)]


# ---------------------------------------------------------------------------------------------------
# points: arkworks affine images <-> the oracle's int tuples
# ---------------------------------------------------------------------------------------------------
def _fq(img8):
    return pyref.from_mont_limbs(np.frombuffer(bytes(img8), dtype=np.uint64), Q)


def g1_point(img):
    img = np.asarray(img, dtype=np.uint8)
    return None if img[64] else (_fq(img[0:32]), _fq(img[32:64]))


def g2_point(img):
    img = np.asarray(img, dtype=np.uint8)
    if img[128]:
        return None
    return (pyref.Fq2(_fq(img[0:32]), _fq(img[32:64])), pyref.Fq2(_fq(img[64:96]), _fq(img[96:128])))


_ONE_FQ = np.array(pyref.to_mont_limbs(1, Q), dtype=np.uint64)


def aff_to_xyz(img, g2=False):
    """arkworks Affine image -> normalised Jacobian image (identity = (1, 1, 0))."""
    w = 8 if g2 else 4
    img = np.asarray(img, dtype=np.uint8)
    out = np.zeros(3 * w, dtype=np.uint64)
    if img[16 * w]:
        out[:4] = _ONE_FQ
        out[w:w + 4] = _ONE_FQ
    else:
        out[:2 * w] = np.frombuffer(img[:16 * w].tobytes(), dtype=np.uint64)
        out[2 * w:2 * w + 4] = _ONE_FQ
    return out


def xyz_to_aff(xyz, g2=False):
    """normalised Jacobian image -> arkworks Affine image."""
    w = 8 if g2 else 4
    xyz = np.asarray(xyz, dtype=np.uint64).reshape(3 * w)
    out = np.zeros(136 if g2 else 72, dtype=np.uint8)
    if not xyz[2 * w:].any():
        out[16 * w] = 1
    else:
        out[:16 * w] = np.frombuffer(xyz[:2 * w].tobytes(), dtype=np.uint8)
    return out


def pad_to_chunks(bases, l):
    """cfg_chunks!(query, pp.l) + det_pack of a short last chunk (groth16/src/proving_key.rs:72-86): the missing
    secrets are the zero-padding of ifft_in_place, i.e. the identity."""
    k = (-bases.shape[0]) % l
    if not k:
        return bases
    pad = np.zeros((k, bases.shape[1]), dtype=np.uint8)
    pad[:, 64 if bases.shape[1] == 72 else 128] = 1
    return np.concatenate([bases, pad])


def fixed_base(scalars, g2=False):
    """[s * G for s in scalars] as arkworks affine images, through the C oracle's fixed-base routine."""
    o = ol.oracle()
    stride = 136 if g2 else 72
    out = np.zeros((len(scalars), stride), dtype=np.uint8)
    if len(scalars):
        sc = ol.fr_np([s % R for s in scalars])
        (o.zko_g2_fixed_base if g2 else o.zko_g1_fixed_base)(_p(sc), len(scalars), out.ctypes.data, stride)
    return out


# ---------------------------------------------------------------------------------------------------
# setup
# ---------------------------------------------------------------------------------------------------
class ProvingKey:
    """ark_groth16::ProvingKey as affine images ((k, 72) / (k, 136) uint8), field names as in arkworks."""


def setup(cs: R1CS, seed):
    """generate_random_parameters_with_reduction::<CircomReduction>: returns (pk, vk); vk = dict of int-tuple points."""
    rnd = random.Random(seed)
    tau, alpha, beta, gamma, delta, rho1, rho2 = (rnd.randrange(1, R) for _ in range(7))
    dom = pyref.Radix2Domain(cs.num_constraints + cs.num_instance)
    m, nc, ni, nv = dom.size, cs.num_constraints, cs.num_instance, cs.num_vars
    # evaluate_all_lagrange_coefficients(tau): L_j(tau) = (tau^m - 1) / (m (tau - w^j)) * w^j
    zt = (pow(tau, m, R) - 1) % R
    u = [zt * dom.size_inv % R * dom.element(j) % R * pow((tau - dom.element(j)) % R, -1, R) % R for j in range(m)]
    a, b, c = [0] * nv, [0] * nv, [0] * nv
    for i in range(ni):                                    # instance_map_with_evaluation: a[i] = u[start + i]
        a[i] = u[nc + i]
    for j in range(nc):
        for co, i in cs.a[j]:
            a[i] = (a[i] + u[j] * co) % R
        for co, i in cs.b[j]:
            b[i] = (b[i] + u[j] * co) % R
        for co, i in cs.c[j]:
            c[i] = (c[i] + u[j] * co) % R
    ginv, dinv = pow(gamma, -1, R), pow(delta, -1, R)
    abc = [(beta * a[i] + alpha * b[i] + c[i]) % R for i in range(nv)]
    # CircomReduction::h_query_scalars(max_power = m - 1): delta^-1 tau^i, i < 2m - 1, ifft over the 2m-domain, odd entries
    dom2 = pyref.Radix2Domain(2 * m)
    hs = dom2.ifft([dinv * pow(tau, i, R) % R for i in range(2 * m - 1)])[1::2]
    pk = ProvingKey()
    s1 = lambda v: [x * rho1 % R for x in v]               # g1 = rho1 * G, g2 = rho2 * G2 (arkworks draws random generators)
    s2 = lambda v: [x * rho2 % R for x in v]
    pk.a_query, pk.b_g1_query, pk.b_g2_query = fixed_base(s1(a)), fixed_base(s1(b)), fixed_base(s2(b), True)
    pk.l_query = fixed_base(s1([x * dinv % R for x in abc[ni:]]))
    pk.h_query = fixed_base(s1(hs))
    pk.alpha_g1, pk.beta_g1, pk.delta_g1 = fixed_base(s1([alpha, beta, delta]))
    pk.beta_g2, pk.delta_g2, gamma_g2 = fixed_base(s2([beta, delta, gamma]), True)
    gamma_abc = fixed_base(s1([x * ginv % R for x in abc[:ni]]))
    vk = {"alpha_g1": g1_point(pk.alpha_g1), "beta_g2": g2_point(pk.beta_g2), "gamma_g2": g2_point(gamma_g2),
          "delta_g2": g2_point(pk.delta_g2), "gamma_abc_g1": [g1_point(p) for p in gamma_abc]}
    pk.domain_size = m
    return pk, vk


# ---------------------------------------------------------------------------------------------------
# clear-text prover and the verifier
# ---------------------------------------------------------------------------------------------------
def _msm1(bases, scalars):
    return ol.g1_xyz_to_point(ol.o_g1_msm(bases, ol.fr_np(scalars))) if len(scalars) else None


def _msm2(bases, scalars):
    return ol.g2_xyz_to_point(ol.o_g2_msm(bases, ol.fr_np(scalars))) if len(scalars) else None


def prove_clear(pk: ProvingKey, cs: R1CS, z, r, s):
    """create_proof_with_reduction_and_matrices (ark-groth16 prover.rs) -> (A, B, C) as int-tuple points."""
    G1, G2 = pyref.G1, pyref.G2
    h = circom_h(*qap_witness(cs, z))
    alpha, beta1, delta1 = g1_point(pk.alpha_g1), g1_point(pk.beta_g1), g1_point(pk.delta_g1)
    beta2, delta2 = g2_point(pk.beta_g2), g2_point(pk.delta_g2)
    A = G1.add(G1.add(alpha, _msm1(pk.a_query, z)), G1.mul(delta1, r))
    B1 = G1.add(G1.add(beta1, _msm1(pk.b_g1_query, z)), G1.mul(delta1, s))
    B2 = G2.add(G2.add(beta2, _msm2(pk.b_g2_query, z)), G2.mul(delta2, s))
    C = G1.add(_msm1(pk.l_query, z[cs.num_instance:]), _msm1(pk.h_query, h))
    C = G1.add(C, G1.add(G1.mul(A, s), G1.mul(B1, r)))
    C = G1.add(C, G1.neg(G1.mul(delta1, r * s % R)))
    return A, B2, C


def verify(vk, public_inputs, proof):
    """verify_proof (ark-groth16 verifier.rs): e(A, B) e(-alpha, beta) e(-IC, gamma) e(-C, delta) == 1."""
    import pairing
    A, B, C = proof
    if A is None or B is None or C is None:
        return False
    G1 = pyref.G1
    if not (G1.on_curve(A) and pyref.G2.on_curve(B) and G1.on_curve(C)):
        return False
    ic = vk["gamma_abc_g1"]
    assert len(public_inputs) + 1 == len(ic)
    acc = ic[0]
    for x, P in zip(public_inputs, ic[1:]):
        acc = G1.add(acc, G1.mul(P, x % R))
    f = pairing.miller_loop(A, B)
    for P, Qp in ((vk["alpha_g1"], vk["beta_g2"]), (acc, vk["gamma_g2"]), (C, vk["delta_g2"])):
        if P is not None:
            f = f * pairing.miller_loop(G1.neg(P), Qp)
    return f.pow(pairing.FINAL_EXP) == pairing.Fq12.one()
