"""GPU: the reference's END-TO-END acceptance test, re-stated against the CUDA path.

groth16/examples/sha256.rs does, for one circuit: setup -> clear-text arkworks proof -> deal CRS / QAP / witness shares
and masks (:201-291) -> every party runs `dsha256` (:32-129: circom_h = 3 d_ifft + 3 d_fft + deg_red, then A, B(G1),
B(G2), C through five d_msm) -> the client unpacks the three shares (:375-377) -> the proof VERIFIES (:400-415).

Here every party-side and king-side computation of that dataflow goes through the C ABI (pack, fft1, king closures,
deg_red, MSM G1/G2, group unpack, wire format); the circuit is synthetic (the reference's sha256.r1cs is missing from
its tree), the CRS / clear-text proof / verifier are tests/groth16_ref.py over the pairing that is pinned to the
reference's vk_alphabeta_12.  Two assertions, the second being exactly the reference's:
  (1) the distributed proof equals the clear-text proof from the same (r, s), element for element;
  (2) the distributed proof verifies.
"""
import random

import numpy as np
import pytest

import groth16_ref as gr
import oracle_lib as ol
from oracle_lib import pyref

pytestmark = pytest.mark.gpu
R = pyref.R_MOD


@pytest.fixture(scope="module")
def z():
    import zksaas_b200
    return zksaas_b200


def rand_points(rnd, k, g2=False):
    return [gr.aff_to_xyz(p, g2) for p in gr.fixed_base([rnd.randrange(1, R) for _ in range(k)], g2)]


@pytest.mark.parametrize("num_constraints,num_instance,dropouts,l",
                         [(60, 2, (), 2), (60, 2, (7,), 2), (1000, 3, (), 2), (120, 2, (), 4), (120, 2, (15,), 4),
                          (32700, 2, (), 2)])        # the last: m = 2^15, the size of BASELINE configs[2] (sha256 circuit)
def test_distributed_groth16_proof_verifies(z, num_constraints, num_instance, dropouts, l):
    from zksaas_b200 import api
    rnd = random.Random(1000 * num_constraints + len(dropouts))
    rng = np.random.default_rng(num_constraints + 17 * len(dropouts))
    cs, w = gr.synthetic_circuit(num_constraints, num_instance, seed=rnd.randrange(1 << 30))
    pk, vk = gr.setup(cs, seed=rnd.randrange(1 << 30))
    r, s = rnd.randrange(R), rnd.randrange(R)
    clear = gr.prove_clear(pk, cs, w, r, s)                                      # sha256.rs:190-200 arkworks_proof
    public = w[1:num_instance]
    assert gr.verify(vk, public, clear)                                          # :400-408

    pp = z.PackedSharingParams.new(l)
    n, t = pp.n, pp.t
    net = z.LocalTestNet(n, dropouts)
    m = pk.domain_size
    dom = z.Radix2EvaluationDomain.new(m)
    mbyl = m // l
    rp = lambda cols=mbyl: ol.rand_fr(rng, cols * t)
    fr = lambda v: api.fr_image(v % R)

    # ---- dealer: QAP::pss, pack_from_arkworks_proving_key, pack_from_witness (:203-214) ----------------------
    qa, qb, qc = (ol.fr_np(v) for v in gr.qap_witness(cs, w))                    # groth16/src/qap.rs:43-90
    qap_shares = api.qap_pss(qa, qb, qc, pp, rp(), rp(), rp())
    crs = {k: api.crs_det_pack(gr.pad_to_chunks(v, l), pp, g2)
           for k, v, g2 in (("s", pk.a_query[1:], False), ("u", pk.h_query, False), ("w", pk.l_query, False),
                            ("h", pk.b_g1_query[1:], False), ("v", pk.b_g2_query[1:], True))}
    chunks = lambda k: (k + l - 1) // l
    a_shares = api.pack_from_witness(pp, ol.fr_np(w[1:]), rp(chunks(len(w) - 1)))
    ax_shares = api.pack_from_witness(pp, ol.fr_np(w[num_instance:]), rp(chunks(len(w) - num_instance)))

    # ---- masks (:218-291) ---------------------------------------------------------------------------------------
    root = z.Radix2EvaluationDomain.new(2 * m).element(1)
    one = api.fr_image(1)
    fmask = lambda rearr, g, gen: z.FftMask.sample(rearr, g, gen, m, pp, ol.rand_fr(rng, m), rp(), rp())
    fft_masks = [fmask(True, root, dom.group_gen_inv()) for _ in range(3)] + [fmask(False, one, dom.group_gen()) for _ in range(3)]
    degred_mask = z.DegRedMask.sample(pp, mbyl, ol.rand_fr(rng, mbyl * l), rp(), rp())
    g1_masks = [z.MsmMask.sample(pp, ol.rand_fr(rng, l), rand_points(rnd, t), rand_points(rnd, t)) for _ in range(4)]
    g2_mask = z.MsmMask.sample(pp, ol.rand_fr(rng, l), rand_points(rnd, t, True), rand_points(rnd, t, True), g2=True)

    # ---- circom_h (groth16/src/ext_wit.rs:104-181) ---------------------------------------------------------------
    coeff = [z.d_ifft([q[i] for q in qap_shares], fft_masks[i], True, dom, root, pp, net, rp()) for i in range(3)]
    ev = [z.d_fft(coeff[i], fft_masks[3 + i], False, dom, pp, net, rp()) for i in range(3)]
    h_eval = [api.qap_h(ev[0][p], ev[1][p], ev[2][p]) for p in range(n)]         # :173-177
    h_shares = z.deg_red(h_eval, degred_mask, pp, net, rp())                     # :179

    # ---- A, B(G1), B(G2), C (groth16/src/prove.rs; r_share = r, s_share = s: `pp.pack(vec![r; pp.n])` at
    #      sha256.rs:202 interpolates a constant polynomial) -----------------------------------------------------------
    J = gr.aff_to_xyz
    L, N, AG1, BG1 = J(pk.a_query[0]), J(pk.delta_g1), J(pk.alpha_g1), J(pk.beta_g1)
    Z1, Z2, K2, BG2 = J(pk.b_g1_query[0]), J(pk.b_g2_query[0], True), J(pk.delta_g2, True), J(pk.beta_g2, True)
    prod_a = z.d_msm(crs["s"], a_shares, g1_masks[0], pp, net)                   # prove.rs:52
    prod_b1 = z.d_msm(crs["h"], a_shares, g1_masks[1], pp, net)                  # :106
    prod_b2 = z.d_msm(crs["v"], a_shares, g2_mask, pp, net, g2=True)             # :154
    prod_w = z.d_msm(crs["w"], ax_shares, g1_masks[2], pp, net)                  # :209
    prod_u = z.d_msm(crs["u"], h_shares, g1_masks[3], pp, net)                   # :219
    A_sh = [api.group_lincomb([L, N, prod_a[p], AG1], [one, fr(r), one, one]) for p in range(n)]            # :46-57
    B1_sh = [api.group_lincomb([Z1, N, prod_b1[p], BG1], [one, fr(s), one, one]) for p in range(n)]        # :100-111
    B2_sh = [api.group_lincomb([Z2, K2, prod_b2[p], BG2], [one, fr(s), one, one], g2=True) for p in range(n)]  # :148-159
    C_sh = [api.group_lincomb([A_sh[p], B1_sh[p], N, prod_w[p], prod_u[p]], [fr(s), fr(r), fr(-(r * s)), one, one])
            for p in range(n)]                                                                            # :229-235

    # ---- client: unpack2(...)[0] (sha256.rs:375-377) --------------------------------------------------------------
    a = ol.g1_xyz_to_point(api.pss_unpack2_group(pp, A_sh)[0][0])
    b = ol.g2_xyz_to_point(api.pss_unpack2_group(pp, B2_sh, g2=True)[0][0])
    c = ol.g1_xyz_to_point(api.pss_unpack2_group(pp, C_sh)[0][0])
    assert (a, b, c) == clear                                                    # :383-388 prints both; equal for equal (r, s)
    assert gr.verify(vk, public, (a, b, c))                                      # :409-415 "Proof verification failed!"
    assert not gr.verify(vk, [(public[0] + 1) % R] + public[1:], (a, b, c))
    if dropouts:
        # the client can also lose a share: unpack_missing_shares over the rest (pss.rs:210-221)
        alive = [p for p in range(n) if p not in dropouts]
        a2 = ol.g1_xyz_to_point(api.pss_unpack2_group(pp, [A_sh[p] for p in alive], alive)[0][0])
        assert a2 == a
