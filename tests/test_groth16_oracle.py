"""CPU: the reference's end-to-end acceptance test (groth16/examples/sha256.rs: the distributed proof VERIFIES,
:400-415) run over the ORACLE's restatements of the path -- pyref's d_fft / d_ifft / deg_red / pack / unpack2 and the C
oracle's MSM and group packing -- with the clear-text Groth16 model of tests/groth16_ref.py and the pairing that is
pinned to the reference's vk_alphabeta_12.  It ties the oracle's PROTOCOL-level functions (not only its field and curve
arithmetic) to the property the reference itself tests, and it is the CPU twin of tests/test_gpu_groth16.py.
The FFT, degree-reduction and MSM masks are SAMPLED as the reference samples them (dfft/mod.rs:30-85, deg_red.rs:40-66,
dmsm/mod.rs:21-48), so the test also checks that they cancel."""
import random

import numpy as np
import pytest

import groth16_ref as gr
import oracle_lib as ol
from oracle_lib import _p, pyref

R = pyref.R_MOD


def test_clear_text_groth16_model_verifies():
    cs, w = gr.synthetic_circuit(60, 2, seed=7)
    pk, vk = gr.setup(cs, seed=11)
    proof = gr.prove_clear(pk, cs, w, 12345, 67890)
    assert gr.verify(vk, w[1:2], proof)
    assert not gr.verify(vk, [(w[1] + 1) % R], proof)                   # wrong public input
    A, B, C = proof
    assert not gr.verify(vk, w[1:2], (A, B, pyref.G1.add(C, A)))        # tampered proof
    assert not gr.verify(vk, w[1:2], (A, pyref.G2.add(B, B), C))


@pytest.mark.parametrize("l,dropouts", [(2, ()), (4, ()), (2, (7,))])
def test_oracle_distributed_groth16_proof_verifies(l, dropouts):
    o = ol.oracle()
    rnd = random.Random(20260)
    cs, w = gr.synthetic_circuit(28, 2, seed=5)
    pk, vk = gr.setup(cs, seed=6)
    r, s = rnd.randrange(R), rnd.randrange(R)
    clear = gr.prove_clear(pk, cs, w, r, s)
    pp = pyref.PackedSharingParams(l)
    n, t = pp.n, pp.t
    m = pk.domain_size
    mbyl = m // l
    rand_cols = lambda cols=mbyl: [[rnd.randrange(R) for _ in range(t)] for _ in range(cols)]
    parties = [q for q in range(n) if q not in dropouts]      # whose messages reach the king (lossy round: multi.rs:330-363)

    def qap_pss_pack(x):                                                # groth16/src/qap.rs:99-112
        x = pyref.fft_in_place_rearrange(x)
        rc = rand_cols()
        return pyref.transpose([pp.pack(x[i::mbyl][:l], rc[i]) for i in range(mbyl)])

    def pack_from_witness(v):                                           # sha256.rs:131-156
        v = list(v) + [0] * ((-len(v)) % l)
        return pyref.transpose(pyref.pack_vec(v, pp, rand_cols(len(v) // l)))

    def crs_det_pack(bases, g2):                                        # groth16/src/proving_key.rs:72-104
        words = 24 if g2 else 12
        bases = gr.pad_to_chunks(bases, l)
        out = [np.zeros((bases.shape[0] // l, bases.shape[1]), dtype=np.uint8) for _ in range(n)]
        for j in range(bases.shape[0] // l):
            sec = np.concatenate([gr.aff_to_xyz(bases[j * l + k], g2) for k in range(l)])
            sh = np.zeros(n * words, dtype=np.uint64)
            (o.zko_pss_pack_g2 if g2 else o.zko_pss_pack_g1)(l, _p(sec), None, _p(sh))
            for p in range(n):
                out[p][j] = gr.xyz_to_aff(sh[p * words:(p + 1) * words], g2)
        return out

    def gadd(a, b, g2=False):
        out = np.zeros_like(a)
        (o.zko_g2_add if g2 else o.zko_g1_add)(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
        return out

    def gmul(a, k, g2=False):
        out = np.zeros_like(a)
        (o.zko_g2_mul if g2 else o.zko_g1_mul)(_p(np.ascontiguousarray(a)), _p(ol.fr_np([k % R])), _p(out))
        return out

    def unpack2(shares, g2=False, who=None):                            # pss.rs:141-166 / :210-221 over group elements
        words = 24 if g2 else 12
        if who is not None and len(who) < n:                            # unpack_missing_shares -> lagrange_unpack (:170-207)
            curve = pyref.G2 if g2 else pyref.G1
            to_pt, to_img = (ol.g2_xyz_to_point, ol.g2_point_to_xyz) if g2 else (ol.g1_xyz_to_point, ol.g1_point_to_xyz)
            pts = pp.unpack_missing_shares([to_pt(x) for x in shares], list(who), pyref.group_ops(curve))
            return [to_img(q) for q in pts]
        u = np.zeros(l * words, dtype=np.uint64)
        (o.zko_pss_unpack2_g2 if g2 else o.zko_pss_unpack2_g1)(l, _p(np.concatenate(shares)), _p(u))
        return [u[i * words:(i + 1) * words].copy() for i in range(l)]

    def msm_mask_sample(g2=False):                                      # MsmMask::sample, dmsm/mod.rs:21-48
        words = 24 if g2 else 12
        gen = gr.aff_to_xyz(gr.fixed_base([1], g2)[0], g2)
        rand_pts = lambda: np.concatenate([gr.aff_to_xyz(a, g2) for a in gr.fixed_base([rnd.randrange(1, R) for _ in range(t)], g2)])
        values = [gmul(gen, rnd.randrange(R), g2) for _ in range(l)]                        # :27-31 gen * x_i
        total = values[0]
        for v in values[1:]:
            total = gadd(total, v, g2)
        out_value = gmul(total, R - 1, g2)                                                  # :38 -(sum of the mask values)
        pack = o.zko_pss_pack_g2 if g2 else o.zko_pss_pack_g1
        ins, outs = np.zeros(n * words, dtype=np.uint64), np.zeros(n * words, dtype=np.uint64)
        pack(l, _p(np.concatenate(values)), _p(rand_pts()), _p(ins))                        # :34 pack(mask values)
        pack(l, _p(np.concatenate([out_value] * l)), _p(rand_pts()), _p(outs))              # :40-41 pack([out; l])
        return [(ins[p * words:(p + 1) * words].copy(), outs[p * words:(p + 1) * words].copy()) for p in range(n)]

    def d_msm(bases_by_party, scalars_by_party, g2=False):              # dmsm/mod.rs:59-102 with sampled masks
        msm = ol.o_g2_msm if g2 else ol.o_g1_msm
        masks = msm_mask_sample(g2)
        c = [gadd(msm(bases_by_party[p], ol.fr_np(scalars_by_party[p])), masks[p][0], g2) for p in range(n)]   # :73-74
        res = unpack2([c[q] for q in parties], g2, parties)                                 # :85
        out = res[0]
        for x in res[1:]:
            out = gadd(out, x, g2)                                                          # :86
        return [gadd(out, masks[p][1], g2) for p in range(n)]                               # :87, :98

    # dealer
    qa, qb, qc = (qap_pss_pack(v) for v in gr.qap_witness(cs, w))
    crs = {k: crs_det_pack(v, g2) for k, v, g2 in (("s", pk.a_query[1:], False), ("u", pk.h_query, False),
                                                    ("w", pk.l_query, False), ("h", pk.b_g1_query[1:], False),
                                                    ("v", pk.b_g2_query[1:], True))}
    a_sh, ax_sh = pack_from_witness(w[1:]), pack_from_witness(w[cs.num_instance:])
    # circom_h (groth16/src/ext_wit.rs:104-181)
    dom = pyref.Radix2Domain(m)
    root = pyref.Radix2Domain(2 * m).element(1)
    fmask = lambda rearr, g_, gen_: pyref.fft_mask_sample(rearr, g_, gen_, m, pp, [rnd.randrange(R) for _ in range(m)], rand_cols(), rand_cols())
    im = [fmask(True, root, dom.group_gen_inv) for _ in range(3)]                                   # sha256.rs:218-267
    fm = [fmask(False, 1, dom.group_gen) for _ in range(3)]
    coeff = [pyref.d_fft_round(q, im[k][0], im[k][1], True, m, pp, rand_cols(), inverse=True, g=root, parties=parties) for k, q in enumerate((qa, qb, qc))]
    ev = [pyref.d_fft_round(cf, fm[k][0], fm[k][1], False, m, pp, rand_cols(), parties=parties) for k, cf in enumerate(coeff)]
    h_eval = [[(x * y - v) % R for x, y, v in zip(ev[0][p], ev[1][p], ev[2][p])] for p in range(n)]
    dm_in, dm_out = pyref.deg_red_mask_sample(pp, mbyl, [rnd.randrange(R) for _ in range(mbyl * l)], rand_cols(), rand_cols())
    masked = [[(x + k) % R for x, k in zip(h_eval[p], dm_in[p])] for p in range(n)]                 # deg_red.rs:94-97
    h_sh = pyref.deg_red_king([masked[q] for q in parties], parties, pp, rand_cols())
    h_sh = [[(x + k) % R for x, k in zip(h_sh[p], dm_out[p])] for p in range(n)]                    # :120-124
    # the shares of h unpack to circom_ref's h (ext_wit.rs:532-537)
    got_h = sum((pp.unpack(col) for col in pyref.transpose(h_sh)), [])
    assert got_h == gr.circom_h(*gr.qap_witness(cs, w))
    # A, B, C (groth16/src/prove.rs)
    J = gr.aff_to_xyz
    L, N, AG1, BG1 = J(pk.a_query[0]), J(pk.delta_g1), J(pk.alpha_g1), J(pk.beta_g1)
    Z1, Z2, K2, BG2 = J(pk.b_g1_query[0]), J(pk.b_g2_query[0], True), J(pk.delta_g2, True), J(pk.beta_g2, True)
    prod_a, prod_b1 = d_msm(crs["s"], a_sh), d_msm(crs["h"], a_sh)
    prod_b2 = d_msm(crs["v"], a_sh, True)
    prod_w, prod_u = d_msm(crs["w"], ax_sh), d_msm(crs["u"], h_sh)
    A_sh = [gadd(gadd(gadd(L, gmul(N, r)), prod_a[p]), AG1) for p in range(n)]
    B1_sh = [gadd(gadd(gadd(Z1, gmul(N, s)), prod_b1[p]), BG1) for p in range(n)]
    B2_sh = [gadd(gadd(gadd(Z2, gmul(K2, s, True), True), prod_b2[p], True), BG2, True) for p in range(n)]
    C_sh = [gadd(gadd(gadd(gadd(gmul(A_sh[p], s), gmul(B1_sh[p], r)), gmul(N, -(r * s))), prod_w[p]), prod_u[p])
            for p in range(n)]
    proof = (ol.g1_xyz_to_point(unpack2(A_sh)[0]), ol.g2_xyz_to_point(unpack2(B2_sh, True)[0]),
             ol.g1_xyz_to_point(unpack2(C_sh)[0]))                       # sha256.rs:375-377
    assert proof == clear
    assert gr.verify(vk, w[1:cs.num_instance], proof)                   # sha256.rs:409-415


def test_oracle_libsnark_h_dataflow_vs_golden():
    """groth16/src/ext_wit.rs:14-102 `libsnark_h` + its test :287-409 over pyref's d_ifft / d_fft rounds (m = 32, zero
    masks): the unpack2'ed coefficients equal libsnark_ref's h (tests/golden/ext_wit.json) -- the CPU twin of
    tests/test_gpu_protocol.py::test_libsnark_h_dataflow_vs_golden."""
    import golden_util as gu
    m, l = 32, 2
    pp = pyref.PackedSharingParams(l)
    n, mbyl = pp.n, m // l
    rnd = random.Random(1)
    rc = lambda: [[rnd.randrange(R) for _ in range(pp.t)] for _ in range(mbyl)]
    a = list(range(m))
    c = [x * x % R for x in a]

    def pss(x):
        x = pyref.fft_in_place_rearrange(x)
        r = rc()
        return pyref.transpose([pp.pack(x[i::mbyl][:l], r[i]) for i in range(mbyl)])

    zero = [[0] * mbyl for _ in range(n)]
    sh = {k: pss(v) for k, v in (("a", a), ("b", a), ("c", c))}
    g, ginv = 5, pow(5, -1, R)
    coeff = {k: pyref.d_fft_round(sh[k], zero, zero, True, m, pp, rc(), inverse=True, g=g) for k in "abc"}
    ev = {k: pyref.d_fft_round(coeff[k], zero, zero, True, m, pp, rc()) for k in "abc"}
    vinv = pow((pow(5, m, R) - 1) % R, -1, R)
    h_eval = [[(x * y - w) * vinv % R for x, y, w in zip(ev["a"][p], ev["b"][p], ev["c"][p])] for p in range(n)]
    h = pyref.d_fft_round(h_eval, zero, zero, False, m, pp, rc(), inverse=True, g=ginv)
    got = sum((pp.unpack2(col) for col in pyref.transpose(h)), [])
    assert got == gu.ints(gu.load("ext_wit.json")[str(m)]["libsnark_h"])
