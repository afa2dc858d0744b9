#!/usr/bin/env python3
"""Regenerates tests/golden/*.json from the big-integer model oracle/pyref.py.

    python tests/golden/make_golden.py

The reference tree holds no literal vectors for this path; these fixtures
are the values its own tests compute at run time, for the deterministic inputs those tests use:
  - groth16/src/ext_wit.rs:287-409 / :411-538  a = b = (0..m), c = a*b  -> libsnark_ref / circom_ref h
  - dist-primitives/examples/local_dfft_test.rs:16-24, dfft_test.rs:23-28   x = (0..m) -> dom.fft(x)
  - secret-sharing/src/pss.rs tests, dist-primitives/src/dfft/tests.rs (seeded random inputs here)
plus the derived group KATs of SURVEY.md section 8c.  Values are canonical integers in hex.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as P  # noqa: E402

R, Q = P.R_MOD, P.Q_MOD
hx = lambda v: hex(v)
pt1 = lambda p: None if p is None else [hx(p[0]), hx(p[1])]
pt2 = lambda p: None if p is None else [[hx(p[0].c0), hx(p[0].c1)], [hx(p[1].c0), hx(p[1].c1)]]


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as fh:
        json.dump(obj, fh, indent=0, separators=(",", ":"))
    print("wrote", name)


def field_golden():
    rng = random.Random(0xF1E1D)
    out = {}
    for name, p in (("fr", R), ("fq", Q)):
        a = [rng.randrange(p) for _ in range(24)] + [0, 1, p - 1, p - 2, (1 << 253), (1 << 64) - 1]
        b = [rng.randrange(p) for _ in range(24)] + [p - 1, p - 1, p - 1, 2, (1 << 253) + 5, (1 << 128) + 1]
        out[name] = {"modulus": hx(p), "a": [hx(x) for x in a], "b": [hx(x) for x in b],
                     "mul": [hx(x * y % p) for x, y in zip(a, b)], "add": [hx((x + y) % p) for x, y in zip(a, b)],
                     "sub": [hx((x - y) % p) for x, y in zip(a, b)],
                     "inv": [hx(pow(x, -1, p)) if x else hx(0) for x in a]}
    out["fr_two_adic_root"] = hx(P.FR_TWO_ADIC_ROOT)
    out["fr_roots_of_unity"] = {str(k): hx(P.Radix2Domain(1 << k).group_gen) for k in (1, 2, 3, 4, 10, 16, 20, 28)}
    dump("fields.json", out)


def group_golden():
    rng = random.Random(0x6E0)
    g1, g2 = P.G1_GEN, P.G2_GEN_PT
    ks = [1, 2, 3, 4, 5, 7, R - 1, R - 2, (1 << 253), rng.randrange(R), rng.randrange(R)]
    out = {"g1_multiples": {hx(k): pt1(P.G1.mul(g1, k)) for k in ks},
           "g2_multiples": {hx(k): pt2(P.G2.mul(g2, k)) for k in ks[:8]}}
    # MSM known answers + seeded cases incl. infinity / repeated bases / P,-P
    cases = []
    bases = [P.G1.mul(g1, k) for k in (1, 2, 3, 4)]
    sc = [5, R - 1, 0, 1 << 253]
    cases.append({"dlogs": [hx(k) for k in (1, 2, 3, 4)], "scalars": [hx(s) for s in sc], "result": pt1(P.msm_naive(P.G1, bases, sc))})
    for n in (1, 7, 33, 64):
        dl = [rng.randrange(R) for _ in range(n)]
        s = [rng.randrange(R) for _ in range(n)]
        if n >= 7:
            dl[1] = dl[0]; s[1] = s[0]            # P + P in one bucket
            dl[3] = R - dl[2]; s[3] = s[2]        # P and -P
            s[4] = 0; s[5] = R - 1; dl[6] = 0     # zero scalar, r-1, infinity base
        total = sum(a * b for a, b in zip(dl, s)) % R
        cases.append({"dlogs": [hx(k) for k in dl], "scalars": [hx(x) for x in s], "result": pt1(P.G1.mul(g1, total))})
    out["g1_msm"] = cases
    cases2 = []
    for n in (1, 5, 17):
        dl = [rng.randrange(R) for _ in range(n)]
        s = [rng.randrange(R) for _ in range(n)]
        if n >= 5:
            dl[1] = dl[0]; s[1] = s[0]; s[2] = 0; dl[3] = 0
        total = sum(a * b for a, b in zip(dl, s)) % R
        cases2.append({"dlogs": [hx(k) for k in dl], "scalars": [hx(x) for x in s], "result": pt2(P.G2.mul(g2, total))})
    out["g2_msm"] = cases2
    dump("groups.json", out)


def pss_golden():
    rng = random.Random(0x955)
    out = {}
    for l in (2, 4):
        pp = P.PackedSharingParams(l)
        sec = [rng.randrange(R) for _ in range(l)]
        rnd = [rng.randrange(R) for _ in range(l)]
        sh = pp.pack(sec, rnd)
        sq = [x * x % R for x in sh]
        out[str(l)] = {
            "pack_matrix": [[hx(v) for v in row] for row in pp.pack_matrix()],
            "unpack_matrix": [[hx(v) for v in row] for row in pp.unpack_matrix()],
            "unpack2_matrix": [[hx(v) for v in row] for row in pp.unpack2_matrix()],
            "secrets": [hx(v) for v in sec], "rand": [hx(v) for v in rnd], "shares": [hx(v) for v in sh],
            "det_shares": [hx(v) for v in pp.det_pack(sec)],
            "squared_shares_unpack2": [hx(v) for v in pp.unpack2(sq)],
            "lagrange_missing_last": [hx(v) for v in pp.lagrange_unpack(sq[:-1], list(range(pp.n - 1)))],
        }
    dump("pss.json", out)


def dfft_golden():
    rng = random.Random(0xDF7)
    out = {}
    for l, m in ((2, 8), (2, 64), (4, 32)):
        pp = P.PackedSharingParams(l)
        dom = P.Radix2Domain(m)
        mbyl = m // l
        x = list(range(m))                                     # local_dfft_test.rs:16-21
        fftx = dom.fft(x)
        xr = P.fft_in_place_rearrange(x)
        rand0 = [[rng.randrange(R) for _ in range(pp.t)] for _ in range(mbyl)]
        packed = P.transpose([pp.pack([xr[i + j * mbyl] for j in range(l)], rand0[i]) for i in range(mbyl)])
        fft1 = [P.fft1_in_place(v, pp, dom.group_gen) for v in packed]
        rand1 = [[rng.randrange(R) for _ in range(pp.t)] for _ in range(mbyl)]
        zeta = P.Radix2Domain(2 * m).element(1)
        king = {}
        for rearr in (0, 1):
            for gname, g in (("one", 1), ("zeta_2m", zeta)):
                king[f"rearrange{rearr}_{gname}"] = [[hx(v) for v in row] for row in
                                                     P.king_fft2(fft1, list(range(pp.n)), pp, dom.group_gen, g, rearr, rand1)]
        s1 = [rng.randrange(R) for _ in range(m)]
        out[f"l{l}_m{m}"] = {
            "x": "0..m", "fft_x": [hx(v) for v in fftx], "ifft_x": [hx(v) for v in dom.ifft(list(range(m)))],
            "rand_pack": [[hx(v) for v in r] for r in rand0], "party_shares": [[hx(v) for v in row] for row in packed],
            "fft1": [[hx(v) for v in row] for row in fft1], "rand_king": [[hx(v) for v in r] for r in rand1],
            "king": king, "fft2_in": [hx(v) for v in s1],
            "fft2_out": [hx(v) for v in P.fft2_in_place(s1, pp, dom.group_gen)],
        }
    dump("dfft.json", out)


def ext_wit_golden():
    """groth16/src/ext_wit.rs:204-285 (libsnark_ref, circom_ref) on the tests' deterministic inputs."""
    out = {}
    for m in (32, 1024):
        dom = P.Radix2Domain(m)
        a = list(range(m)); b = list(range(m)); c = [x * y % R for x, y in zip(a, b)]
        # circom_ref :239-285
        root = P.Radix2Domain(2 * m).element(1)
        ac, bc, cc = dom.ifft(a), dom.ifft(b), dom.ifft(c)
        ae, be, ce = (dom.fft(P.distribute_powers(v, root)) for v in (ac, bc, cc))
        circom = [(x * y - z) % R for x, y, z in zip(ae, be, ce)]
        # libsnark_ref :204-237
        coset = dom.get_coset(P.FR_GENERATOR)
        ae, be, ce = coset.fft(ac), coset.fft(bc), coset.fft(cc)
        vanish_inv = pow((pow(P.FR_GENERATOR, m, R) - 1) % R, -1, R)      # evaluate_vanishing_polynomial(g)^-1
        lib = coset.ifft([(x * y - z) * vanish_inv % R for x, y, z in zip(ae, be, ce)])
        out[str(m)] = {"circom_h": [hx(v) for v in circom], "libsnark_h": [hx(v) for v in lib]}
    dump("ext_wit.json", out)


if __name__ == "__main__":
    field_golden()
    group_golden()
    pss_golden()
    dfft_golden()
    ext_wit_golden()
