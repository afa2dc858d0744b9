"""CPU: the multi-GPU host logic (point-range sharding + one all-gather + combine) on a world of 2
gloo processes, with the CPU oracle standing in for the per-rank CUDA MSM."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_every_point_once():
    from zksaas_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000, (1 << 22) + 3):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                cover.append((lo, hi))
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cover]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as ol
    from oracle_lib import _p
    from zksaas_b200.sharding import sharded_msm, xyz_to_xyzz_g1
    o = ol.oracle()
    rng = np.random.default_rng(123)                       # same data on every rank
    bases = np.zeros((n, 72), dtype=np.uint8)
    if n:
        o.zko_g1_sequence(_p(ol.rand_fr(rng, 1)), _p(ol.rand_fr(rng, 1)), n, bases.ctypes.data, 72)
    scalars = ol.rand_fr(rng, n)

    def partial(lo, hi):
        return xyz_to_xyzz_g1(ol.o_g1_msm(bases[lo:hi], scalars[lo:hi])) if hi > lo else np.zeros(16, dtype=np.uint64)

    def all_gather(mine):
        t = torch.from_numpy(mine.view(np.int64).copy())
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [x.numpy().view(np.uint64) for x in outs]

    def combine(parts):
        acc = ol.g1_point_to_xyz(None)
        for p in parts:
            if p[8:12].any():                               # zz != 0
                xyz = np.concatenate([p[0:8], p[8:12]])
                out = np.zeros(12, dtype=np.uint64)
                o.zko_g1_add(_p(acc), _p(np.ascontiguousarray(xyz)), _p(out))
                acc = out
        return acc

    got = sharded_msm(n, world, rank, partial, all_gather, combine)
    full = ol.o_g1_msm(bases, scalars) if n else ol.g1_point_to_xyz(None)
    q.put((rank, bool((got == full).all())))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 301])
def test_sharded_msm_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------------------------------------
# sharded king pipeline: host logic (column ranges + ONE sum reduce-scatter) with the big-integer
# model standing in for the two CUDA stages
# ------------------------------------------------------------------------------------------------
def _king_worker(rank, world, port, l, m, rearrange, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    import oracle_lib as ol
    from oracle_lib import pyref
    from zksaas_b200.sharding import king_sharded
    R = pyref.R_MOD
    rng = random.Random(77)                                  # same inputs on every rank
    pp = pyref.PackedSharingParams(l)
    dom = pyref.Radix2Domain(m)
    mbyl = m // l
    gen, g = dom.group_gen, pyref.Radix2Domain(2 * m).element(1)
    cols = [pp.pack([rng.randrange(R) for _ in range(l)], [rng.randrange(R) for _ in range(l)]) for _ in range(mbyl)]
    shares = pyref.transpose([[v * v % R for v in c] for c in cols])          # party-major, degree 2(l+t)-2
    rand = [[rng.randrange(R) for _ in range(pp.t)] for _ in range(mbyl)]
    expect = pyref.king_fft2(shares, list(range(pp.n)), pp, gen, g, rearrange, rand)
    log_m = m.bit_length() - 1

    def brev(x):
        return int(format(x, f"0{log_m}b")[::-1], 2)

    def stage1(lo, hi):
        """what k_king_stage1 computes for columns [lo, hi): column-local fft2 closed form, pack-order scatter"""
        S = [0] * m
        for k in range(lo, hi):
            ent = {k: pp.unpack2([shares[p][k] for p in range(pp.n)])}
            C, E = mbyl, l
            for i in range(l.bit_length() - 1, 0, -1):
                new = {}
                for kap, e in ent.items():
                    tw = pow(gen, (1 << (i - 1)) * (kap + 1), R)
                    new[kap] = [(e[2 * j] + e[2 * j + 1] * tw) % R for j in range(E // 2)]
                    new[kap + C] = [(e[2 * j] - e[2 * j + 1] * tw) % R for j in range(E // 2)]
                ent, C, E = new, C * 2, E // 2
            for kap, e in ent.items():
                pos = (kap + 1) % m
                v = e[0] * pow(g, pos, R) % R
                if rearrange:
                    p = brev(pos)
                    S[(p % mbyl) * l + p // mbyl] = v
                else:
                    S[pos] = v
        return torch.from_numpy(ol.fr_np(S).view(np.int64).copy())

    def reduce_scatter(S):
        # gloo has no reduce_scatter_tensor: all_reduce + slice has the same semantics
        t = S.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t[rank * (m // world):(rank + 1) * (m // world)]

    def stage2(lo, hi, S_r):
        vals = ol.np_fr(S_r.numpy().view(np.uint64))
        out = [pp.pack(vals[c * l:(c + 1) * l], rand[lo + c]) for c in range(hi - lo)]
        return pyref.transpose(out)

    got = king_sharded(mbyl, world, rank, stage1, reduce_scatter, stage2)
    lo, hi = rank * (mbyl // world), (rank + 1) * (mbyl // world)
    ok = all(got[p] == expect[p][lo:hi] for p in range(pp.n))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("l,m,rearrange", [(2, 32, 0), (2, 32, 1), (4, 64, 1)])
def test_sharded_king_world2_gloo(l, m, rearrange):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + l + m + rearrange
    procs = [ctx.Process(target=_king_worker, args=(r, 2, port, l, m, rearrange, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------------------------------------
# sharded fft1: host logic (block ownership, ONE all-to-all, output index map) with big-integer
# stand-ins for the two CUDA steps, against the literal fft1_in_place of the oracle
# ------------------------------------------------------------------------------------------------
def _fft1_worker(rank, world, port, l, m, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    import oracle_lib as ol
    from oracle_lib import pyref
    from zksaas_b200.sharding import fft1_sharded, fft1_sharded_index
    R = pyref.R_MOD
    rng = random.Random(5)                                   # same lane on every rank
    pp = pyref.PackedSharingParams(l)
    gen = pyref.Radix2Domain(m).group_gen
    mbyl = m // l
    px = [rng.randrange(R) for _ in range(mbyl)]
    expect = pyref.fft1_in_place(px, pp, gen)
    n2, cnt = mbyl // world, mbyl // world // world
    w = pow(gen, l, R)
    lg_w, lg_n2 = world.bit_length() - 1, n2.bit_length() - 1

    def brev(x, bits):
        return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0

    def local():
        blk = px[rank * n2:(rank + 1) * n2]
        i1 = brev(rank, lg_w)
        wi = pow(w, world, R)
        inner = [sum(blk[brev(i2, lg_n2)] * pow(wi, i2 * k2, R) for i2 in range(n2)) % R for k2 in range(n2)]
        return torch.from_numpy(ol.fr_np([v * pow(w, i1 * k2, R) % R for k2, v in enumerate(inner)]).view(np.int64).copy())

    def all_to_all(send):
        # gloo stand-in with all_to_all_single semantics: chunk g of the result is chunk `rank` of rank g's buffer
        bufs = [torch.zeros_like(send) for _ in range(world)]
        dist.all_gather(bufs, send)
        return torch.cat([b[rank * cnt:(rank + 1) * cnt] for b in bufs])

    def outer(recv):
        v = ol.np_fr(recv.numpy().view(np.uint64))
        wg = pow(w, n2, R)
        return [[sum(v[g * cnt + j] * pow(wg, brev(g, lg_w) * k1, R) for g in range(world)) % R for j in range(cnt)]
                for k1 in range(world)]

    got = fft1_sharded(mbyl, world, rank, local, all_to_all, outer)
    idx = fft1_sharded_index(mbyl, world, rank)
    flat = [x for row in got for x in row]
    ok = len(idx) == len(flat) and all(expect[int(k)] == x for k, x in zip(idx, flat))
    q.put((rank, ok, sorted(int(k) for k in idx)))
    dist.destroy_process_group()


@pytest.mark.parametrize("l,m", [(2, 32), (4, 64), (2, 8)])
def test_sharded_fft1_world2_gloo(l, m):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + l + m
    procs = [ctx.Process(target=_fft1_worker, args=(r, 2, port, l, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted((r, ok) for r, ok, _ in res) == [(0, True), (1, True)]
    # the two ranks' index sets partition the lane
    assert sorted(res[0][2] + res[1][2]) == list(range(m // l))
