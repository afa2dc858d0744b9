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
