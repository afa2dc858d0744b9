"""GPU, >= 2 devices: the NCCL paths (one process per GPU) against the single-GPU results.
  - one king pipeline sharded by share columns: stage 1 -> ONE sum reduce-scatter -> stage 2
  - one large MSM sharded by point range: partial sums -> all-gather -> combine
  - one fft1 lane sharded by contiguous blocks: inner transforms -> ONE all-to-all -> outer DFT
Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import zksaas_b200 as z
    from zksaas_b200 import capi, sharding
    from zksaas_b200.api import fr_image
    lib = z.lib()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = capi.ctx_p()
    capi.check(lib.zkg_ctx_create(rank, C.c_void_p(stream.cuda_stream), C.byref(ctx)))
    ok = {}

    def rand_fr(k, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device=dev, generator=g)
        t[:, 3] &= (1 << 61) - 1
        return t

    # ---- sharded king vs the single-GPU king (same seeds on every rank) ----
    for l, mbyl, rearr in ((2, 1 << 12, 1), (2, 1 << 12, 0), (4, 1 << 10, 1)):
        n, t = 4 * l, l
        m = mbyl * l
        dom = z.Radix2EvaluationDomain.new(m)
        gen, g = dom.group_gen(), z.Radix2EvaluationDomain.new(2 * m).element(1)
        shares = rand_fr(n * mbyl, 11).reshape(n, mbyl, 4)
        rnd = rand_fr(mbyl * t, 12)
        full = torch.empty((n, mbyl, 4), dtype=torch.int64, device=dev)
        capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, n, mbyl, l, gen.ctypes.data,
                                               g.ctypes.data, rearr, C.c_void_p(rnd.data_ptr()), C.c_void_p(full.data_ptr())))
        lo, hi = sharding.shard_range(mbyl, world, rank)
        loc = shares[:, lo:hi, :].contiguous()
        rloc = rnd[lo * t:hi * t].contiguous()
        got = sharding.king_fft2_sharded_cuda(ctx, lib, torch, dist, loc, mbyl, l, gen, g, rearr, rloc, rank, world)
        torch.cuda.synchronize()
        ok[f"king_l{l}_r{rearr}"] = bool((got == full[:, lo:hi, :]).all())

    # ---- sharded fft1 (four-step, ONE all-to-all) vs the single-GPU fft1 ----
    for l, mbyl in ((2, 1 << 14), (2, 1 << 21), (8, 1 << 6)):
        gen = z.Radix2EvaluationDomain.new(mbyl * l).group_gen()
        px = rand_fr(mbyl, 31 + l)
        full = px.clone()
        capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(full.data_ptr()), mbyl, l, gen.ctypes.data, None, None))
        n2 = mbyl // world
        blk = px[rank * n2:(rank + 1) * n2].clone()
        got = sharding.fft1_sharded_cuda(ctx, lib, torch, dist, blk, mbyl, l, gen, rank, world)
        idx = torch.from_numpy(sharding.fft1_sharded_index(mbyl, world, rank)).to(dev)
        torch.cuda.synchronize()
        ok[f"fft1_l{l}_n{mbyl}"] = bool((got.reshape(-1, 4) == full[idx]).all())

    # ---- sharded MSM vs the single-GPU MSM ----
    npts = 1 << 14
    a, s = rand_fr(npts, 21), rand_fr(npts, 22)
    bases = torch.empty((npts, 64), dtype=torch.uint8, device=dev)
    capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s.data_ptr()), npts, C.c_void_p(bases.data_ptr())))
    ref = torch.zeros(12, dtype=torch.int64, device=dev)
    capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a.data_ptr()), npts, C.c_void_p(ref.data_ptr())))
    out = torch.zeros(12, dtype=torch.int64, device=dev)

    def partial(lo, hi):
        p = torch.zeros(16, dtype=torch.int64, device=dev)
        capi.check(lib.zkg_msm_bn254_partial_dev(ctx, 1, C.c_void_p(bases[lo:].data_ptr()), C.c_void_p(a[lo:].data_ptr()),
                                                 hi - lo, C.c_void_p(p.data_ptr())))
        return p

    def all_gather(mine):
        g = torch.zeros(16 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(g, mine)
        return g

    def combine(g):
        capi.check(lib.zkg_msm_combine_dev(ctx, 1, C.c_void_p(g.data_ptr()), world, C.c_void_p(out.data_ptr())))
        return out

    got = sharding.sharded_msm(npts, world, rank, partial, lambda mine: [all_gather(mine)] * world, lambda parts: combine(parts[0]))
    torch.cuda.synchronize()
    ok["msm"] = bool((got == ref).all())
    q.put((rank, ok))
    dist.barrier()
    lib.zkg_ctx_destroy(ctx)
    dist.destroy_process_group()


def test_sharded_king_and_msm_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok in res:
        assert all(ok.values()), (rank, ok)
