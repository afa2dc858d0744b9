"""GPU, >= 2 devices: the multi-GPU paths against the single-GPU results.
  one process per GPU (torchrun-style):
  - one king pipeline sharded by share columns: stage 1 storing straight into the owners' memory over NVLink (CUDA IPC peer
    buffers) + a 4-byte barrier + stage 2; and the round-1 path (zero-filled buffer + ONE sum reduce-scatter)
  - one fft1 lane sharded by contiguous blocks: inner transforms, twiddles fused with the peer-store all-to-all, outer DFT;
    and the round-1 path (ONE NCCL all-to-all)
  - one large MSM sharded by point range: partial sums -> all-gather -> combine
  one process, a device list (the C-ABI entry points a Rust host calls): zkg_*_sharded
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
        peers = sharding.PeerBuffers(ctx, lib, dist, m // world * 32, rank, world)
        token = torch.zeros(1, dtype=torch.int32, device=dev)
        for rep in range(3):                                       # repeated: the two buffer copies alternate
            got = sharding.king_fft2_sharded_cuda(ctx, lib, torch, dist, loc, mbyl, l, gen, g, rearr, rloc, rank, world,
                                                  peers=peers, token=token)
            torch.cuda.synchronize()
            ok[f"king_peer_l{l}_r{rearr}_{rep}"] = bool((got == full[:, lo:hi, :]).all())
        dist.barrier()
        peers.close()

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
        peers = sharding.PeerBuffers(ctx, lib, dist, n2 * 32, rank, world)
        token = torch.zeros(1, dtype=torch.int32, device=dev)
        for rep in range(3):
            blk = px[rank * n2:(rank + 1) * n2].clone()
            got = sharding.fft1_sharded_cuda(ctx, lib, torch, dist, blk, mbyl, l, gen, rank, world, peers=peers, token=token)
            torch.cuda.synchronize()
            ok[f"fft1_peer_l{l}_n{mbyl}_{rep}"] = bool((got.reshape(-1, 4) == full[idx]).all())
        dist.barrier()
        peers.close()

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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_king_and_msm_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
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


@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_single_process_device_list_entry_points(n_dev):
    """zkg_*_sharded (one process, a device list: what the unchanged Rust caller of INTEGRATION.md uses) against the
    single-GPU entry points, bit for bit: MSM G1 / G2, registered bases, king closure (both packings, with a dropout),
    deg_red king, fft1 (with pre-scale and in-mask)."""
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs >= {n_dev} GPUs")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from oracle_lib import _p
    import zksaas_b200 as z
    from zksaas_b200 import capi
    lib = z.lib()
    o = ol.oracle()
    rng = np.random.default_rng(40 + n_dev)
    devs = (C.c_int32 * n_dev)(*range(n_dev))
    u64p = C.POINTER(C.c_uint64)
    # ---- MSM ----
    n = (1 << 16) + 123                                              # not divisible by the device count
    bases = np.zeros((n, 72), dtype=np.uint8)
    o.zko_g1_fixed_base(_p(ol.rand_fr(rng, n)), n, bases.ctypes.data, 72)
    sc = ol.rand_fr(rng, n)
    ref = z.msm_g1(bases, sc)
    out = np.zeros(12, dtype=np.uint64)
    capi.check(lib.zkg_msm_bn254_g1_sharded(devs, n_dev, bases.ctypes.data, 72, n, _p(sc), n, _p(out)))
    assert (out == ref).all()
    h = C.c_uint64(0)
    capi.check(lib.zkg_bases_register_sharded(devs, n_dev, 1, bases.ctypes.data, 72, n, C.byref(h)))
    for _ in range(2):
        out[:] = 0
        capi.check(lib.zkg_msm_bn254_registered(h.value, _p(sc), n, _p(out)))
        assert (out == ref).all()
    capi.check(lib.zkg_bases_release(h.value))
    n2 = 1 << 13
    b2 = np.zeros((n2, 136), dtype=np.uint8)
    o.zko_g2_fixed_base(_p(ol.rand_fr(rng, n2)), n2, b2.ctypes.data, 136)
    s2 = ol.rand_fr(rng, n2)
    out2 = np.zeros(24, dtype=np.uint64)
    capi.check(lib.zkg_msm_bn254_g2_sharded(devs, n_dev, b2.ctypes.data, 136, n2, _p(s2), n2, _p(out2)))
    assert (out2 == z.msm_g2(b2, s2)).all()
    # ---- king closure / deg_red ----
    l, mbyl = 2, 1 << 14
    pp = z.PackedSharingParams.new(l)
    dom = z.Radix2EvaluationDomain.new(mbyl * l)
    gen, g = dom.group_gen(), z.Radix2EvaluationDomain.new(2 * mbyl * l).element(1)
    shares = [ol.rand_fr(rng, mbyl) for _ in range(pp.n)]
    rnd = ol.rand_fr(rng, mbyl * pp.t)
    for parties, rearr in ((list(range(8)), 1), (list(range(8)), 0), ([0, 1, 2, 4, 5, 6, 7], 1)):
        sh = [shares[p] for p in parties]
        ref_out = z.king_fft2(sh, parties, pp, gen, g, bool(rearr), rnd)
        outs = [np.zeros((mbyl, 4), dtype=np.uint64) for _ in range(pp.n)]
        in_arr = (u64p * len(sh))(*[x.ctypes.data_as(u64p) for x in sh])
        out_arr = (u64p * pp.n)(*[x.ctypes.data_as(u64p) for x in outs])
        par = (C.c_uint32 * len(parties))(*parties)
        capi.check(lib.zkg_king_fft2_bn254_sharded(devs, n_dev, in_arr, par, len(parties), mbyl, l, _p(gen), _p(g), rearr, _p(rnd), out_arr))
        assert all((a == b).all() for a, b in zip(outs, ref_out)), (parties, rearr)
    ref_out = z.deg_red_king(shares, list(range(8)), pp, rnd)
    outs = [np.zeros((mbyl, 4), dtype=np.uint64) for _ in range(pp.n)]
    in_arr = (u64p * 8)(*[x.ctypes.data_as(u64p) for x in shares])
    out_arr = (u64p * 8)(*[x.ctypes.data_as(u64p) for x in outs])
    capi.check(lib.zkg_deg_red_king_bn254_sharded(devs, n_dev, in_arr, None, 8, mbyl, l, _p(rnd), out_arr))
    assert all((a == b).all() for a, b in zip(outs, ref_out))
    # ---- fft1 ----
    for lg in (16, 20):
        mb = 1 << lg
        d2 = z.Radix2EvaluationDomain.new(mb * l)
        px, mask = ol.rand_fr(rng, 64)[np.arange(mb) % 64].copy(), ol.rand_fr(rng, 64)[(np.arange(mb) * 7) % 64].copy()
        px[:, 0] += np.arange(mb, dtype=np.uint64)                   # distinct values (low limb cannot overflow past 2^253)
        e = z.fft1_in_place(px.copy(), pp, d2.group_gen_inv(), pre_scale=d2.size_inv(), in_mask=mask)
        got = px.copy()
        capi.check(lib.zkg_fft1_bn254_sharded(devs, n_dev, _p(got), mb, l, _p(d2.group_gen_inv()), _p(d2.size_inv()), _p(mask)))
        assert (got == e).all(), lg
