"""Point-range sharding of one large MSM across the GPUs of a box.

MSM is linear, so rank g computes the partial sum over points [lo_g, hi_g) with the full
single-GPU pipeline, the 128-byte XYZZ partials are exchanged with ONE all-gather (NCCL over
NVLink on the GPU box, gloo in the CPU tests) and every rank adds them.  The per-rank compute and
the final add are injected so the same host logic runs against the CUDA library (bench.py, GPU
tests) and against the CPU oracle (world_size-2 gloo test)."""
from __future__ import annotations

import numpy as np

Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_ONE_FQ = (1 << 256) % Q_MOD
ONE_FQ_LIMBS = np.array([(_ONE_FQ >> (64 * i)) & ((1 << 64) - 1) for i in range(4)], dtype=np.uint64)


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of n points: the first n % world ranks get one extra point."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def xyz_to_xyzz_g1(xyz) -> np.ndarray:
    """normalised Jacobian image (x, y, 1) / (1, 1, 0) -> XYZZ image (x, y, 1, 1) / zeros."""
    xyz = np.asarray(xyz, dtype=np.uint64).reshape(3, 4)
    out = np.zeros((4, 4), dtype=np.uint64)
    if xyz[2].any():
        out[0], out[1], out[2], out[3] = xyz[0], xyz[1], ONE_FQ_LIMBS, ONE_FQ_LIMBS
    return out.reshape(16)


def sharded_msm(n_total: int, world: int, rank: int, partial_fn, all_gather_fn, combine_fn):
    """partial_fn(lo, hi) -> this rank's partial (any fixed-size array);
    all_gather_fn(partial) -> list of `world` partials in rank order; combine_fn(list) -> result."""
    lo, hi = shard_range(n_total, world, rank)
    mine = partial_fn(lo, hi)
    parts = all_gather_fn(mine)
    if len(parts) != world:
        raise RuntimeError("all_gather returned %d partials for world size %d" % (len(parts), world))
    return combine_fn(parts)


def king_sharded(mbyl: int, world: int, rank: int, stage1_fn, reduce_scatter_fn, stage2_fn):
    """One king pipeline (dist-primitives/src/dfft/mod.rs:264-304) sharded by share columns.

    stage1_fn(lo, hi)      -> FULL-size S (m x 4 int64 words, pack order), zero except the slots this
                              rank's columns [lo, hi) produce (unpack -> fft2 -> g^i);
    reduce_scatter_fn(S)   -> this rank's contiguous slice of the element-wise SUM over ranks.  Every slot
                              is written by exactly one rank, so the 64-bit lane sums never carry: the one
                              collective of the pipeline *is* the rotate / bit-reverse / stride permutation;
    stage2_fn(lo, hi, S_r) -> the rank's output columns [lo, hi) (pack), party-major.
    """
    if mbyl % world:
        raise ValueError("king_sharded: m/l must be divisible by the number of ranks")
    lo, hi = shard_range(mbyl, world, rank)
    return stage2_fn(lo, hi, reduce_scatter_fn(stage1_fn(lo, hi)))


class PeerBuffers:
    """`copies` peer-visible device buffers of `nbytes` per rank (one process per GPU).  Allocated by the library
    (zkg_shared_alloc), CUDA IPC handles exchanged through the process group, peers mapped with zkg_shared_open.
    ptrs(i) is the ctypes array of buffer i as seen from this rank, indexed by rank (own entry = local allocation).
    Successive exchanges alternate between the copies, so ONE barrier per exchange suffices: a rank can only start
    writing copy i again after every peer has passed the barrier of the exchange in between, i.e. finished reading it."""

    def __init__(self, ctx, lib, dist, nbytes, rank, world, copies=2):
        import ctypes as C
        from .capi import check
        self.ctx, self.lib, self.rank, self.world, self.copies = ctx, lib, rank, world, copies
        self.local, self.peer, self._arrays, self.turn = [], [], [], 0
        for _ in range(copies):
            p = C.c_void_p()
            h = (C.c_uint8 * 64)()
            check(lib.zkg_shared_alloc(ctx, nbytes, C.byref(p), h))
            handles = [None] * world
            if world > 1:
                dist.all_gather_object(handles, bytes(h))
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(p.value)
                else:
                    q = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                    check(lib.zkg_shared_open(ctx, hb, C.byref(q)))
                    ptrs.append(q.value)
            self.local.append(p.value)
            self.peer.append(ptrs)
            self._arrays.append((C.c_void_p * world)(*ptrs))

    def next(self):
        i = self.turn
        self.turn = (self.turn + 1) % self.copies
        return i

    def ptrs(self, i):
        return self._arrays[i]

    def close(self):
        from .capi import check
        for i in range(self.copies):
            for r, q in enumerate(self.peer[i]):
                if r != self.rank:
                    check(self.lib.zkg_shared_close(self.ctx, q))
        for p in self.local:
            check(self.lib.zkg_shared_free(self.ctx, p))
        self.local, self.peer, self._arrays = [], [], []


def stream_barrier(torch, dist, token):
    """Orders `every rank has finished what it enqueued so far` before what this rank enqueues next, without stalling the
    host: a 4-byte all-reduce on the current stream (NCCL over NVLink; `token` is a 1-element device tensor)."""
    dist.all_reduce(token)


def king_fft2_sharded_cuda(ctx, lib, torch, dist, shares_local, mbyl, l, gen, g, rearrange, rand_local, rank, world,
                           peers=None, token=None):
    """CUDA instantiation of king_sharded.  shares_local: (n, cols, 4) int64 device tensor holding this rank's columns
    of every party's vector; rand_local: (cols*t, 4).  Returns (n, cols, 4).

    With `peers` (a PeerBuffers of m/world*32 bytes per copy) stage 1 stores every value straight into the memory of
    the rank that owns its output column (NVLink peer stores: zkg_king_stage1_scatter_bn254_dev) and the only
    collective left is a 4-byte barrier.  Without it (round-1 path, kept for comparison): stage 1 scatters into a
    zero-filled FULL-size buffer and one NCCL sum reduce-scatter moves world-times the bytes."""
    import ctypes as C
    from .capi import check
    n = shares_local.shape[0]
    cols = shares_local.shape[1]
    m = mbyl * l
    dev = shares_local.device
    if peers is not None:
        if mbyl % world:
            raise ValueError("king_sharded: m/l must be divisible by the number of ranks")
        lo, hi = shard_range(mbyl, world, rank)
        i = peers.next()
        check(lib.zkg_king_stage1_scatter_bn254_dev(ctx, C.c_void_p(shares_local.data_ptr()), None, n, lo, hi - lo, mbyl, l,
                                                    gen.ctypes.data, g.ctypes.data, 1 if rearrange else 0, peers.ptrs(i), world))
        if world > 1:
            stream_barrier(torch, dist, token)
        out = torch.empty((n, cols, 4), dtype=torch.int64, device=dev)
        check(lib.zkg_king_stage2_bn254_dev(ctx, C.c_void_p(peers.local[i]), C.c_void_p(rand_local.data_ptr()), hi - lo, l,
                                            C.c_void_p(out.data_ptr())))
        return out

    def stage1(lo, hi):
        S = torch.zeros((m, 4), dtype=torch.int64, device=dev)
        check(lib.zkg_king_stage1_bn254_dev(ctx, C.c_void_p(shares_local.data_ptr()), None, n, lo, hi - lo, mbyl, l,
                                            gen.ctypes.data, g.ctypes.data, 1 if rearrange else 0,
                                            C.c_void_p(S.data_ptr())))
        return S

    def reduce_scatter(S):
        out = torch.empty((m // world, 4), dtype=torch.int64, device=dev)
        if world == 1:
            out.copy_(S)
        else:
            dist.reduce_scatter_tensor(out, S, op=dist.ReduceOp.SUM)
        return out

    def stage2(lo, hi, S_r):
        out = torch.empty((n, cols, 4), dtype=torch.int64, device=dev)
        check(lib.zkg_king_stage2_bn254_dev(ctx, C.c_void_p(S_r.data_ptr()), C.c_void_p(rand_local.data_ptr()), hi - lo, l,
                                            C.c_void_p(out.data_ptr())))
        return out

    return king_sharded(mbyl, world, rank, stage1, reduce_scatter, stage2)


# ------------------------------------------------------------------------------------------------
# fft1 of ONE lane sharded over the ranks: four-step with a single all-to-all
# ------------------------------------------------------------------------------------------------
def fft1_sharded(mbyl: int, world: int, rank: int, local_fn, all_to_all_fn, outer_fn):
    """fft1_in_place (dist-primitives/src/dfft/mod.rs:178-208) of a lane of mbyl = m/l elements whose
    contiguous block [rank*N2, (rank+1)*N2), N2 = mbyl/world, lives on this rank.

    local_fn()            -> send buffer of N2 elements (inner transform + twiddles), ordered by column:
                             chunk d (N2/world elements) is what rank d needs;
    all_to_all_fn(send)   -> receive buffer, chunk g from rank g (world x N2/world elements);
    outer_fn(recv)        -> (world x N2/world) array: row k1, column j  =  X[(rank*N2/world + j) + N2*k1],
                             with fft1(px)[k] = X[(k+1) mod mbyl]  (see fft1_sharded_index).
    """
    if world & (world - 1) or mbyl % (world * world):
        raise ValueError("fft1_sharded: ranks must be a power of two and m/l divisible by ranks^2")
    return outer_fn(all_to_all_fn(local_fn()))


def fft1_sharded_index(mbyl: int, world: int, rank: int) -> np.ndarray:
    """fft1 output positions held by `rank` after fft1_sharded, in the (k1, j) row-major order of its result."""
    n2 = mbyl // world
    cnt = n2 // world
    k1 = np.arange(world, dtype=np.int64)[:, None]
    j = np.arange(cnt, dtype=np.int64)[None, :]
    return ((rank * cnt + j + n2 * k1 - 1) % mbyl).reshape(-1)


def fft1_sharded_cuda(ctx, lib, torch, dist, block, mbyl, l, gen, rank, world, pre_scale=None, peers=None, token=None):
    """CUDA instantiation.  block: (N2, 4) int64 device tensor (this rank's slice of the lane; overwritten).  Returns a
    (world, N2/world, 4) device tensor laid out as fft1_sharded describes.

    With `peers` (PeerBuffers of N2*32 bytes per copy) the twiddle pass of the local step stores each chunk straight
    into the destination rank's receive buffer (zkg_fft1_shard_local_scatter_bn254_dev): the all-to-all is fused into
    the kernel and only a 4-byte barrier remains.  Without it: one NCCL all_to_all_single."""
    import ctypes as C
    from .capi import check
    n2 = mbyl // world
    cnt = n2 // world
    if peers is not None:
        if world & (world - 1) or mbyl % (world * world):
            raise ValueError("fft1_sharded: ranks must be a power of two and m/l divisible by ranks^2")
        i = peers.next()
        check(lib.zkg_fft1_shard_local_scatter_bn254_dev(ctx, C.c_void_p(block.data_ptr()), n2, l, world, rank, gen.ctypes.data,
                                                         pre_scale.ctypes.data if pre_scale is not None else None, peers.ptrs(i)))
        if world > 1:
            stream_barrier(torch, dist, token)
        out = torch.empty((world, cnt, 4), dtype=torch.int64, device=block.device)
        check(lib.zkg_fft1_shard_outer_bn254_dev(ctx, C.c_void_p(peers.local[i]), cnt, n2, l, world, gen.ctypes.data,
                                                 C.c_void_p(out.data_ptr())))
        return out

    def local():
        check(lib.zkg_fft1_shard_local_bn254_dev(ctx, C.c_void_p(block.data_ptr()), n2, l, world, rank, gen.ctypes.data,
                                                 pre_scale.ctypes.data if pre_scale is not None else None))
        return block

    def all_to_all(send):
        if world == 1:
            return send
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        return recv

    def outer(recv):
        out = torch.empty((world, cnt, 4), dtype=torch.int64, device=block.device)
        check(lib.zkg_fft1_shard_outer_bn254_dev(ctx, C.c_void_p(recv.data_ptr()), cnt, n2, l, world, gen.ctypes.data,
                                                 C.c_void_p(out.data_ptr())))
        return out

    return fft1_sharded(mbyl, world, rank, local, all_to_all, outer)
