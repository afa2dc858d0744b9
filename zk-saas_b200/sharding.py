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
