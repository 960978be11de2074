"""Multi-GPU layout: one process per GPU, independent proofs sharded with no collective; the only
exchange step is the optional split of ONE large MSM.

torch.distributed is plumbing here (NCCL on the GPU box, gloo in the CPU tests): a barrier, a
max-reduction of elapsed times in bench.py, and a 96-byte-per-rank all-gather for the split MSM.
"""
from __future__ import annotations

import numpy as np


def shard_indices(num_items: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment of independent proofs (batch items) to ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, num_items, world))


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) of rank's slice of an n-point MSM (balanced to within 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def split_msm(ctx, scalars: np.ndarray, bases: np.ndarray, rank: int, world: int, dist=None, device=None) -> np.ndarray:
    """sum_i s_i P_i computed cooperatively: each rank reduces its index range on its own GPU to one
    point (zkw_msm_bn254_g1 over caller bases), the `world` results (96 B each) are all-gathered, and
    every rank folds them with a `world`-point MSM with unit scalars.  Returns the Jacobian (x, y, 1).

    `ctx` needs .msm(scalars, bases); `dist` is torch.distributed (initialised) or None for world == 1."""
    lo, hi = shard_range(scalars.shape[0], rank, world)
    part = ctx.msm(scalars[lo:hi], bases[lo:hi]) if hi > lo else _identity()
    if world == 1 or dist is None:
        return part
    import torch
    t = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    parts = np.stack([g.cpu().numpy().view(np.uint64) for g in gathered])  # (world, 12), each (x, y, 1) or Z = 0
    pts = parts[:, :8].copy()
    pts[~parts[:, 8:].any(axis=1)] = 0                                     # identity -> affine (0,0)
    ones = np.tile(_FR_ONE_MONT, (world, 1))
    return ctx.msm(ones, pts)


# 1 in Montgomery form (R mod r) — halo2curves Fr::one()
_FR_ONE_MONT = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=np.uint64)
_FQ_ONE_MONT = np.array([0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f], dtype=np.uint64)


def _identity() -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    out[4:8] = _FQ_ONE_MONT
    return out
