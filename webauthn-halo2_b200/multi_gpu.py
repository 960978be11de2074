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
    """sum_i s_i P_i computed cooperatively from HOST arrays: each rank reduces its index range on its own GPU to one
    point (zkw_msm_bn254_g1 over caller bases), the `world` results (96 B each) are all-gathered and folded on the host
    with world - 1 point additions (zkw_g1_sum).  Returns the Jacobian (x, y, 1).  For repeated MSMs over a fixed basis
    use SplitMsm, which keeps each rank's slice (with its window tables) resident.

    `ctx` needs .msm(scalars, bases); `dist` is torch.distributed (initialised) or None for world == 1."""
    lo, hi = shard_range(scalars.shape[0], rank, world)
    part = ctx.msm(scalars[lo:hi], bases[lo:hi]) if hi > lo else _identity()
    return _gather_and_fold(part, world, dist, device)


def _gather_and_fold(part: np.ndarray, world: int, dist, device) -> np.ndarray:
    if world == 1 or dist is None:
        return part
    import torch
    from . import native
    t = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        t = t.to(device, non_blocking=True)
    gathered = torch.empty((world, 12), dtype=torch.int64, device=t.device)
    dist.all_gather([gathered[i] for i in range(world)], t)             # world x 96 bytes
    return native.g1_sum(gathered.cpu().numpy().view(np.uint64))       # world - 1 point additions on the host


class SplitMsm:
    """ONE large MSM split over the ranks of a job (BASELINE configs[4]): rank r keeps the window tables of its
    contiguous slice [lo, hi) of the basis resident in HBM; a call reduces the rank's slice of the scalars (already on the
    device) to one point, all-gathers the 96-byte partial results over NCCL and folds them on the host (world - 1 point
    additions).  Latency bound: the collective moves world x 96 bytes."""

    def __init__(self, ctx, bases_slice, n_total: int, rank: int, world: int, dist=None, device=None):
        self.ctx, self.rank, self.world, self.dist, self.device = ctx, rank, world, dist, device
        self.lo, self.hi = shard_range(n_total, rank, world)
        if hasattr(bases_slice, "data_ptr"):
            ctx.srs_load_dev(bases_slice, None, self.hi - self.lo)
        else:
            ctx.srs_load(bases_slice)
        from . import native
        self._which = native.BASES_G

    def __call__(self, scalars_slice_dev) -> np.ndarray:
        """scalars_slice_dev: this rank's (hi - lo, 4) device tensor / address.  Returns the full sum (x, y, 1) on every rank."""
        m = self.hi - self.lo
        part = self.ctx.msm_dev(scalars_slice_dev, m, self._which) if m else _identity()
        return _gather_and_fold(part, self.world, self.dist, self.device)


# 1 in Montgomery form (R mod r) — halo2curves Fr::one()
_FR_ONE_MONT = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=np.uint64)
_FQ_ONE_MONT = np.array([0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f], dtype=np.uint64)


def _identity() -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    out[4:8] = _FQ_ONE_MONT
    return out
