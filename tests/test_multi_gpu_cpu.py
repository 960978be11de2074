"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: proof sharding and the split-MSM
exchange.  The device context is replaced by a stand-in that answers .msm() with the CPU oracle, so
only the plumbing (ranges, all-gather, fold) is under test here; the kernels are covered by -m gpu."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleCtx:
    """Stand-in for zkw.Context in CPU tests: same .msm contract, answered by the oracle."""

    def msm(self, scalars, bases=None, which=None):
        from oracle import cpu
        xyz = cpu.best_multiexp(np.ascontiguousarray(scalars), np.ascontiguousarray(bases), 1)
        aff = cpu.g1_to_affine(xyz)[0]
        out = np.zeros(12, dtype=np.uint64)
        if aff.any():
            out[:8] = aff
            out[8:] = cpu.fq_to_mont_one(1)
        else:
            out[4:8] = cpu.fq_to_mont_one(1)
        return out


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = importlib.import_module("webauthn-halo2_b200.multi_gpu")
    from oracle import cpu
    s = cpu.fr_random(n, 5)
    b = cpu.g1_fixed_base_mul(cpu.fr_random(n, 6), 1)
    out = mg.split_msm(OracleCtx(), s, b, rank, world, dist)
    want = cpu.g1_to_affine(cpu.best_multiexp(s, b, 1))[0]
    q.put((rank, bool(np.array_equal(out[:8], want)), mg.shard_indices(7, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [2, 101])
def test_split_msm_world2_gloo(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]


def test_shard_ranges_cover_everything():
    mg = importlib.import_module("webauthn-halo2_b200.multi_gpu")
    for n in (0, 1, 7, 8, 1 << 19):
        for world in (1, 2, 3, 8):
            ranges = [mg.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
            assert sorted(sum((mg.shard_indices(n if n < 100 else 50, r, world) for r in range(world)), [])) == list(range(n if n < 100 else 50))
    with pytest.raises(ValueError):
        mg.shard_range(10, 2, 2)
