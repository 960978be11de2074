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
    """Stand-in for zkw.Context in CPU tests: same .msm / .srs_load / .msm_dev contract, answered by the oracle."""

    def srs_load(self, g, g_lagrange=None):
        self.g = np.ascontiguousarray(g)

    def msm_dev(self, scalars, n, which=0, bases_dev=None):
        return self.msm(np.ascontiguousarray(scalars)[:n], self.g[:n])

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
    ok = bool(np.array_equal(out[:8], want))
    # the resident form: each rank holds its slice of the basis, scalars arrive slice by slice
    lo, hi = mg.shard_range(n, rank, world)
    sm = mg.SplitMsm(OracleCtx(), b[lo:hi], n, rank, world, dist)
    ok = ok and bool(np.array_equal(sm(s[lo:hi])[:8], want))
    q.put((rank, ok, mg.shard_indices(7, rank, world)))
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


def test_g1_sum_host_fold_matches_oracle(zkw, oracle):
    """zkw_g1_sum (host code of the library): the fold of per-rank partial results, against the oracle's point addition."""
    import ctypes as C
    native = importlib.import_module("webauthn-halo2_b200.native")
    pts = oracle.g1_fixed_base_mul(oracle.fr_random(5, 77))
    xyz = np.zeros((7, 12), dtype=np.uint64)
    one = oracle.fq_to_mont_one(1)
    for i in range(5):
        xyz[i, :8] = pts[i]
        xyz[i, 8:] = one
    xyz[5, 4:8] = one                                     # identity (z = 0)
    xyz[6] = xyz[2]                                       # a repeated point: the doubling branch
    u64p = C.POINTER(C.c_uint64)
    acc = xyz[0].copy()
    for i in (1, 2, 3, 4, 6):
        nxt = np.empty(12, dtype=np.uint64)
        oracle.lib().zko_g1_add(nxt.ctypes.data_as(u64p), acc.ctypes.data_as(u64p), xyz[i].ctypes.data_as(u64p))
        acc = nxt
    want = oracle.g1_to_affine(acc.reshape(1, 12))[0]
    got = native.g1_sum(xyz)
    assert np.array_equal(got[:8], want) and np.array_equal(got[8:], one)
    # P + (-P) = identity
    neg = xyz[:2].copy()
    neg[1] = xyz[0]
    from oracle import pyref as pr
    x, y = oracle.g1_affine_to_ints(pts[0])
    neg[1, :8] = oracle.g1_ints_to_affine((x, (-y) % pr.P))
    assert not native.g1_sum(neg)[8:].any()
    assert not native.g1_sum(np.zeros((0, 12), dtype=np.uint64))[8:].any()
