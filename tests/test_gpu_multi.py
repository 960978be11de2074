"""Multi-GPU on real devices (skipped with fewer than 2 GPUs): the split of ONE MSM across ranks over NCCL
(per-rank partial + 96-byte all-gather) and sharded independent proofs with no collective."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    zkw = importlib.import_module("webauthn-halo2_b200")
    mg = importlib.import_module("webauthn-halo2_b200.multi_gpu")
    from oracle import cpu
    ctx = zkw.Context(rank)
    s = cpu.fr_random(n, 5)
    b = cpu.g1_fixed_base_mul(cpu.fr_random(n, 6), 2)
    out = mg.split_msm(ctx, s, b, rank, world, dist, device=torch.device("cuda", rank))
    want = cpu.g1_to_affine(cpu.best_multiexp(s, b, 2))[0]
    ok_msm = bool(np.array_equal(out[:8], want))
    # the resident form: this rank's slice of the basis (with window tables) stays in HBM, scalars are on the device
    lo, hi = mg.shard_range(n, rank, world)
    sctx = zkw.Context(rank)
    sm = mg.SplitMsm(sctx, b[lo:hi], n, rank, world, dist, device=torch.device("cuda", rank))
    s_dev = torch.from_numpy(s[lo:hi].view(np.int64)).to(torch.device("cuda", rank))
    ok_msm = ok_msm and bool(np.array_equal(sm(s_dev)[:8], want))
    sctx.close()
    # independent proofs, sharded round-robin: each rank proves its share on its own GPU
    st = zkw.ProverState(zkw.CircuitParams("Simple", 10, 2, 1, 1, 8, 88, 3), rank, synthetic=True)
    mine = mg.shard_indices(6, rank, world)
    proofs = {i: st.prove(b"assertion-%d" % i, zkw.TRANSCRIPT_EVM, seed=100 + i) for i in mine}
    q.put((rank, ok_msm, {i: p.hex() for i, p in proofs.items()}))
    dist.barrier()
    st.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_gpus() < 2, reason="needs 2 GPUs")
def test_split_msm_and_sharded_proofs_nccl():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5000, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _ in res)
    proofs = {}
    for _, _, d in res:
        proofs.update(d)
    assert sorted(proofs) == list(range(6)) and len(set(proofs.values())) == 6
    # the same assertion / seed proven on a single GPU gives the same bytes: sharding changes nothing
    zkw = importlib.import_module("webauthn-halo2_b200")
    st = zkw.ProverState(zkw.CircuitParams("Simple", 10, 2, 1, 1, 8, 88, 3), 0, synthetic=True)
    try:
        for i in (0, 3, 5):
            assert st.prove(b"assertion-%d" % i, zkw.TRANSCRIPT_EVM, seed=100 + i).hex() == proofs[i]
    finally:
        st.close()


@pytest.mark.skipif(_gpus() < 2, reason="needs 2 GPUs")
def test_one_process_batch_over_two_gpus():
    """zkw_prove_batch with provers on two different GPUs, driven by the library's threads from ONE process: every proof is
    accepted under the (device-independent) verifying key, and both GPUs produced some."""
    sys.path.insert(0, ROOT)
    zkw = importlib.import_module("webauthn-halo2_b200")
    from oracle import cpu, halo2_ref as h
    pool = zkw.ProverPool(zkw.CircuitParams.for_degree(15), device=[0, 1], workers=2)
    try:
        assert [st.ctx.device for st in pool.states] == [0, 0, 1, 1]
        vks = [st.pk.vk() for st in pool.states]
        assert all(np.array_equal(v[0], vks[0][0]) and np.array_equal(v[2], vks[0][2]) for v in vks)     # same key on both GPUs
        assertions = [zkw.synthetic_assertion(500 + i) for i in range(12)]
        before = [st.ctx.launch_count for st in pool.states]
        proofs = pool.prove_many(assertions, zkw.TRANSCRIPT_EVM)
        after = [st.ctx.launch_count for st in pool.states]
        assert after[0] + after[1] > before[0] + before[1] and after[2] + after[3] > before[2] + before[3]
        fx, pm, dg = vks[0]
        vk = h.VerifyingKey(h.Shape(15, 17, 3, 1), [cpu.g1_affine_to_ints(p) for p in fx], [cpu.g1_affine_to_ints(p) for p in pm],
                            cpu.fr_from_mont(dg.reshape(1, 4))[0])
        assert all(h.verify_proof(vk, p, "evm", tau=zkw.prover.DEV_TAU_CANONICAL) for p in proofs)
    finally:
        pool.close()
