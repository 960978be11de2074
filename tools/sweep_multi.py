"""BASELINE.json configs[4], multi-GPU leg: ONE BN254 G1 MSM of 2^k points split across the ranks of a
torchrun job: every rank keeps the window tables of its contiguous slice of the SRS
resident, reduces its slice to one point, the 96-byte partial results are all-gathered over NCCL and folded on the host
(multi_gpu._gather_and_fold, the code path of multi_gpu.SplitMsm and of bench.py's split_msm leg).
Timed end to end (device MSM + collective + fold), max over ranks; the split result is checked against a
single-GPU MSM over the whole SRS.  NTTs are not sharded (every size of the sweep fits one GPU): N GPUs run N
independent transforms, so their aggregate rate is N x the single-GPU figure of tools/sweep.py.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi.py --ks 20 22 24 --out gpurun_out/x.json
"""
import argparse, importlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
zkw = importlib.import_module("webauthn-halo2_b200")
from importlib import import_module
mg = import_module("webauthn-halo2_b200.multi_gpu")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", type=int, nargs="+", default=[20, 22, 24])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    tau = np.array([0x1234567890ABCDEF, 0x0FEDCBA987654321, 0x1111111111111111, 0x0222222222222222], dtype=np.uint64)
    rows = []
    for k in args.ks:
        n = 1 << k
        lo, hi = mg.shard_range(n, rank, world)
        full = zkw.Context(local)
        full.msm_config(0, False)            # whole SRS without window tables: only used to cut slices and to check
        full.srs_setup(k, tau)
        g = full.srs_get(zkw.BASES_G, n)
        ctx = zkw.Context(local)
        ctx.srs_load(g[lo:hi])               # this rank's slice, with window tables
        gen = torch.Generator(device=dev); gen.manual_seed(1000 + k)   # same scalars on every rank
        s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device=dev, generator=gen); s[:, 3] &= (1 << 60) - 1
        mine = s[lo:hi].contiguous()

        def split_once():
            # the product path: per-rank MSM over resident window tables, NCCL all-gather of world x 96 bytes, host fold
            part = ctx.msm_dev(mine, hi - lo, zkw.BASES_G)
            return mg._gather_and_fold(part, world, dist if world > 1 else None, dev)

        res = split_once()
        ts = []
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = split_once()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            ts.append(float(dt.item()) * 1e3)
        ts.sort()
        ms = ts[len(ts) // 2]
        ok = None
        if k <= 22:
            want = full.msm_dev(s, n, zkw.BASES_G)
            ok = bool(np.array_equal(want, res))
        row = {"k": k, "n": n, "gpus": world, "split_msm_ms": ms, "points_per_s": n / ms * 1e3, "alg_gbs": 96 * n / ms / 1e6,
               "matches_single_gpu_msm": ok}
        if rank == 0:
            print(json.dumps(row), file=sys.stderr, flush=True)
        rows.append(row)
        ctx.close(); full.close()
        del s, mine, g
        torch.cuda.empty_cache()
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        doc = {"gpu": torch.cuda.get_device_name(local), "gpus": world, "hbm_peak_gbs_per_gpu": peak, "rows": rows,
               "note": "wall clock around device MSM + NCCL all-gather of 96-byte partials + fold, max over ranks"}
        text = json.dumps(doc, indent=1)
        if args.out:
            open(args.out, "w").write(text)
        print(text)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
