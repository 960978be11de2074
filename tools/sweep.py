"""BASELINE.json configs[4]: BN254 G1 MSM sweep 2^16..2^24 points and Fr NTT sweep 2^17..2^24 on one GPU,
timed with CUDA events on the context stream, reported against the algorithmic-byte roofline
(MSM 96*N bytes, NTT 64*N bytes, coset extension 160*n bytes; MEASURED_PEAKS.json copy bandwidth) and as
Montgomery products per second against the measured ALU ceiling (profiles/r1_modmul_peak.txt).
Writes one JSON document to stdout / --out."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
zkw = importlib.import_module("webauthn-halo2_b200")


def timed(stream, fn, reps):
    fn()
    stream.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def rand_fr(n, gen, dev):
    t = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device=dev, generator=gen)
    t[:, 3] &= (1 << 60) - 1
    return t


def sweep_rows(zkw_mod, min_k, max_k, peak, device=0, log=None):
    """One row per size: MSM over the resident window tables, inverse / forward transform, coset extension (k >= 17)."""
    dev = torch.device("cuda", device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    tau = np.array([0x1234567890ABCDEF, 0x0FEDCBA987654321, 0x1111111111111111, 0x0222222222222222], dtype=np.uint64)
    rows = []
    for k in range(min_k, max_k + 1):
        n = 1 << k
        ctx = zkw_mod.Context(device)
        stream = torch.cuda.ExternalStream(ctx.stream, device=device)
        torch.cuda.set_stream(stream)
        row = {"k": k, "n": n}
        reps = 5 if k <= 21 else 3
        ctx.srs_setup(k, tau)
        s = rand_fr(n, gen, dev)
        torch.cuda.synchronize()
        ms = timed(stream, lambda: ctx.msm_dev(s, n, zkw_mod.BASES_G), reps)
        row["msm_ms"] = ms
        row["msm_alg_gbs"] = 96 * n / ms / 1e6
        row["msm_frac_of_hbm"] = row["msm_alg_gbs"] / peak
        row["msm_points_per_s"] = n / ms * 1e3
        row["msm_windows"] = -(-255 // ctx.msm_window_bits(n))     # mixed additions per point: the plan the library picks for n
        row["msm_modmul_per_s"] = row["msm_windows"] * n * 10 / ms * 1e3    # windows x (8M + 2S) per mixed addition, whole MSM time
        if k >= 17:
            a = rand_fr(n, gen, dev)
            ms = timed(stream, lambda: ctx.lagrange_to_coeff_dev(a, k), reps)
            row["intt_ms"] = ms
            row["intt_alg_gbs"] = 64 * n / ms / 1e6
            row["intt_frac_of_hbm"] = row["intt_alg_gbs"] / peak
            row["intt_modmul_per_s"] = (n // 2) * k / ms * 1e3
            ms = timed(stream, lambda: ctx.coeff_to_lagrange_dev(a, k), reps)
            row["ntt_ms"] = ms
            row["ntt_alg_gbs"] = 64 * n / ms / 1e6
            if k + 2 <= 26:
                e = torch.empty((4 * n, 4), dtype=torch.int64, device=dev)
                ms = timed(stream, lambda: ctx.coeff_to_extended_dev(a, k, k + 2, e), reps)
                row["coset_ext_ms"] = ms
                row["coset_ext_alg_gbs"] = 160 * n / ms / 1e6
                del e
            del a
        rows.append(row)
        if log:
            print(json.dumps(row), file=log, flush=True)
        ctx.close()
        del s
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-k", type=int, default=16)
    ap.add_argument("--max-k", type=int, default=24)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = sweep_rows(zkw, args.min_k, args.max_k, peak, 0, sys.stderr)
    # measured 254-bit Montgomery products per second (tools/modmul_bench.cu), from the newest committed record
    import glob
    import re
    modmul_peak, modmul_src = None, None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_modmul_peak.txt")), reverse=True):
        vals = [float(x) for x in re.findall(r"([0-9.]+)\s*G\s*(?:products|modmul)", open(path).read())]
        if vals:
            modmul_peak, modmul_src = max(vals) * 1e9, os.path.relpath(path, ROOT)
            break
    doc = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "modmul_peak_per_s": modmul_peak, "modmul_peak_source": modmul_src, "rows": rows}
    text = json.dumps(doc, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
