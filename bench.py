#!/usr/bin/env python
"""bench.py — proofs/s for the P-256 ECDSA circuit's prover at k = 19 on N B200s.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line from rank 0.

A "step" is ONE complete proof (`--workload proof`, the default): zkw_create_proof on the device for the k = 19 shape of
halo2-circuits/src/configs/bench_ecdsa.config:1 (1 advice / 1 lookup / 1 fixed column; degree-5 constraint system, extended
domain 2^21), EVM transcript + GWC (BASELINE.json configs[1]), over the REAL ECDSA verification circuit's witness for a
synthetic signed WebAuthn assertion: 15 MSMs of 2^19 points, 5 iNTT(2^19), 5 coset NTTs 2^19 -> 2^21, the quotient over 2^21
rows and 14 cosets, 1 iNTT(2^21), plus lookup permutation, grand products, evaluations, openings and the transcript.

    value       proofs/s with the advice column already resident in HBM (device-timed, CUDA events, max over ranks)
    e2e         the same through the public API (ProverState.prove -> zkw_prover_prove): assertion bytes in host memory ->
                witness synthesis on the host -> H2D of the assigned cells -> proof bytes back; `e2e.three_seam` is the
                hot-path call sequence through the three host-pointer seams a [patch]ed halo2_proofs would bind
    batch       BASELINE configs[2]/[3]: 64 proofs per GPU through zkw_prove_batch with three provers in flight
    split_msm   (N > 1) BASELINE configs[4]'s multi-GPU leg: one 2^20 / 2^22-point MSM split over the ranks
    sweep       (N = 1) BASELINE configs[4]'s single-GPU leg: MSM 2^16 .. 2^24 points and iNTT / coset extension 2^17 .. 2^24,
                milliseconds and fraction of the HBM roofline by algorithmic bytes (tools/sweep.py's code; --no-sweep skips it)
    roofline    the dominant kernel (msm_accumulate_kernel): algorithmic bytes / measured launch time against the measured
                HBM copy bandwidth; DRAM traffic and the ALU ceiling are read from the committed ncu export / profiles
    cpu_baseline / --impl reference
                the CPU oracle (a C restatement of the upstream CPU algorithms; the Rust prover cannot be built here)
                running the hot-path call sequence of one whole proof per step on the host cores

`--workload hotpath` times the bare MSM / NTT / quotient sequence on uniformly random columns instead (round 1's first bench).
One process per GPU; independent proofs shard with no collective (torch.distributed is used only for the barrier, the
max-over-ranks of the elapsed time and the optional split-MSM leg).
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 19
DOMINANT_KERNEL = "msm_accumulate_kernel"
METRIC = "P-256 ECDSA proofs/sec at k=19"
UNIT = "proofs/s"
N_MSM_LAGRANGE, N_MSM_G = 5, 10
N_INTT, N_EXT, N_IEXT, N_QUOT = 5, 5, 1, 1
WORKLOAD_HOT = ("ECDSA circuit k=19 (1 advice/1 lookup/1 fixed, ext 2^21), prover hot path per proof: "
                "15 MSM(2^19) + 5 iNTT(2^19) + 5 cosetNTT(2^19->2^21) + quotient(2^21 rows) + 1 iNTT(2^21), GWC opening count")
WORKLOAD = ("single P-256 ECDSA-circuit proof, k=19 (bench_ecdsa.config:1: 1 advice/1 lookup/1 fixed, ext 2^21), EVM transcript, GWC: "
            "full create_proof on the device (15 MSM(2^19), 5 iNTT, 5 coset NTT, quotient, lookup permutation, grand products, "
            "18 evaluations, 5 openings) on the P-256 ECDSA verification circuit's witness for a synthetic signed WebAuthn assertion")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_dram_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, read from the newest committed
    `ncu --set full --page raw --csv` export under profiles/ (per-launch, like `achieved`); (None, why) if absent."""
    import csv
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_msm_acc*_raw.csv")), reverse=True):
        try:
            with open(path, newline="") as f:
                rows = list(csv.reader(f))
        except OSError:
            continue
        hdr = next((r for r in rows if "Kernel Name" in r), None)
        if not hdr:
            continue
        ik = hdr.index("Kernel Name")
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        units = rows[rows.index(hdr) + 1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[rows.index(hdr) + 2:]:
            if len(r) > max(ik, ir, iw) and kernel in r[ik]:
                try:
                    vals.append(float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0))
                except ValueError:
                    pass
        if vals:
            best = (sum(vals) / len(vals), f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(vals)} launch(es) in {os.path.relpath(path, ROOT)}")
            break
    return best or (None, "no ncu raw export of this kernel under profiles/")


def modmul_peak_from_profiles():
    """Measured 254-bit Montgomery products per second of one B200 (tools/modmul_bench.cu), from the newest profiles/r*_modmul_peak.txt."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_modmul_peak.txt")), reverse=True):
        try:
            vals = [float(x) for x in re.findall(r"([0-9.]+)\s*G\s*(?:products|modmul)", open(path).read())]
            m = max(vals) if vals else None
        except OSError:
            continue
        if m:
            return m * 1e9, os.path.relpath(path, ROOT)
    return None, "no profiles/r*_modmul_peak.txt"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class HotPathProof:
    """Device-resident state for one prover: SRS + proving-key cosets (fixed per circuit) and one
    proof's worth of synthetic witness-derived columns."""

    def __init__(self, zkw, ctx, torch, k: int, seed: int):
        self.zkw, self.ctx, self.torch, self.k = zkw, ctx, torch, k
        self.shape = zkw.CircuitShape.from_config(k, 1, 1, 1)
        self.ek = self.shape.ext_k
        self.n, self.en = 1 << k, 1 << self.ek
        dev = torch.device("cuda", ctx.device)
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self.gen = g
        self.dev = dev
        # dev SRS (gen_srs analogue) — tau is a fixed development value
        tau = np.array([0x1234567890ABCDEF, 0x0FEDCBA987654321, 0x1111111111111111, 0x0222222222222222], dtype=np.uint64)
        ctx.srs_setup(k, tau)
        # proving-key cosets: constants, table, q_enable, q_lookup, 2 sigma, l0, l_last, l_active
        self.pk = {name: self.rand_fr(self.en) for name in
                   ("constants", "table", "q_enable", "q_lookup", "sigma0", "sigma1", "l0", "l_last", "l_active")}
        self.challenges = {c: self.rand_fr(1).cpu().numpy().view(np.uint64).reshape(4) for c in ("y", "beta", "gamma", "theta")}
        # per-proof columns (Lagrange basis): advice, A', S', Z_perm, Z_lookup; coefficient form: random poly, 5 GWC witnesses
        self.lagrange = [self.rand_fr(self.n) for _ in range(5)]
        self.coeff_polys = [self.rand_fr(self.n) for _ in range(6)]
        self.work = [torch.empty((self.n, 4), dtype=torch.int64, device=dev) for _ in range(5)]
        self.ext = [torch.empty((self.en, 4), dtype=torch.int64, device=dev) for _ in range(5)]
        self.h = torch.empty((self.en, 4), dtype=torch.int64, device=dev)
        self.commitments = []

    def rand_fr(self, n: int):
        """uniform field elements < 2^253 (valid Montgomery residues: any value < r is one)"""
        t = self.torch.randint(0, 1 << 62, (n, 4), dtype=self.torch.int64, device=self.dev, generator=self.gen)
        t[:, 3] &= (1 << 60) - 1
        return t

    def input_bytes(self) -> int:
        return (len(self.lagrange) + len(self.coeff_polys)) * self.n * 32

    def step(self):
        z, ctx, n, k, ek = self.zkw, self.ctx, self.n, self.k, self.ek
        out = []
        # commitments to the Lagrange-basis columns
        for col in self.lagrange:
            out.append(ctx.msm_dev(col, n, z.BASES_G_LAGRANGE))
        out.append(ctx.msm_dev(self.coeff_polys[0], n, z.BASES_G))  # random polynomial
        # to coefficient form, then onto the extended coset
        for col, w, e in zip(self.lagrange, self.work, self.ext):
            w.copy_(col)
            ctx.lagrange_to_coeff_dev(w, k)
            ctx.coeff_to_extended_dev(w, k, ek, e)
        cols = {"advice": [self.ext[0]], "constants": [self.pk["constants"]], "table": self.pk["table"],
                "q_enable": [self.pk["q_enable"]], "q_lookup": self.pk["q_lookup"],
                "sigma": [self.pk["sigma0"], self.pk["sigma1"]], "perm_z": [self.ext[3]],
                "lookup_z": [self.ext[4]], "lookup_a": [self.ext[1]], "lookup_s": [self.ext[2]],
                "l0": self.pk["l0"], "l_last": self.pk["l_last"], "l_active": self.pk["l_active"]}
        ctx.quotient_dev(self.shape, cols, self.challenges, self.h)
        ctx.extended_to_coeff_dev(self.h, ek)
        # h pieces (4 x n coefficients) and the GWC witness polynomials
        for i in range(4):
            out.append(ctx.msm_dev(self.h[i * n:(i + 1) * n], n, z.BASES_G))
        for p in self.coeff_polys[1:]:
            out.append(ctx.msm_dev(p, n, z.BASES_G))
        self.commitments = out
        return out


class HostAbiProof:
    """The same sequence through the host-pointer C ABI - what a drop-in FFI call from halo2_proofs pays per best_multiexp /
    best_fft / evaluate_h call - with every caller buffer in page-locked host memory and reused across steps (inputs AND
    outputs: the in-place transforms work on the caller's array, like upstream's `&mut [F]`)."""

    def __init__(self, zkw, ctx, torch, dev_state: HotPathProof):
        self.zkw, self.ctx, self.s = zkw, ctx, dev_state
        self.torch = torch

        def pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h.numpy().view(np.uint64)

        def pinned_empty(rows):
            return torch.empty((rows, 4), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)

        self.lagrange = [pinned(t) for t in dev_state.lagrange]
        self.coeff_polys = [pinned(t) for t in dev_state.coeff_polys]
        self.pk = {k: pinned(v) for k, v in dev_state.pk.items()}
        self.work = [pinned_empty(dev_state.n) for _ in self.lagrange]
        self.ext = [pinned_empty(dev_state.en) for _ in self.lagrange]
        self.h = pinned_empty(dev_state.en)
        self.h2d = 0
        self.d2h = 0

    def step(self):
        z, ctx, s = self.zkw, self.ctx, self.s
        n, k, ek, en = s.n, s.k, s.ek, s.en
        h2d = d2h = 0
        out = []
        for col in self.lagrange:
            out.append(ctx.msm(col, which=z.BASES_G_LAGRANGE)); h2d += n * 32; d2h += 96
        out.append(ctx.msm(self.coeff_polys[0], which=z.BASES_G)); h2d += n * 32; d2h += 96
        for col, w, e in zip(self.lagrange, self.work, self.ext):
            w[:] = col                                         # upstream's lagrange_to_coeff consumes its argument
            ctx.lagrange_to_coeff(w, inplace=True); h2d += n * 32; d2h += n * 32
            ctx.coeff_to_extended(w, ek, out=e); h2d += n * 32; d2h += en * 32
        ext = self.ext
        cols = {"advice": [ext[0]], "constants": [self.pk["constants"]], "table": self.pk["table"],
                "q_enable": [self.pk["q_enable"]], "q_lookup": self.pk["q_lookup"],
                "sigma": [self.pk["sigma0"], self.pk["sigma1"]], "perm_z": [ext[3]],
                "lookup_z": [ext[4]], "lookup_a": [ext[1]], "lookup_s": [ext[2]],
                "l0": self.pk["l0"], "l_last": self.pk["l_last"], "l_active": self.pk["l_active"]}
        h = ctx.quotient(s.shape, cols, s.challenges, out=self.h); h2d += 14 * en * 32; d2h += en * 32
        hc = ctx.extended_to_coeff(h, inplace=True); h2d += en * 32; d2h += en * 32
        for i in range(4):
            out.append(ctx.msm(hc[i * n:(i + 1) * n], which=z.BASES_G)); h2d += n * 32; d2h += 96
        for p in self.coeff_polys[1:]:
            out.append(ctx.msm(p, which=z.BASES_G)); h2d += n * 32; d2h += 96
        self.h2d, self.d2h = h2d, d2h
        return out


class FullProof:
    """One complete create_proof per step through zkw_create_proof_ex (EVM transcript, GWC) over the real P-256
    ECDSA verification circuit (csrc/ecdsa_circuit.cpp) for synthetic signed WebAuthn assertions."""

    def __init__(self, zkw, torch, k: int, device: int, seed: int):
        self.zkw, self.torch = zkw, torch
        self.state = zkw.ProverState(zkw.CircuitParams.for_degree(k), device)      # gen_srs + keygen, resident
        self.ctx = self.state.ctx
        self.assertions = [zkw.synthetic_assertion(1000 * seed + i) for i in range(4)]
        a = self.assertions[0]
        cols = self.state.circuit.synthesize(*[a[32 * i: 32 * i + 32] for i in range(5)])     # canonical integers
        self.rows = [c.shape[0] for c in cols]
        self.dev_cols = [torch.from_numpy(c.view(np.int64)).to(torch.device("cuda", device)) for c in cols]
        self.step_no = 0
        self.proof = b""
        self.h2d = sum(c.shape[0] * 32 for c in cols)      # the public API ships the assigned advice cells, 32 bytes each
        self.synth_ms = []

    def step(self):
        """advice already in HBM (canonical integers); blinding seed changes every step."""
        self.step_no += 1
        self.proof = self.zkw.create_proof(self.ctx, self.state.pk, self.dev_cols, seed=self.step_no, transcript=self.zkw.TRANSCRIPT_EVM,
                                           canonical=True, device_rows=self.rows)
        return self.proof

    def step_e2e(self):
        """the public API: assertion bytes in, proof bytes out (host witness synthesis + H2D inside)."""
        self.step_no += 1
        self.proof = self.state.prove(self.assertions[self.step_no % len(self.assertions)], self.zkw.TRANSCRIPT_EVM, seed=self.step_no)
        self.synth_ms.append(self.state.last_synth_ms)
        return self.proof


class CpuHotPath:
    """The reference arm / cpu_baseline leg: the prover's hot-path call sequence for ONE k = 19 proof on the CPU oracle
    (a restatement of the upstream CPU algorithms), every call on its own buffers:

        15 best_multiexp(2^k) on 15 distinct scalar vectors (5 over g_lagrange, 10 over g)
         5 lagrange_to_coeff(2^k), 5 coeff_to_extended(2^k -> 2^ek) on 5 distinct columns
         1 evaluate_h + divide_by_vanishing_poly over 14 distinct cosets, 1 extended_to_coeff(2^ek)

    `fraction` < 1 runs a bounded sample: ceil(fraction * count) calls of each kind (the quotient over the first
    fraction of the rows is not expressible, so it is run whole and its measured time is scaled); the value reported is
    always proofs per second = fraction_of_a_proof_done / measured_seconds, and ms_per_step is the measured time."""

    def __init__(self, k: int, threads: int):
        from oracle import cpu
        self.cpu, self.k, self.threads = cpu, k, threads
        n = 1 << k
        self.shape = cpu.make_shape(k, 1, 0, 1)
        self.ek = self.shape.ext_k
        en = 1 << self.ek
        self.dom = cpu.Domain.new(self.shape.cs_degree, k)
        self.g = cpu.g1_fixed_base_mul(cpu.fr_random(n, 1), threads)
        self.gl = cpu.g1_fixed_base_mul(cpu.fr_random(n, 2), threads)
        self.scalars = [cpu.fr_random(n, 10 + i) for i in range(N_MSM_G + N_MSM_LAGRANGE)]
        self.columns = [cpu.fr_random(n, 40 + i) for i in range(N_INTT)]
        names = ("constants", "table", "q_enable", "q_lookup", "sigma0", "sigma1", "l0", "l_last", "l_active")
        self.pk = {nm: cpu.fr_random(en, 60 + i) for i, nm in enumerate(names)}
        self.ch = {nme: cpu.fr_random(1, 90 + i)[0] for i, nme in enumerate(("y", "beta", "gamma", "theta"))}

    def step(self) -> dict:
        cpu, dom, T = self.cpu, self.dom, self.threads
        t = {}
        t0 = time.perf_counter()
        for i, s in enumerate(self.scalars):
            cpu.best_multiexp(s, self.gl if i < N_MSM_LAGRANGE else self.g, T)
        t["msm"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        coeffs = [dom.lagrange_to_coeff(c, T) for c in self.columns]
        t["intt"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ext = [dom.coeff_to_extended(c, T) for c in coeffs]
        t["ext"] = time.perf_counter() - t0
        pk = self.pk
        cols = {"advice": [ext[0]], "constants": [pk["constants"]], "table": pk["table"], "q_enable": [pk["q_enable"]],
                "q_lookup": pk["q_lookup"], "sigma": [pk["sigma0"], pk["sigma1"]], "perm_z": [ext[3]], "lookup_z": [ext[4]],
                "lookup_a": [ext[1]], "lookup_s": [ext[2]], "l0": pk["l0"], "l_last": pk["l_last"], "l_active": pk["l_active"]}
        t0 = time.perf_counter()
        h = cpu.quotient_ecdsa(self.shape, cols, self.ch, T)
        t["quot"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        dom.extended_to_coeff(h, T)
        t["iext"] = time.perf_counter() - t0
        t["total"] = sum(t.values())
        return t

    def sample(self) -> str:
        return (f"the whole hot-path sequence of one proof, each call on its own buffers: {N_MSM_G + N_MSM_LAGRANGE} best_multiexp(2^{self.k}), "
                f"{N_INTT} lagrange_to_coeff(2^{self.k}), {N_EXT} coeff_to_extended(2^{self.k}->2^{self.ek}), evaluate_h(2^{self.ek} rows, 14 cosets), "
                f"extended_to_coeff(2^{self.ek}); {self.threads} OpenMP threads; not included: witness synthesis, lookup permutation, grand products, "
                f"evaluations, openings, transcript (the GPU arm does all of those)")


def host_threads() -> int:
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1, so the OpenMP default is
    not trustworthy: the CPU arm passes this count to the oracle's num_threads clauses explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def split_msm_leg(zkw, torch, dist, rank, world, local, ks=(20, 22), reps=7):
    """One 2^k-point MSM over `world` GPUs: rank r holds slice r of the basis (2^k / world points tau_r^i G with their window
    tables), the scalars are already on each device.  Timed on the host around device MSM + NCCL all-gather + host fold,
    max over ranks, median of `reps`."""
    mg = importlib.import_module("webauthn-halo2_b200.multi_gpu")
    dev = torch.device("cuda", local)
    rows = []
    for k in ks:
        n = 1 << k
        if n % world:
            continue
        m = n // world
        ctx = zkw.Context(local)
        tau = np.array([0x1234567890ABCDEF + rank, 0x0FEDCBA987654321, 0x1111111111111111, 0x0222222222222222], dtype=np.uint64)
        ctx.srs_setup(m.bit_length() - 1, tau)                       # this rank's slice, with window tables
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 + k)                                    # the same scalar vector on every rank
        s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device=dev, generator=gen)
        s[:, 3] &= (1 << 60) - 1
        mine = s[rank * m:(rank + 1) * m].contiguous()

        def once():
            part = ctx.msm_dev(mine, m, zkw.BASES_G)
            return mg._gather_and_fold(part, world, dist, dev)

        res = once()
        ts = []
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = once()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            ts.append(float(dt.item()) * 1e3)
        ts.sort()
        # check: rank 0 gathers every slice of the basis and computes the whole MSM alone (caller bases)
        g_mine = torch.from_numpy(ctx.srs_get(zkw.BASES_G, m).view(np.int64)).to(dev)
        g_all = [torch.empty_like(g_mine) for _ in range(world)]
        dist.all_gather(g_all, g_mine)
        ok = None
        if rank == 0:
            whole = ctx.msm_dev(s, n, zkw.BASES_CALLER, bases_dev=torch.cat(g_all))
            ok = bool(np.array_equal(whole, res))
        rows.append({"k": k, "points": n, "gpus": world, "ms": ts[len(ts) // 2], "points_per_s": n / ts[len(ts) // 2] * 1e3,
                     "matches_single_gpu_msm": ok})
        ctx.close()
        del s, mine, g_all, g_mine
        torch.cuda.empty_cache()
    return {"rows": rows, "path": "multi_gpu._gather_and_fold: per-rank zkw_msm_bn254_g1_dev over resident window tables, NCCL all-gather of "
                                  "world x 96 bytes, host fold (zkw_g1_sum)"}


def config_dict(args, workload: str) -> dict:
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload, "k": args.k, "proofs_per_step_per_gpu": 1,
            "l2": "working set per step ~1.4 GB (14 cosets x 64 MiB + SRS window tables 2 x 512 MiB) exceeds the 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                    # N > 1: rank 0 alone runs the CPU arm, the other ranks exit without work
    threads = host_threads()
    cpu = CpuHotPath(args.k, threads)
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu.step()
        if i >= args.warmup:
            times.append(t)
    total = sum(t["total"] for t in times)
    value = len(times) / total                    # one step = the hot path of one whole proof, actually run
    parts = {nm: sum(t[nm] for t in times) / len(times) for nm in ("msm", "intt", "ext", "quot", "iext")}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 Montgomery (254-bit modular integers)", "data": "synthetic",
        "config": config_dict(args, WORKLOAD),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu.sample()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "seconds_per_step_by_function": parts,
        "note": ("CPU restatement of the upstream algorithms (oracle/, kind \"port\"): the Rust prover cannot be built in this image. It runs "
                 "LESS work per step than the GPU arm (hot-path functions only, uniformly random columns)"),
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    zkw = importlib.import_module("webauthn-halo2_b200")
    if args.workload == "hotpath":
        ctx = zkw.Context(local)
        state = HotPathProof(zkw, ctx, torch, args.k, seed=1234 + rank)
    else:
        state = FullProof(zkw, torch, args.k, local, seed=1234 + rank)
        ctx = state.ctx
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    torch.cuda.set_stream(stream)  # torch-side copies are ordered with the ctx's kernels
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        state.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # inside the timed region only the dominant kernel is bracketed by CUDA events (30 events per proof; bracketing all 215
    # launches costs ~1 ms per proof); the per-kernel table comes from one extra, untimed proof below
    ctx.profile_reset()
    ctx.profile_filter(DOMINANT_KERNEL)
    ctx.profile_enable(True)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        state.step()
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    prof = ctx.profile_all()
    ctx.profile_enable(False)
    ctx.profile_filter(None)
    clocks = sampler.stop() if rank == 0 else None
    ctx.profile_reset()
    ctx.profile_enable(True)
    state.step()                                   # untimed: every launch bracketed, for the per-kernel table
    prof_all = ctx.profile_all()
    ctx.profile_enable(False)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * args.steps / (max_ms / 1000.0)

    # end-to-end through the host-pointer C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e and args.workload != "hotpath":
        state.step_e2e()
        barrier()
        e2e_steps = max(1, args.steps)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            state.step_e2e()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        te = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": state.h2d,
               "d2h_bytes_per_step": len(state.proof), "steps": e2e_steps,
               "host_synthesis_ms": (sum(state.synth_ms) / len(state.synth_ms)) if state.synth_ms else None,
               "path": "generate_proof_evm mirror (ProverState.prove): assertion bytes -> ECDSA circuit witness synthesis on the host "
                       "(zkw_ecdsa_synthesize into page-locked staging, signature checked) -> one H2D copy of the assigned advice cells -> "
                       "zkw_create_proof_ex -> proof bytes"}
        # the function-level drop-in (what a [patch]ed halo2_proofs pays): the same hot-path sequence through the three
        # host-pointer seams, host<->device copies of every call inside the timed region
        if rank == 0 and args.three_seam:
            hctx = zkw.Context(local)
            hs = HotPathProof(zkw, hctx, torch, args.k, seed=77)
            host = HostAbiProof(zkw, hctx, torch, hs)
            host.step()
            t0 = time.perf_counter()
            for _ in range(2):
                host.step()
            hctx.sync()
            e2e["three_seam"] = {"value": 2 / (time.perf_counter() - t0), "unit": UNIT, "h2d_bytes_per_step": host.h2d, "d2h_bytes_per_step": host.d2h,
                                 "path": "host-pointer C ABI, one call per upstream function (15 zkw_msm_bn254_g1, 5 zkw_lagrange_to_coeff, "
                                         "5 zkw_coeff_to_extended, zkw_quotient_ecdsa, zkw_extended_to_coeff), every caller buffer in page-locked host memory"}
            del host, hs
            hctx.close()
    elif not args.no_e2e:
        host = HostAbiProof(zkw, ctx, torch, state)
        host.step()
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host.step()
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        te = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": host.h2d,
               "d2h_bytes_per_step": host.d2h, "steps": e2e_steps,
               "path": "host-pointer C ABI (zkw_msm_bn254_g1 / zkw_lagrange_to_coeff / zkw_coeff_to_extended / zkw_quotient_ecdsa / zkw_extended_to_coeff), pinned host buffers"}

    # the dominant kernel in isolation (no lanes in flight): one uniform-scalar MSM of 2^k points, the case
    # DESIGN.md's roofline arithmetic is written for
    isolated = None
    msm_windows = -(-255 // ctx.msm_window_bits(1 << args.k))      # mixed additions per point and MSM (the plan the library picks)
    if args.workload != "hotpath":
        g = torch.Generator(device=torch.device("cuda", local))
        g.manual_seed(99)
        us = torch.randint(0, 1 << 62, ((1 << args.k), 4), dtype=torch.int64, device=torch.device("cuda", local), generator=g)
        us[:, 3] &= (1 << 60) - 1
        torch.cuda.synchronize()
        for _ in range(2):
            ctx.msm_dev(us, 1 << args.k, zkw.BASES_G)
        ctx.profile_reset()
        ctx.profile_enable(True)
        for _ in range(5):
            ctx.msm_dev(us, 1 << args.k, zkw.BASES_G)
        iso = ctx.profile_all()
        ctx.profile_enable(False)
        ims, icnt = iso.get("msm_accumulate_kernel", (0.0, 0))
        if icnt:
            isolated = {"avg_launch_ms": ims / icnt, "launches": icnt, "scalars": "uniform",
                        "msm_total_ms": sum(v[0] for v in iso.values()) / icnt}
        del us

    # the reference's other flavours through the public API (host witness synthesis + H2D inside, wall clock):
    # generate_proof = Blake2b + SHPLONK (what bench_secp256r1_ecdsa and the published csv time) at k = 19 and at the
    # server's k = 17 (BASELINE configs[0]); generate_proof_evm at k = 17 (the /prove_evm production path)
    flavours = None
    if args.flavours and args.workload != "hotpath" and rank == 0:
        flavours = {}

        def time_flavour(st, transcript, shplonk, reps=5):
            for i in range(2):
                st.prove(state.assertions[i % 4], transcript, seed=i, shplonk=shplonk)
            t0 = time.perf_counter()
            for i in range(reps):
                proof = st.prove(state.assertions[i % 4], transcript, seed=100 + i, shplonk=shplonk)
            return {"ms_per_proof": (time.perf_counter() - t0) / reps * 1e3, "proof_bytes": len(proof)}

        flavours["k19_blake2b_shplonk"] = time_flavour(state.state, zkw.TRANSCRIPT_BLAKE2B, True)
        st17 = zkw.ProverState(zkw.CircuitParams.for_degree(17), local)
        flavours["k17_blake2b_shplonk"] = time_flavour(st17, zkw.TRANSCRIPT_BLAKE2B, True)
        flavours["k17_evm_gwc"] = time_flavour(st17, zkw.TRANSCRIPT_EVM, False)
        st17.close()
        st17.ctx.close()
        flavours["published_reference_ms"] = {"k19_blake2b_shplonk": 14846.2, "k17_blake2b_shplonk": 5388.0,
                                               "source": "halo2-circuits/src/results/ecdsa_bench.csv:2,4 (M1 Pro, includes halo2-ecc witness synthesis)"}

    # configs[2]-style throughput: a batch of independent proofs through the public API with several
    # provers in flight on this GPU (host witness synthesis included)
    batch = None
    if args.batch > 0 and args.workload != "hotpath":
        pool = zkw.ProverPool(zkw.CircuitParams.for_degree(args.k), local, workers=args.workers)
        assertions = [zkw.synthetic_assertion(100000 * (rank + 1) + i) for i in range(args.batch)]
        pool.prove_many(assertions[: args.workers], zkw.TRANSCRIPT_EVM)          # warm-up (arenas, lanes)
        barrier()
        t0 = time.perf_counter()
        proofs = pool.prove_many(assertions, zkw.TRANSCRIPT_EVM)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tb_ = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tb_, op=dist.ReduceOp.MAX)
        batch = {"proofs_per_gpu": args.batch, "workers_per_gpu": args.workers, "value": world * args.batch / float(tb_.item()), "unit": UNIT,
                 "path": "ProverPool.prove_many (generate_proof_evm mirror: ECDSA witness synthesis on the host + H2D inside, OS-seeded blinding)",
                 "all_proofs_distinct": len(set(proofs)) == len(proofs)}
        pool.close()

    # BASELINE configs[4], multi-GPU leg: ONE MSM of 2^20 / 2^22 points split over the ranks (each rank keeps the window
    # tables of its slice resident; NCCL all-gather of the 96-byte partial results; host fold), checked against the same
    # MSM computed by rank 0 alone
    split = None
    if world > 1 and args.split_msm and args.workload != "hotpath":
        split = split_msm_leg(zkw, torch, dist, rank, world, local)

    # BASELINE configs[4], single-GPU leg (N = 1 only): the MSM sweep 2^16 .. 2^24 and the NTT sweep 2^17 .. 2^24 against the
    # algorithmic-byte roofline, so that the driver's bench record carries them (tools/sweep.py is the same code, stand-alone)
    sweep = None
    if world == 1 and args.sweep and args.workload != "hotpath":
        spec = importlib.util.spec_from_file_location("zkw_sweep", os.path.join(ROOT, "tools", "sweep.py"))
        sw = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(sw)
        hbm_peak = peaks()[0]["hbm_gbs"]
        rows = sw.sweep_rows(zkw, 16, args.sweep_max_k, hbm_peak, local)
        torch.cuda.set_stream(stream)
        sweep = {"hbm_peak_gbs": hbm_peak,
                 "algorithmic_bytes": "MSM 96*N, NTT 64*N, coset extension 160*n (SURVEY.md 8d)",
                 "msm_ms": {str(r["k"]): round(r["msm_ms"], 4) for r in rows},
                 "msm_frac_of_hbm": {str(r["k"]): round(r["msm_frac_of_hbm"], 5) for r in rows},
                 "intt_ms": {str(r["k"]): round(r["intt_ms"], 4) for r in rows if "intt_ms" in r},
                 "intt_frac_of_hbm": {str(r["k"]): round(r["intt_frac_of_hbm"], 5) for r in rows if "intt_frac_of_hbm" in r},
                 "coset_ext_ms": {str(r["k"]): round(r["coset_ext_ms"], 4) for r in rows if "coset_ext_ms" in r}}

    if rank == 0:
        pk, pk_src = peaks()
        n = 1 << args.k
        acc_ms, acc_cnt = prof.get(DOMINANT_KERNEL, (0.0, 0))
        per_launch_ms = acc_ms / acc_cnt if acc_cnt else None
        alg_bytes = 96 * n
        achieved = (alg_bytes / (per_launch_ms / 1000.0) / 1e9) if per_launch_ms else None
        total_kernel_ms = sum(v[0] for v in prof_all.values())                  # share: from the fully bracketed proof
        traffic, traffic_src = ncu_dram_traffic("msm_accumulate_kernel")
        modmul_peak, modmul_src = modmul_peak_from_profiles()
        roofline = {
            "kernel": "msm_accumulate_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": (achieved / pk["hbm_gbs"]) if achieved else None,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": pk_src + " copy bandwidth (MEASURED_PEAKS.json)",
            "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": per_launch_ms, "launches": acc_cnt,
            "share_of_kernel_time": (prof_all.get(DOMINANT_KERNEL, (0.0, 0))[0] / total_kernel_ms) if total_kernel_ms else None,
            "note": ("integer-ALU bound (254-bit Montgomery products), see DESIGN.md. avg_launch_ms is over the timed region, where "
                     "MSM lanes and the NTT stream overlap this kernel and witness-shaped scalars make launches uneven; "
                     "`isolated` times the same kernel alone on uniform scalars"),
            "isolated": isolated,
            "isolated_achieved_gbs": (alg_bytes / (isolated["avg_launch_ms"] / 1000.0) / 1e9) if isolated else None,
            "msm_windows": msm_windows,
            "isolated_modmul_per_s": (msm_windows * n * 10 / (isolated["avg_launch_ms"] / 1000.0)) if isolated else None,
            "modmul_peak_per_s_measured": modmul_peak, "modmul_peak_source": modmul_src,
        }
        kernels = {name: {"ms_per_step": v[0], "launches_per_step": v[1]} for name, v in sorted(prof_all.items())}
        cpu_val, cpu_sample = None, "skipped"
        cores = None
        if world == 1 and not args.no_cpu:
            cores = host_threads()
            cpu = CpuHotPath(args.k, cores)
            cpu.step()                                         # warm-up (page faults, OpenMP pool)
            reps = [cpu.step()["total"] for _ in range(2)]     # ~2 x 7 s on 16 cores: a bounded sample of the reference arm's step
            cpu_val, cpu_sample = len(reps) / sum(reps), cpu.sample() + f"; {len(reps)} timed steps after one warm-up"
            del cpu
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 Montgomery (254-bit modular integers)", "data": "synthetic",
            "config": config_dict(args, WORKLOAD if args.workload != "hotpath" else WORKLOAD_HOT),
            "proof_bytes": len(getattr(state, "proof", b"")),
            "circuit": state.state.circuit.stats() if args.workload != "hotpath" else None,
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample},
            "e2e": e2e, "batch": batch, "split_msm": split, "sweep": sweep, "flavours": flavours, "gpu_launches": launches, "clocks": clocks, "kernels": kernels,
            "published_reference": {"value": 1.0 / 14.846241542, "unit": UNIT, "hardware": "M1 Pro (halo2-circuits/src/results/ecdsa_bench.csv:2)",
                                    "note": "full create_proof incl. halo2-ecc witness synthesis, Blake2b + SHPLONK; not the same hardware or witness"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--workload", default="proof", choices=["proof", "hotpath"])
    ap.add_argument("--batch", type=int, default=64, help="proofs per GPU in the batch-throughput leg: BASELINE configs[2] (64 on one GPU) and [3] (512 over 8); 0 = skip")
    ap.add_argument("--workers", type=int, default=3, help="concurrent provers per GPU in the batch leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-flavours", dest="flavours", action="store_false", help="skip the Blake2b/SHPLONK and k = 17 timings")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-split-msm", dest="split_msm", action="store_false", help="N > 1: skip the single-MSM-split-over-ranks leg")
    ap.add_argument("--no-three-seam", dest="three_seam", action="store_false", help="skip the host-pointer three-seam figure")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="N = 1: skip the MSM / NTT size sweep (BASELINE configs[4])")
    ap.add_argument("--sweep-max-k", type=int, default=24)
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the
    # first collective), so file descriptor 1 is pointed at stderr for the whole run and the JSON line alone goes
    # to the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
