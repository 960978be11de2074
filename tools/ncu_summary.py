"""Key metrics of every launch in an .ncu-rep (via `ncu -i ... --page raw --csv`), one line per launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--raw-out profiles/x_raw.csv]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if "--raw-out" in sys.argv:
    open(sys.argv[sys.argv.index("--raw-out") + 1], "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
for d in data:
    out = [d[ix["Kernel Name"]].split("(")[0]]
    for m, lab in cols:
        if m in ix:
            out.append(f"{lab}={d[ix[m]]}{units[ix[m]] if lab in ('time', 'dram_rd', 'dram_wr') else ''}")
    print("  ".join(out))
