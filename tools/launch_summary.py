"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-proof table
(markdown on stdout): python tools/launch_summary.py gpurun_out/r1c_launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [r[4].split('(')[0].replace('zkw::', '') for r in rows]
t = [float(r[14]) / 1e3 for r in rows]   # us
starts = [i for i, n in enumerate(names) if n in ('to_mont_kernel', 'u64_to_mont_kernel') and i + 1 < len(names) and names[i + 1] in ('zero_fill_kernel', 'rand_fill_kernel')]
proofs = [(s, e) for s, e in zip(starts, starts[1:] + [len(names)])]
s, e = proofs[1]    # the first timed proof (after one warm-up); later segments also hold the isolated MSMs and other flavours
agg = collections.OrderedDict()
for n, x in zip(names[s:e], t[s:e]):
    a = agg.setdefault(n, [0, 0.0, 0.0]); a[0] += 1; a[1] += x; a[2] = max(a[2], x)
tot = sum(a[1] for a in agg.values())
print(f"{len(proofs)} proofs in the list; one timed proof: {e - s} launches, {tot / 1e3:.2f} ms of kernel time when serialised by ncu\n")
print("| kernel | launches | avg us | max us | total ms | share |\n|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {a[0]} | {a[1] / a[0]:.1f} | {a[2]:.1f} | {a[1] / 1e3:.2f} | {100 * a[1] / tot:.1f}% |")
