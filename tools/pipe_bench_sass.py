"""Instruction mix of each test kernel of tools/pipe_bench.cu, from the SASS of the cross-compiled binary (no GPU needed):
proves which mnemonic each timing line of profiles/r*_pipe_rates.txt measures.  The loops are fully unrolled 64 x 8, so a
kernel's dominant mnemonics ARE its loop body.   python tools/pipe_bench_sass.py [binary] > profiles/r2_pipe_sass.txt"""
import collections
import re
import subprocess
import sys

NAMES = ["mad.wide.u32", "mad.lo.cc+madc.hi.cc pair", "mad.lo.u32", "mad.hi.u32", "fma.rz.f64", "add.cc+addc pair", "lop3", "mad.wide + add.cc/addc",
         "dfma + mad.wide", "4-pair carry chain"]


def main():
    binary = sys.argv[1] if len(sys.argv) > 1 else "tools/bin/pipe_bench"
    sass = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True, check=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            counts[fn][m.group(1)] += 1
    for fn, c in sorted(counts.items(), key=lambda kv: int(re.search(r"ILi(\d+)E", kv[0]).group(1))):
        idx = int(re.search(r"ILi(\d+)E", fn).group(1))
        body = ", ".join(f"{op} x{n}" for op, n in c.most_common() if n >= 100)
        print(f"test {idx} ({NAMES[idx]}): {body}")


if __name__ == "__main__":
    main()
