#!/usr/bin/env python3
"""Per-opcode executed warp-instruction histogram of an ncu report (first kernel in it).
usage: tools/ncu_sass.py report.ncu-rep [divisor] [--dump]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
div = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.Counter()
samp = collections.Counter()
tot = 0.0
lines = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        try:
            n = float(r[hdr.index("Instructions Executed")].replace(",", ""))
            sm = float(r[hdr.index("# Samples")].replace(",", ""))
        except ValueError:
            continue
        src = r[hdr.index("Source")]
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        agg[op] += n
        samp[op] += sm
        tot += n
        lines.append((n, sm, src))
print("total executed warp-instructions / divisor: %.1f" % (tot / div))
ts = sum(samp.values()) or 1.0
for op, n in agg.most_common(25):
    print("%-10s %9.1f  %5.1f%%   stall-samples %5.1f%%" % (op, n / div, 100 * n / tot, 100 * samp[op] / ts))
if "--dump" in sys.argv:
    for n, sm, src in lines:
        print("%10.1f %6d  %s" % (n / div, sm, src))
