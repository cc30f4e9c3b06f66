#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: warp-instructions per env, stall samples, active lanes.
usage: tools/ncu_lines.py report.ncu-rep [num_envs] [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
envs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
src = {}


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and r and r[0] not in ("Function Name", "") and len(r) > 2 and r[2] == "-":
        ln = int(r[0])
        ie, sm, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        a = agg[(cur, ln)]
        a[0] += f(r[ie]); a[1] += f(r[sm]); a[2] += f(r[ti])
        src[(cur, ln)] = r[1]
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values()) or 1.0
print("total warp-instructions per env: %.1f   samples: %d" % (tot / envs, tots))
for (fl, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%s:%4d inst=%7.1f/env (%4.1f%%) samp=%4.1f%% lanes=%4.1f | %s" %
          (fl, ln, v[0] / envs, 100 * v[0] / tot, 100 * v[1] / tots, v[2] / max(v[0], 1), src[(fl, ln)][:95]))
