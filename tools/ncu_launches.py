"""Summarise an `ncu --csv` launch list: per kernel, mean of each metric and the launch count."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    agg[r[ki]][r[mi]].append(float(r[vi].replace(",", "")))
for k, m in agg.items():
    if "at::" in k or "fp64_probe" in k:
        continue
    print(k[:70].ljust(70), " ".join("%s=%.4g(x%d)" % (a.split("__")[-1][:28], sum(v) / len(v), len(v)) for a, v in m.items()))
