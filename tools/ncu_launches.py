#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`).

  python tools/ncu_launches.py gpurun_out/launches.csv > profiles/rNN_launches_bench.csv
"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[kn])
    ns = float(r[mv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[r[mu]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print("kernel,launches,total_ms,share_pct,avg_ms")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.3f,%.2f,%.4f" % (name, n, ns / 1e6, 100.0 * ns / total, ns / n / 1e6))
