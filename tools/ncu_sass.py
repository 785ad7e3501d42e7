#!/usr/bin/env python
"""SASS of a source-line range from `ncu --page source --csv --print-source cuda,sass` output, with
executions per `--per` (e.g. loop iterations) and average active lanes.
  python tools/ncu_sass.py cs.csv file.cuh first last [--per N] [--min X]"""
import csv, sys
args = sys.argv[1:]
path, fname, a, b = args[0], args[1], int(args[2]), int(args[3])
per = float(args[args.index("--per") + 1]) if "--per" in args else 1.0
mn = float(args[args.index("--min") + 1]) if "--min" in args else 0.0
cur = hdr = line = None
out, seen, tot = [], set(), 0.0
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); ct = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or r[0] == "Function Name": continue
    if r[0]: line = (cur, int(r[0])); continue
    if line and line[0] == fname and a <= line[1] <= b and r[2] not in seen:
        try: n = int(r[ci]); tn = int(r[ct])
        except ValueError: continue
        seen.add(r[2])
        if n / per >= mn: out.append((r[2], line[1], n / per, tn / max(n, 1), r[3].strip()))
for adr, l, x, lanes, s in sorted(out):
    tot += x
    print("%s L%3d %6.2f %4.1f  %s" % (adr[-5:], l, x, lanes, s))
print("total", tot)
