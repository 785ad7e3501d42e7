#!/usr/bin/env python
"""Per-source-line summary of an ncu report: warp instructions executed, stall samples, shared wavefronts.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40] [--kernel scan_fast]

Runs `ncu -i <rep> --page source --csv --print-source cuda,sass` (needs -lineinfo at compile time and
--import-source on at capture time) and sums the per-line rows.
"""
import argparse
import csv
import io
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--kernel", default="")
ap.add_argument("--sort", default="inst", choices=["inst", "samples"])
args = ap.parse_args()
out = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, func, hdr = "", "", None
agg = {}
tot_inst = tot_samp = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
        ct = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or not r[0] or (args.kernel and args.kernel not in func):
        continue
    try:
        inst, samp, wave, tinst = int(r[ci]), int(r[cs]), int(r[cw]), int(r[ct])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, 0, r[1].strip()])
    a[0] += inst; a[1] += samp; a[2] += wave; a[3] += tinst
    tot_inst += inst; tot_samp += samp
k = 0 if args.sort == "inst" else 1
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
print("%-22s %5s %7s %7s %6s %9s  %s" % ("file", "line", "inst%", "samp%", "lanes", "smem_wave", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][k])[:args.top]:
    print("%-22s %5d %6.2f%% %6.2f%% %6.1f %9d  %s" % (key[0], key[1], 100.0 * a[0] / max(tot_inst, 1),
                                                      100.0 * a[1] / max(tot_samp, 1), a[3] / max(a[0], 1), a[2], a[4][:110]))
