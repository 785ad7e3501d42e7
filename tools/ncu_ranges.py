#!/usr/bin/env python
"""Instruction and stall-sample shares of an ncu report per source-line range.

  python tools/ncu_ranges.py rep.ncu-rep file.cuh name:first-last [name:first-last ...]
"""
import csv, io, subprocess, sys
rep, fname, specs = sys.argv[1], sys.argv[2], sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, inst, samp = None, None, {}, {}
stalls = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
        scols = [(k, n) for k, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or not r[0] or r[0] == "Function Name": continue
    try:
        key = (cur, int(r[0])); inst[key] = inst.get(key, 0) + int(r[ci]); samp[key] = samp.get(key, 0) + int(r[cs])
        st = stalls.setdefault(key, {})
        for k, n in scols: st[n] = st.get(n, 0) + int(r[k])
    except ValueError: pass
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp instructions %d, samples %d" % (ti, ts))
covered = set()
for spec in specs:
    name, rng = spec.split(":"); a, b = map(int, rng.split("-"))
    keys = [k for k in inst if k[0] == fname and a <= k[1] <= b]
    covered.update(keys)
    agg = {}
    for k in keys:
        for n, v in stalls.get(k, {}).items(): agg[n] = agg.get(n, 0) + v
    top = ", ".join("%s %.1f%%" % (n.replace("stall_", ""), 100.0 * v / max(ts, 1)) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4])
    print("%-22s inst %6.2f%%  samples %6.2f%%   [%s]" % (name, 100.0 * sum(inst[k] for k in keys) / ti, 100.0 * sum(samp[k] for k in keys) / ts, top))
rest = [k for k in inst if k not in covered]
print("%-22s inst %6.2f%%  samples %6.2f%%" % ("(elsewhere)", 100.0 * sum(inst[k] for k in rest) / ti, 100.0 * sum(samp[k] for k in rest) / ts))
