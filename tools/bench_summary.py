#!/usr/bin/env python
"""The headline fields of a bench.py line (the last JSON line of a file): python tools/bench_summary.py file.json"""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    c, e = d.get("config", {}), d.get("e2e", {})
    print("%s: N=%s value %.3f G (%.3f ms/step, kernel %s ms) | e2e %.3f G (%.3f ms) | issue %.3f | cal %s"
          % (path, d.get("n_gpus"), d["value"] / 1e9, d.get("ms_per_step", 0.0), c.get("scan_kernel_ms"),
             e.get("value", 0.0) / 1e9, e.get("ms_per_step", 0.0), d.get("roofline", {}).get("issue", {}).get("frac", 0.0),
             d.get("calibration", {}).get("value")))
