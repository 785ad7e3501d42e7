#!/usr/bin/env python
"""Development aid: scan the same synthetic chromosome with two builds of the library (RMD_LIB) and name the (window, model)
groups whose gate passes or hit counts differ.

  python tools/ab_scan.py --mb 10 --old rmdetect_b200/librmd_old.so [--steps 3]
"""
import argparse
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=10.0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--old", default="")
ap.add_argument("--child", default="")
args = ap.parse_args()
if args.child:
    from bench import LN_CUTOFF, LN_UNKNOWN, MIN_BPP, SEED, WIN_LEN, WIN_STEP, load_models
    from rmdetect_b200 import _native
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.scanner import ScanEngine
    if os.environ.get("RMD_LIB"):
        _native.LIB_PATH = os.path.abspath(os.environ["RMD_LIB"])
    models, evs = load_models()
    eng = ScanEngine(models, evs, LN_UNKNOWN)
    n = int(args.mb * 1e6)
    gc = gc_table({b: math.log(0.25) for b in "ACGU"})
    n_win, n_bpp, ms_gen = eng.ctx.synth_job(n, 0, SEED, SEED, WIN_LEN, WIN_STEP, MIN_BPP, LN_CUTOFF, gc)
    out = {}
    for k in range(args.steps):
        r = eng.ctx.scan_resident(0, n_win)
        print(args.child, "step", k, "scan %.3f ms" % r["ms_scan"], "tries", r["tries"], "raw", r["n_raw"], "hits", len(r["hits"]), flush=True)
        out["gate%d" % k] = r["gate_pass"]; out["raw%d" % k] = r["raw_count"]
    np.savez(args.child, **out)
    eng.close()
    sys.exit(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
files = []
for name, lib in (("new", ""), ("old", args.old)):
    if name == "old" and not lib:
        continue
    env = dict(os.environ)
    if lib:
        env["RMD_LIB"] = lib
    f = os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % name)
    subprocess.check_call([sys.executable, __file__, "--mb", str(args.mb), "--steps", str(args.steps), "--child", f], env=env)
    files.append(np.load(f))
new = files[0]
ref = files[1] if len(files) > 1 else None
for k in range(args.steps):
    for other, label in ((new, "new step 0"), (ref, "old step %d" % k)):
        if other is None:
            continue
        kk = 0 if other is new else k
        for what in ("gate", "raw"):
            a, b = new["%s%d" % (what, k)], other["%s%d" % (what, kk)]
            d = np.argwhere(a != b)
            print("new step %d vs %s: %s differs in %d groups" % (k, label, what, len(d)), [(int(w), int(m), int(a[w, m]), int(b[w, m])) for w, m in d[:12]])
