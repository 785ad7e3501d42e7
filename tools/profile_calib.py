#!/usr/bin/env python
"""Times calibrate_kernel: 2^20 device-drawn samples per shipped model against the 165 GC classes (development aid;
used under ncu for profiles/r*_calibrate_ncu_full.csv)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LN_CUTOFF, LN_UNKNOWN, SEED, load_models   # noqa: E402
from rmdetect_b200 import _native                             # noqa: E402
from rmdetect_b200.calibrate import class_logs               # noqa: E402
from rmdetect_b200.gcx import GC                              # noqa: E402
from rmdetect_b200.scanner import ScanEngine                  # noqa: E402

if os.environ.get("RMD_LIB"):
    _native.LIB_PATH = os.path.abspath(os.environ["RMD_LIB"])
flags = [a for a in sys.argv[1:] if a.startswith("--")]          # --fresh: a new histogram array per call; --job: a 10 Mb job resident; --torch: torch's CUDA context too
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 1 << 20
if "--torch" in flags:
    import torch
    torch.zeros(1, device="cuda")
models, evs = load_models()
eng = ScanEngine(models, evs, LN_UNKNOWN)
cl = class_logs(GC.get_classes(5, 15, 85))
if "--job" in flags:
    import math
    from bench import LN_CUTOFF as _c, MIN_BPP, WIN_LEN, WIN_STEP
    from rmdetect_b200.gcx import gc_table
    eng.ctx.synth_job(10_000_000, 0, SEED, SEED, WIN_LEN, WIN_STEP, MIN_BPP, _c, gc_table({b: math.log(0.25) for b in "ACGU"}))
out = []
for mi, m in enumerate(models):
    h = np.zeros((len(cl), 512))
    eng.ctx.calibrate(mi, None, cl, LN_CUTOFF, h, seed=SEED, first=0, n=1 << 14, L=len(m.nodes))
    best, wall, walls = 1e9, 1e9, []
    for rep in range(3):
        if "--fresh" in flags:
            h = np.zeros((len(cl), 512))
        else:
            h[:] = 0
        t0 = time.perf_counter()
        n_ok, ms = eng.ctx.calibrate(mi, None, cl, LN_CUTOFF, h, seed=SEED, first=0, n=n, L=len(m.nodes))
        walls.append(round((time.perf_counter() - t0) * 1e3, 3))
        wall = min(walls)
        best = min(best, ms)
    out.append("%s L=%d %.3f ms (calls %s ms) ok %d sum %.9e" % (m.name, len(m.nodes), best, walls, n_ok, h.sum()))
print(" | ".join(out))
