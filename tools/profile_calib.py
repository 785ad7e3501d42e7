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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
models, evs = load_models()
eng = ScanEngine(models, evs, LN_UNKNOWN)
cl = class_logs(GC.get_classes(5, 15, 85))
out = []
for mi, m in enumerate(models):
    h = np.zeros((len(cl), 512))
    eng.ctx.calibrate(mi, None, cl, LN_CUTOFF, h, seed=SEED, first=0, n=1 << 14, L=len(m.nodes))
    best, wall = 1e9, 1e9
    for rep in range(3):
        h[:] = 0
        t0 = time.perf_counter()
        n_ok, ms = eng.ctx.calibrate(mi, None, cl, LN_CUTOFF, h, seed=SEED, first=0, n=n, L=len(m.nodes))
        wall = min(wall, (time.perf_counter() - t0) * 1e3)
        best = min(best, ms)
    out.append("%s L=%d %.3f ms (call %.3f ms) ok %d sum %.9e" % (m.name, len(m.nodes), best, wall, n_ok, h.sum()))
print(" | ".join(out))
