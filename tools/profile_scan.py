#!/usr/bin/env python
"""Small driver for profiling: synthetic chromosome of --mb Mb generated on the device, --steps scans.

Used under ncu (launch list / --set full) and for phase-breakdown experiments (RMD_DEBUG bits:
1 = skip scoring, 2 = skip the exact gate; results are then wrong on purpose).
"""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LN_CUTOFF, LN_UNKNOWN, MIN_BPP, SEED, WIN_LEN, WIN_STEP, load_models   # noqa: E402
from rmdetect_b200 import _native                                                          # noqa: E402
from rmdetect_b200.gcx import gc_table                                                    # noqa: E402
from rmdetect_b200.scanner import ScanEngine                                              # noqa: E402

if os.environ.get("RMD_LIB"):                       # A/B runs of differently built libraries (development only)
    _native.LIB_PATH = os.path.abspath(os.environ["RMD_LIB"])
ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--models", default="")
ap.add_argument("--n-frac", type=float, default=0.0, help="also scan the same job with this fraction of the letters replaced by runs of N")
ap.add_argument("--e2e", action="store_true", help="also time rmd_scan from host buffers (compact transport) and print its parts")
args = ap.parse_args()
models, evs = load_models()
if args.models:
    keep = [i for i, m in enumerate(models) if m.name in args.models.split(",")]
    models, evs = [models[i] for i in keep], [evs[i] for i in keep]
eng = ScanEngine(models, evs, LN_UNKNOWN)
n = int(args.mb * 1e6)
gc = gc_table({b: math.log(0.25) for b in "ACGU"})
n_win, n_bpp, ms_gen = eng.ctx.synth_job(n, 0, SEED, SEED, WIN_LEN, WIN_STEP, MIN_BPP, LN_CUTOFF, gc)
for k in range(args.steps):
    r = eng.ctx.scan_resident(0, n_win)
print("models %s windows %d bpp %d gen %.2f ms | scan %.3f ms post %.3f ms total %.3f ms | tries %s raw %d hits %d | %.1f M nt·model/s"
      % ([m.name for m in models], n_win, n_bpp, ms_gen, r["ms_scan"], r["ms_post"], r["ms_total"], r["tries"], r["n_raw"],
         len(r["hits"]), n * len(models) / r["ms_scan"] / 1e3))
if args.n_frac > 0:
    # the same chromosome with runs of N (real genomes: assembly gaps, masked repeats); lists unchanged
    off, bi, bj, bp = eng.ctx.download_job(n_win, n_bpp)
    codes = eng.ctx.download_codes(n).copy()
    rng = np.random.default_rng(1)
    run = 200
    for at in rng.integers(0, n - run, int(args.n_frac * n / run)):
        codes[at:at + run] = 5
    counts = np.bincount(codes, minlength=16)
    gcn = gc_table({"ACGUN"[min(c, 4)]: math.log(counts[c] / n) for c in (0, 1, 2, 3, 5) if counts[c]}, others="N")
    job = dict(seq_off=np.array([0, n], dtype=np.int64), codes=codes, gc_log=gcn.reshape(1, -1),
               gc_class=np.full((1, len(models)), -1, dtype=np.int32), win_len=WIN_LEN, win_step=WIN_STEP, n_win=n_win,
               bpp_off=off, bpp_i=bi, bpp_j=bj, bpp_p=bp, bpp_ij=None, bpp_q=None, min_bpp_mean=MIN_BPP, cutoff=LN_CUTOFF)
    eng.ctx.upload(job)
    for k in range(args.steps):
        r2 = eng.ctx.scan_resident(0, n_win)
    n_other = int(sum(1 for w in range(n_win) if (codes[w * WIN_STEP:w * WIN_STEP + WIN_LEN] > 3).any())) if n_win < 200000 else -1
    print("with %.2f %% N in runs of %d: %d of %d windows hold an N | scan %.3f ms post %.3f ms | tries %s raw %d | %.1f M nt·model/s"
          % (100 * args.n_frac, run, n_other, n_win, r2["ms_scan"], r2["ms_post"], r2["tries"], r2["n_raw"], n * len(models) / r2["ms_scan"] / 1e3))
if args.e2e:
    off, bi, bj, bp = eng.ctx.download_job(n_win, n_bpp, pinned=True)
    codes = _native.pinned_array(n, np.uint8)
    codes[:] = eng.ctx.download_codes(n)
    ij, q = _native.compact_bpp(bi, bj, bp, pinned=True)
    job = dict(seq_off=np.array([0, n], dtype=np.int64), codes=codes, gc_log=gc.reshape(1, -1),
               gc_class=np.full((1, len(models)), -1, dtype=np.int32), win_len=WIN_LEN, win_step=WIN_STEP, n_win=n_win,
               bpp_off=off, bpp_i=None, bpp_j=None, bpp_p=None, bpp_ij=ij, bpp_q=q, min_bpp_mean=MIN_BPP, cutoff=LN_CUTOFF)
    for k in range(args.steps):
        e = eng.ctx.scan(job, copy=False)
    print("e2e: total %.3f ms | h2d(first stage) %.3f  scan %.3f  post %.3f  d2h %.3f | launches %d"
          % (e["ms_total"], e["ms_h2d"], e["ms_scan"], e["ms_post"], e["ms_d2h"], e["n_launches"]))
    for k in range(args.steps):                     # the job stays resident in the compact form: the FUSED kernel without any copy
        r3 = eng.ctx.scan_resident(0, n_win)
    print("resident after rmd_scan (values rebuilt per window, FUSED): scan %.3f ms" % r3["ms_scan"])
eng.close()
