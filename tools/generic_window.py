#!/usr/bin/env python
"""Scan kernel time for windows the shared-memory kernel does not take (--win-len above 160): scan_generic_kernel.

  python tools/generic_window.py [--kb 150]
"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LN_CUTOFF, LN_UNKNOWN, MIN_BPP, SEED, load_models          # noqa: E402
from rmdetect_b200.scanner import ScanEngine                                 # noqa: E402
from rmdetect_b200.synth import SynthFold, synth_sequence                    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kb", type=int, default=150)
args = ap.parse_args()
models, evs = load_models()
eng = ScanEngine(models, evs, LN_UNKNOWN)
seq = synth_sequence(args.kb * 1000, SEED)
for W, step in ((150, 75), (160, 80), (200, 100), (300, 150)):
    job = eng.build_job([("chr", seq, None)], W, step, SynthFold(SEED), "", MIN_BPP, LN_CUTOFF)
    eng.ctx.upload(job)
    for _ in range(3):
        r = eng.ctx.scan_resident(0, job["n_win"])
    print("win %d/%d: %d windows, %d entries, tries %s, %d hits | scan %.3f ms = %.1f M nt·model/s"
          % (W, step, job["n_win"], len(job["bpp_p"]), r["tries"], len(r["hits"]), r["ms_scan"], len(seq) * len(models) / r["ms_scan"] / 1e3))
eng.close()
