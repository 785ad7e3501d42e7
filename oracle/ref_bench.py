#!/usr/bin/env python
"""Time the reference's OWN Python code (staged in oracle/_ref by oracle/stage_ref.py) on this host.  TEST / BENCH
INFRASTRUCTURE: run as a subprocess by bench.py's cpu_baseline leg, never imported by the product.

    python oracle/ref_bench.py --windows 20 --procs 8 --cal-samples 4000

Scan: `lib/scanner.py: Scanner.scan` (unmodified, shims D1-D3 of SURVEY.md §0.1 applied by monkey-patch) on the
first `--windows` windows of bench.py's synthetic chromosome with the same stand-in bpp provider, four models; first on
one core, then one such sample per process on `--procs` cores (the reference has no parallelism of its own; window
ranges are independent).  Calibration: the body of `rmcalibrate.py` (its own main loop, rpy2 stubbed — defect D8) for
KT_1.0 on `--cal-samples` samples.  Prints one JSON object.
"""
import argparse
import contextlib
import io
import json
import math
import multiprocessing as mp
import os
import random
import sys
import time
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SEED = 20261019
MODEL_ORDER = ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]


def _import_reference():
    warnings.simplefilter("ignore")
    sys.dont_write_bytecode = True
    for p in (os.path.dirname(HERE), REF, os.path.join(REF, "lib")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import config as ref_config
    import scanner as ref_scanner
    from lib.bpprobs import BPProbs
    return ref_config, ref_scanner, BPProbs


def scan_sample(args):
    first_window, n_windows = args
    ref_config, ref_scanner, BPProbs = _import_reference()
    from rmdetect_b200.synth import SynthFold, synth_sequence
    cfg = ref_config.Config()
    cfg.data_path = os.path.join(REF, "models")
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        cfg.configure()
    by = {m.full_name(): (m, e) for m, e in zip(cfg.models, cfg.evalues)}
    cfg.models = [by[n][0] for n in MODEL_ORDER]
    cfg.evalues = [by[n][1] for n in MODEL_ORDER]
    for m in cfg.models:
        m.pairs_list = list(m.pairs_list)                        # shim D1
    cfg.win_len, cfg.win_step, cfg.min_bpp_mean = 150, 75, 0.001
    cfg.cutoff, cfg.unknown_prob = math.log(0.1), math.log(1e-4)
    sc = ref_scanner.Scanner(cfg)
    sc.rna_tools = SynthFold(SEED, factory=BPProbs)              # shims D2 / D3: the bpp provider bench.py uses
    n = (n_windows - 1) * 75 + 150
    seq = synth_sequence(n, SEED, start=first_window * 75)
    box = types.SimpleNamespace(sequence_list=[types.SimpleNamespace(key="s", seq=seq, col_index=None)])
    # fold time is not part of the metric (ViennaRNA's place): pre-compute the matrices, then time the scan alone
    cache = {}
    for k in range(n_windows):
        w = seq[k * 75:k * 75 + 150]
        cache[w] = sc.rna_tools.bpprobs(w, "")
    sc.rna_tools = types.SimpleNamespace(bpprobs=lambda w, c: cache[w])
    t0 = time.perf_counter()
    cands, tries = sc.scan(box)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, tries=list(tries), hits=len(cands), nt_model=n_windows * 75 * len(cfg.models))


def calibrate_sample(n_samples):
    _import_reference()
    stub = types.ModuleType("rpy2")
    stub.robjects = types.ModuleType("rpy2.robjects")
    stub.robjects.r = None
    sys.modules["rpy2"], sys.modules["rpy2.robjects"] = stub, stub.robjects       # shim D8
    src = open(os.path.join(REF, "rmcalibrate.py")).read()
    argv = sys.argv
    sys.argv = ["rmcalibrate.py", "--data-path", os.path.join(REF, "models"), "--samples", str(n_samples), "KT_1.0",
                "/tmp/ref_bench_%d.evalues" % os.getpid()]
    random.seed(SEED)
    g = {"__name__": "__main__"}
    t0 = time.perf_counter()
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, "rmcalibrate.py", "exec"), g)
    dt = time.perf_counter() - t0
    sys.argv = argv
    return dict(seconds=dt, samples=n_samples)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=20)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--cal-samples", type=int, default=4000)
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "lib")):
        print(json.dumps({"unavailable": "oracle/_ref is not staged (run oracle/stage_ref.py where /root/reference exists)"}))
        return
    one = scan_sample((0, a.windows))
    out = {"scan_1core": dict(one, value=one["nt_model"] / one["seconds"], unit="nt·model/s", cores=1,
                              sample="lib/scanner.py Scanner.scan, first %d windows, 4 models, bpp precomputed" % a.windows)}
    if a.procs > 1:
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(a.procs) as pool:
            parts = pool.map(scan_sample, [(k * a.windows, a.windows) for k in range(a.procs)])
        wall = time.perf_counter() - t0
        busy = max(p["seconds"] for p in parts)
        out["scan_all_cores"] = dict(value=sum(p["nt_model"] for p in parts) / busy, unit="nt·model/s", cores=a.procs,
                                     seconds=busy, wall_incl_import=wall,
                                     sample="%d processes x %d windows each (consecutive window ranges)" % (a.procs, a.windows))
    if a.cal_samples > 0:
        c = calibrate_sample(a.cal_samples)
        out["calibrate_1core"] = dict(c, value=c["samples"] / c["seconds"], unit="samples/s", cores=1,
                                      sample="rmcalibrate.py main loop, KT_1.0, %d samples" % a.cal_samples)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
