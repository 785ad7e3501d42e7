#!/usr/bin/env python
"""BASELINE configs[3] at its full size, from the reference itself: the body of rmcalibrate.py run in-process for KT_1.0
with --samples 1000000 and a seeded RNG (the reference's own loop, checkpoints and stop rule; about five minutes of one
core), the `.evalues` text it leaves behind and the number of simulations it ran.

  python tests/golden/make_calibrate_full.py        (build container, reference at /root/reference, read-only)

Shim D8 of SURVEY.md section 0.1 (an empty `rpy2` module: the script imports it for plots it does not draw here); nothing
of the reference is copied or edited.
"""
import contextlib
import gzip
import hashlib
import io
import json
import os
import random
import re
import sys
import time
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/lib")
SEED = 20261019

stub = types.ModuleType("rpy2")
stub.robjects = types.ModuleType("rpy2.robjects")
stub.robjects.r = None
sys.modules["rpy2"] = stub
sys.modules["rpy2.robjects"] = stub.robjects
src = open(REF + "/rmcalibrate.py").read()
out = {}
for name, samples in (("KT_1.0", 1000000),):
    tmp = "/tmp/golden_full_%s.evalues" % name
    argv = sys.argv
    sys.argv = ["rmcalibrate.py", "--data-path", REF + "/models", "--samples", str(samples), name, tmp]
    random.seed(SEED)
    g = {"__name__": "__main__"}
    err, so = io.StringIO(), io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stderr(err), contextlib.redirect_stdout(so):
        exec(compile(src, "rmcalibrate.py", "exec"), g)
    sys.argv = argv
    text = open(tmp).read()
    said = re.findall(r"Converged after (\d+) simulations", so.getvalue() + err.getvalue())
    saved = re.findall(r"data saved after (\d+) simulations", err.getvalue() + so.getvalue())
    out[name] = dict(samples=samples, seed=SEED, text=text, sha256=hashlib.sha256(text.encode()).hexdigest(),
                     converged_after=int(said[-1]) if said else None, last_saved_after=int(saved[-1]) if saved else None,
                     reference_seconds=round(time.time() - t0, 1))
    print(name, samples, "converged after", out[name]["converged_after"], "last save after", out[name]["last_saved_after"],
          "rows", text.count("\n"), "%.0f s" % (time.time() - t0))
raw = json.dumps(out, separators=(",", ":"), sort_keys=True).encode()
with gzip.GzipFile(os.path.join(HERE, "calibrate_full.json.gz"), "wb", mtime=0) as fh:
    fh.write(raw)
print("wrote calibrate_full.json.gz", len(raw))
