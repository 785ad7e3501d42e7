"""E-value calibration (the reference's rmcalibrate.py) with sample scoring on the GPU.

Host side keeps the reference's bookkeeping exactly: the sample stream
(`draw_seq`, rmcalibrate.py:15-16), the table construction (`get_gc_evalues`,
:64-99), the convergence measure (`diff_gc_evalues`, :134-157), the 10 000
sample checkpoints and stop rule (:237-257) and the `.evalues` text layout
(`save_data`, :101-131).  The per-sample work — evaluating the model on the
spliced sequence and adding 165 weighted scores to histograms (:265-295) — is
`rmd_calibrate` / `rmd_calibrate_synth` (csrc/calib_kernels.cuh).
"""
from __future__ import annotations

import math
import random
import sys
import time

import numpy as np

from . import _native
from .gcx import GC
from .modelflat import flatten_model

N_SAVE = 10000            # rmcalibrate.py:163
N_BINS = 512              # score buckets 0.0 .. 51.1
CONVERGED_DIFF = 0.001    # rmcalibrate.py:249
CONVERGED_SAMPLES = 100000


def draw_samples(n, L, seed=None):
    """`n` samples of `L` letters as codes, consuming Python's RNG exactly like `draw_seq`."""
    if seed is not None:
        random.seed(seed)
    ri = random.randint
    return np.array([[ri(0, 3) for _ in range(L)] for _ in range(n)], dtype=np.uint8).reshape(n, L)


def class_logs(classes):
    """log(gc_i[letter]) in ACGU order, taken with math.log like rmcalibrate.py:43,58."""
    return np.array([[math.log(c[ch]) for ch in "ACGU"] for c in classes], dtype=np.float64)


def hist_dicts(hist):
    """Dense [class][bucket] array -> the reference's dicts keyed by round(score, 1)."""
    out = []
    for row in hist:
        nz = np.flatnonzero(row)
        out.append({round(int(k) / 10.0, 1): float(row[k]) for k in nz})
    return out


def gc_evalues(gc_hists, evalue_min):
    """rmcalibrate.py:64-99: reverse cumulative probabilities per class, floored at `evalue_min`."""
    score_max, score_min, score_step = -100.0, 100.0, 0.1
    for hist in gc_hists:
        if hist:
            score_max = max(max(hist.keys()), score_max)
            score_min = min(min(hist.keys()), score_min)
    score, scores = score_min, []
    while score <= score_max:
        scores.append(round(score, 1))
        score += score_step
    gc_count = [float(sum(h.values())) for h in gc_hists]
    gc_evalue = []
    for i, hist in enumerate(gc_hists):
        tab, last = {}, ""
        for s in scores[::-1]:
            p = hist.get(s, 0.0) / gc_count[i]
            tab[s] = max(tab.get(last, 0.0) + p, evalue_min)
            last = s
        gc_evalue.append(tab)
    return gc_evalue, scores


def diff_gc_evalues(gce1, gce2, scores):
    """rmcalibrate.py:134-157: relative area difference of the middle class between two tables."""
    gc = len(gce1) // 2
    diff = atotal = 0.0
    for i in range(1, len(scores)):
        a, b = scores[i - 1], scores[i]
        ya1, yb1 = gce1[gc].get(a), gce1[gc].get(b)
        ya2, yb2 = gce2[gc].get(a), gce2[gc].get(b)
        if None not in (ya1, yb1, ya2, yb2):
            x = b - a
            a1 = x * (yb1 + abs(yb1 - ya1) / 2.0)
            a2 = x * (yb2 + abs(yb2 - ya2) / 2.0)
            diff += abs(a1 - a2)
            atotal += a1
    return diff / atotal


def table_text(gc_classes, gc_evalue, scores):
    """The `.evalues` file body (rmcalibrate.py:101-131)."""
    lines = ["[GC_CLASSES]\n"]
    header = "#SCORES"
    for i, gc in enumerate(gc_classes):
        name = "GC_%04d" % i
        header += "\t%10s" % name
        lines.append(name + "".join("\t%s\t%.4f" % (k, v) for k, v in gc.items()) + "\n")
    lines.append("[E_VALUES]\n%s\n" % header)
    for s in scores:
        lines.append("%s" % s + "".join("\t%10.3E" % gc_evalue[i].get(s, 0.0) for i in range(len(gc_classes))) + "\n")
    return "".join(lines)


class Calibrator:
    """One model resident on one GPU, scoring calibration samples."""

    def __init__(self, model, unknown_prob, cutoff, device=0, ctx=None, model_index=0):
        self.model = model
        self.L = len(model.nodes)
        self.cutoff = cutoff
        self.classes = GC.get_classes(resolution=5, min_p=15, max_p=85)     # rmcalibrate.py:224
        self.class_log = class_logs(self.classes)
        if ctx is None:
            ctx = _native.Context(device)
            ctx.set_models([flatten_model(model, unknown_prob)])
            model_index = 0
        self.ctx, self.model_index = ctx, model_index
        self.hist = np.zeros((len(self.classes), N_BINS), dtype=np.float64)
        self.ms_kernel = 0.0

    def add_samples(self, samples):
        n_ok, ms = self.ctx.calibrate(self.model_index, samples, self.class_log, self.cutoff, self.hist)
        self.ms_kernel += ms
        return n_ok

    def add_synthetic(self, seed, first, n):
        n_ok, ms = self.ctx.calibrate(self.model_index, None, self.class_log, self.cutoff, self.hist,
                                      seed=seed, first=first, n=n, L=self.L)
        self.ms_kernel += ms
        return n_ok

    def table(self):
        return gc_evalues(hist_dicts(self.hist), 1.0 / float(4 ** self.L))


def calibrate_model(model, out_file, samples=1000000, unknown_prob=math.log(0.0001), cutoff=math.log(0.1),
                    seed=None, device=0, device_samples=False, log=sys.stderr):
    """The main loop of rmcalibrate.py:235-298 for one model; returns (gc_evalue, scores, n_done)."""
    cal = Calibrator(model, unknown_prob, cutoff, device)
    total = min(samples, 4 ** cal.L)
    if seed is not None:
        random.seed(seed)
    last, stable, done = None, 0, 0
    t0 = time.time()
    while done < total:
        # every N_SAVE samples the reference recomputes the table, saves it and tests convergence
        if done > 0:
            gce, scores = cal.table()
            if out_file:
                with open(out_file, "w") as fh:
                    fh.write(table_text(cal.classes, gce, scores))
            diff = 0.0
            if last is not None:
                diff = diff_gc_evalues(last, gce, scores)
                if diff < CONVERGED_DIFF:
                    stable += N_SAVE
                    if stable > CONVERGED_SAMPLES:
                        print("\nConverged after %d simulations" % done)
                        break
                else:
                    stable = 0
            if log is not None:
                el = time.time() - t0
                log.write("%sdata saved after %d simulations, TTF: %ds, diff: %5.4f (%d)"
                          % ("\b" * 100, done, (el * total / done) - el, diff, stable))
                log.flush()
            last = gce
        step = min(N_SAVE, total - done)
        if device_samples:
            cal.add_synthetic(seed or 0, done, step)
        else:
            cal.add_samples(draw_samples(step, cal.L))
        done += step
    gce, scores = cal.table()
    if out_file:
        with open(out_file, "w") as fh:
            fh.write(table_text(cal.classes, gce, scores))
    cal.ctx.close()
    return gce, scores, done
