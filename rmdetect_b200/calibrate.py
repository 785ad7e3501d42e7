"""E-value calibration (the reference's rmcalibrate.py) with sample scoring on the GPU.

Host side keeps the reference's bookkeeping exactly: the sample stream
(`draw_seq`, rmcalibrate.py:15-16), the table construction (`get_gc_evalues`,
:64-99), the convergence measure (`diff_gc_evalues`, :134-157), the 10 000
sample checkpoints and stop rule (:237-257) and the `.evalues` text layout
(`save_data`, :101-131).  The per-sample work — evaluating the model on the
spliced sequence and adding 165 weighted scores to histograms (:265-295) — is
`rmd_calibrate` / `rmd_calibrate_synth` (csrc/calib_kernels.cuh).

Several GPUs (BASELINE.json configs[3]): samples are independent, so the ranks of a
`torch.distributed` group score disjoint slices of every checkpoint block and all-reduce (sum) the block's
165 x N_BINS fp64 histogram; the reference's table construction and stop rule then run on the reduced
histogram, identically on every rank.  A checkpoint block is N_SAVE samples per rank.
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


def draw_samples_loop(n, L, seed=None):
    """`n` samples of `L` letters as codes, letter by letter through `random.randint(0, 3)` as `draw_seq` does
    (rmcalibrate.py:16-22): the definition that `draw_samples` reproduces."""
    if seed is not None:
        random.seed(seed)
    ri = random.randint
    return np.array([[ri(0, 3) for _ in range(L)] for _ in range(n)], dtype=np.uint8).reshape(n, L)


def draw_samples(n, L, seed=None):
    """`n` samples of `L` letters as codes, consuming Python's RNG exactly like `draw_seq` — without 16 million calls of
    `random.randint` for a million samples (they took longer than everything else in a calibration run).

    `randint(0, 3)` is `_randbelow(4)`: the top three bits of one 32-bit output of the Mersenne Twister, drawn again while
    they are 4 or more.  The generator's state is handed to numpy's MT19937 (same algorithm, same state layout), the raw
    outputs are taken in bulk, the rejected ones dropped, and the state after exactly the outputs consumed is handed back to
    `random`, so that whatever draws next — the reference's own code included — continues the same stream."""
    if seed is not None:
        random.seed(seed)
    need = int(n) * int(L)
    if need == 0:
        return np.zeros((n, L), dtype=np.uint8)
    version, internal, gauss = random.getstate()
    if version != 3 or len(internal) != 625:                      # an unknown generator layout: the plain loop
        return draw_samples_loop(n, L)
    start = {"bit_generator": "MT19937", "state": {"key": np.array(internal[:-1], dtype=np.uint32), "pos": int(internal[-1])}}
    mt = np.random.MT19937()
    mt.state = start
    parts, have = [], 0
    while have < need:
        # half of the outputs are rejected: a first piece just short of what is needed, small pieces for the rest, so that
        # little has to be drawn again to leave the generator exactly behind the last output used
        before = mt.state
        raw = mt.random_raw(int(1.9 * need) if have == 0 and need > 4096 else max(256, int(2.2 * (need - have)) + 32))
        v = (raw >> np.uint64(29)).astype(np.uint8)
        ok = np.flatnonzero(v < 4)
        take = min(len(ok), need - have)
        parts.append(v[ok[:take]])
        have += take
        if have == need and int(ok[take - 1]) + 1 < len(raw):
            mt.state = before
            mt.random_raw(int(ok[take - 1]) + 1)
    end = mt.state["state"]
    random.setstate((version, tuple(int(x) for x in end["key"]) + (int(end["pos"]),), gauss))
    return np.concatenate(parts).reshape(n, L)


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


class _Scores(list):
    """A score grid as the list the reference builds, with what the dense functions derive from it again and again:
    `idx[k]` = histogram bucket of row k, `at[score]` = row."""
    __slots__ = ("idx", "at")


_GRIDS = {}


def _grid(first, last):
    """The grid from bucket `first` to bucket `last` (rmcalibrate.py:65-83: repeated `+= 0.1` steps, each rounded for use as
    a key — the float accumulation, and with it a possibly missing last row, is the reference's); cached: a calibration run
    asks for the same few grids at every checkpoint."""
    g = _GRIDS.get((first, last))
    if g is None:
        score, score_max, scores = round(first / 10.0, 1), round(last / 10.0, 1), []
        while score <= score_max:
            scores.append(round(score, 1))
            score += 0.1
        g = (tuple(scores), np.array([int(round(v * 10)) for v in scores], dtype=np.int64), {v: k for k, v in enumerate(scores)})
        if len(_GRIDS) > 4096:
            _GRIDS.clear()
        _GRIDS[(first, last)] = g
    out = _Scores(g[0])
    out.idx, out.at = g[1], g[2]
    return out


def score_grid(hist, cols=None):
    """The score rows `get_gc_evalues` walks (rmcalibrate.py:65-83) for a dense [class][bucket] histogram: from the
    smallest to the largest occupied bucket; `cols` = (first, last) occupied bucket when the caller knows them."""
    if cols is None:
        occupied = np.flatnonzero(hist.any(axis=0))
        cols = (int(occupied[0]), int(occupied[-1])) if len(occupied) else (-1, -1)
    if cols[1] < 0:
        out = _Scores()
        out.idx, out.at = np.zeros(0, dtype=np.int64), {}
        return out
    return _grid(int(cols[0]), int(cols[1]))


def gc_evalues_dense(hist, evalue_min, classes=None, cols=None):
    """`gc_evalues` on the dense histogram: (table [class][row], scores); `classes` restricts the table to those rows
    of the histogram (the score grid still spans all of them); `cols` = (first, last) occupied bucket of the WHOLE
    histogram when `hist` holds only some of its rows.  Same additions in the same order (a
    sequential cumulative sum from the highest score down, floored at `evalue_min`; after the first row the floor
    never binds because every term is >= 0), so the values equal the dict version's bit for bit."""
    scores = score_grid(hist, cols)
    idx = scores.idx
    sub = hist if classes is None else hist[list(classes)]
    # Python's own sum(), like :85 — not a numpy running sum: since Python 3.12 the built-in compensates its float sums
    # (Neumaier), so only the built-in gives the number the reference's `sum(hist.values())` gives under the same interpreter
    count = np.array([float(sum(row[np.flatnonzero(row)].tolist())) for row in sub])
    p = sub[:, idx[::-1]] / count[:, None]
    p[:, 0] = np.maximum(p[:, 0], evalue_min)
    tab = np.maximum(np.cumsum(p, axis=1), evalue_min)[:, ::-1]
    return tab, scores


def diff_dense(tab1, scores1, tab2, scores2, gc=None):
    """`diff_gc_evalues` (rmcalibrate.py:134-157) for two dense tables: the middle class (row `gc` of the tables,
    default their middle row), rows of the NEWER grid `scores2`, a row pair skipped when the older table lacks one of
    the scores.  The terms are formed row by row with the reference's operations (numpy, element by element) and summed
    in its order (a running sum, `cumsum`), so the value is the loop's bit for bit."""
    gc = tab1.shape[0] // 2 if gc is None else gc
    at1 = getattr(scores1, "at", None)
    if at1 is None:
        at1 = {s: k for k, s in enumerate(scores1)}
    diff = atotal = 0.0
    if len(scores2) > 1:
        k1 = np.array([at1.get(v, -1) for v in scores2], dtype=np.int64)
        use = np.flatnonzero((k1[:-1] >= 0) & (k1[1:] >= 0))
        if len(use):
            s2 = np.asarray(scores2, dtype=np.float64)
            r1, r2 = np.asarray(tab1[gc], dtype=np.float64), np.asarray(tab2[gc], dtype=np.float64)
            x = s2[use + 1] - s2[use]
            ya1, yb1, ya2, yb2 = r1[k1[use]], r1[k1[use + 1]], r2[use], r2[use + 1]
            a1 = x * (yb1 + np.abs(yb1 - ya1) / 2.0)
            a2 = x * (yb2 + np.abs(yb2 - ya2) / 2.0)
            diff, atotal = float(np.cumsum(np.abs(a1 - a2))[-1]), float(np.cumsum(a1)[-1])
    return diff / atotal


def table_text_dense(gc_classes, tab, scores):
    """`table_text` for a dense table [class][row] — the same bytes, one format call per score row instead of one per
    number (a checkpoint writes 165 x ~200 numbers; call by call that was two thirds of a checkpoint's time)."""
    lines = ["[GC_CLASSES]\n"]
    header = "#SCORES"
    for i, gc in enumerate(gc_classes):
        name = "GC_%04d" % i
        header += "\t%10s" % name
        lines.append(name + "".join("\t%s\t%.4f" % (k, v) for k, v in gc.items()) + "\n")
    lines.append("[E_VALUES]\n%s\n" % header)
    row_format = "%s" + "\t%10.3E" * len(gc_classes) + "\n"
    cols = np.asarray(tab, dtype=np.float64).T.tolist()              # [row][class]
    for s, values in zip(scores, cols):
        lines.append(row_format % (s, *values))
    return "".join(lines)


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


class _DeviceArray:
    """A device allocation owned by the library, presented through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=2)


class Calibrator:
    """One model resident on one GPU, scoring calibration samples into histograms that stay on the device
    (a calibration session of the C ABI: rmd_calib_open / _add / _device_hist / _read)."""

    def __init__(self, model, unknown_prob, cutoff, device=0, ctx=None, model_index=0):
        self.model = model
        self.L = len(model.nodes)
        self.cutoff = cutoff
        self.classes = GC.get_classes(resolution=5, min_p=15, max_p=85)     # rmcalibrate.py:224
        self.class_log = class_logs(self.classes)
        self._own = ctx is None
        if ctx is None:
            ctx = _native.Context(device)
            ctx.set_models([flatten_model(model, unknown_prob)])
            model_index = 0
        self.ctx, self.model_index = ctx, model_index
        self.hist = np.zeros((len(self.classes), N_BINS), dtype=np.float64)
        self.ms_kernel = 0.0
        self.n_ok = 0
        self._pinned = None
        self.ctx.calib_open(self.model_index, self.L, self.class_log, self.cutoff, N_BINS)

    def add_samples(self, samples):
        self.ctx.calib_add(samples=samples)

    def add_synthetic(self, seed, first, n):
        self.ctx.calib_add(None, seed, first, n)

    def reduced_hist(self, dist=None):
        """The histogram of everything added so far, summed over the ranks of `dist`.  With NCCL the all-reduce runs
        on the device, on a copy of the session's histogram (the session keeps accumulating this rank's share)."""
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl":
            import torch
            ptr, n = self.ctx.calib_device_hist()                  # waits for the enqueued batches
            t = torch.as_tensor(_DeviceArray(ptr, n), device=torch.device("cuda", self.ctx.device)).clone()
            dist.all_reduce(t)
            if self._pinned is None:                               # one pinned landing buffer for every checkpoint
                self._pinned = torch.empty(n, dtype=torch.float64, pin_memory=True)
            self._pinned.copy_(t)                                  # synchronous for the host
            self.hist = self._pinned.numpy().reshape(self.hist.shape)
        else:
            # self.hist is the landing buffer of every checkpoint (the table built from the previous one is a copy)
            local, self.n_ok, ms = self.ctx.calib_read(out=self.hist)
            self.ms_kernel += ms
            self.hist = all_reduce_hist(local, dist)
        return self.hist

    def reduced_row(self, cls, dist=None):
        """(row [1][N_BINS] of class `cls`, (first, last) occupied bucket over all classes) of the histogram of everything
        added so far, summed over the ranks of `dist`: what the stop rule reads (rmcalibrate.py:241-257) and what spans the
        table's score rows (:65-83) — 4 KB per checkpoint instead of the whole table.  Over several ranks the bucket range
        rides in the same all-reduce as the row: a 0/1 vector that is 1 from a rank's first to its last occupied bucket,
        whose sum is non-zero exactly from the smallest first to the largest last."""
        world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        if world == 1:
            rows, _, cols = self.ctx.calib_read_rows(cls, 1)
            return rows, cols
        if dist.get_backend() == "nccl":
            import torch
            _, ptr, cols = self.ctx.calib_read_rows(cls, 1, want_rows=False)     # waits for the enqueued batches
            t = torch.zeros(2 * N_BINS, dtype=torch.float64, device=torch.device("cuda", self.ctx.device))
            t[:N_BINS] = torch.as_tensor(_DeviceArray(ptr, N_BINS), device=t.device)
            if cols[1] >= 0:
                t[N_BINS + cols[0]:N_BINS + cols[1] + 1] = 1.0
            dist.all_reduce(t)
            both = t.cpu().numpy()
        else:
            rows, _, cols = self.ctx.calib_read_rows(cls, 1)
            both = np.zeros(2 * N_BINS)
            both[:N_BINS] = rows[0]
            if cols[1] >= 0:
                both[N_BINS + cols[0]:N_BINS + cols[1] + 1] = 1.0
            both = all_reduce_hist(both, dist)
        return both[:N_BINS].reshape(1, N_BINS), span_of(both[N_BINS:])

    def table(self):
        return gc_evalues(hist_dicts(self.reduced_hist()), 1.0 / float(4 ** self.L))

    def close(self):
        _, self.n_ok, ms = self.ctx.calib_read(want_hist=False)    # this rank's counters and device time
        self.ms_kernel += ms
        if self._own:
            self.ctx.close()


def all_reduce_hist(block, dist):
    """Sum a [class][bucket] fp64 histogram over the ranks (gloo in the CPU tests; a host-side fallback for NCCL):
    the one collective of the calibration path, 165 x N_BINS doubles per checkpoint block."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return block
    import torch
    nccl = dist.get_backend() == "nccl"
    t = torch.from_numpy(np.array(block, dtype=np.float64, copy=True))
    if nccl:
        t = t.cuda()
    dist.all_reduce(t)
    return t.cpu().numpy() if nccl else t.numpy()


def span_of(marks):
    """(first, last) index of the non-zero entries; (-1, -1) when there is none."""
    nz = np.flatnonzero(marks)
    return (int(nz[0]), int(nz[-1])) if len(nz) else (-1, -1)


def all_reduce_cols(cols, dist):
    """(first, last) occupied bucket over the ranks: the union of their occupied buckets; -1 marks an empty histogram."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return cols
    import torch
    first, last = cols
    t = torch.tensor([-first if first >= 0 else -(1 << 30), last], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    first, last = int(-t[0]), int(t[1])
    return (first, last) if last >= 0 else (-1, -1)


def calibrate_model(model, out_file, samples=1000000, unknown_prob=math.log(0.0001), cutoff=math.log(0.1),
                    seed=None, device=0, device_samples=False, log=sys.stderr, dist=None, scorer=None,
                    save_every_block=True, stats=None):
    """The main loop of rmcalibrate.py:235-298 for one model; returns (table [class][row], scores, samples done).

    One rank: exactly the reference's schedule — a checkpoint (table, file, convergence test) every N_SAVE samples.
    `dist` (an initialised torch.distributed): a checkpoint block is N_SAVE samples PER RANK; rank r scores the r-th
    slice of the block, the histograms are all-reduced, and every rank runs the stop rule on the same reduced
    table (rank 0 writes the file).  Host-drawn samples (`device_samples=False`, the reference's `random.randint`
    stream, seeded alike on every rank) are drawn in full by every rank so that the union of the slices is the
    single-rank stream; `device_samples=True` draws on the GPU, sample index = position in the global stream.
    `scorer` replaces the GPU `Calibrator` (tests: a CPU stand-in with add_samples / add_synthetic / reduced_hist).
    `save_every_block=False` (throughput runs): a checkpoint tabulates only the class the stop rule reads (the middle
    one) and the file is written once, at the end.  `stats` (a dict) receives blocks, kernel ms and the final
    histogram."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    cal = scorer if scorer is not None else Calibrator(model, unknown_prob, cutoff, device)
    total = min(samples, 4 ** cal.L)
    if seed is not None:
        random.seed(seed)
    evalue_min = 1.0 / float(4 ** cal.L)
    mid = len(cal.classes) // 2
    block = N_SAVE * world
    last, stable, done, blocks = None, 0, 0, 0
    hist = None
    t0 = time.time()

    def save(tab, scores):
        if out_file and rank == 0:
            with open(out_file, "w") as fh:
                fh.write(table_text_dense(cal.classes, tab, scores))

    while done < total:
        # every block the reference recomputes the table, saves it and tests convergence (:237-263)
        if done > 0:
            if save_every_block:
                hist = cal.reduced_hist(dist)
                tab, scores = gc_evalues_dense(hist, evalue_min)
                save(tab, scores)
                tab_mid = tab[mid:mid + 1]
            elif hasattr(cal, "reduced_row"):         # only the class the stop rule reads crosses the bus
                row, cols = cal.reduced_row(mid, dist)
                tab_mid, scores = gc_evalues_dense(row, evalue_min, cols=cols)
            else:
                tab_mid, scores = gc_evalues_dense(cal.reduced_hist(dist), evalue_min, classes=[mid])
            diff = 0.0
            if last is not None:
                diff = diff_dense(last[0], last[1], tab_mid, scores, gc=0)
                if diff < CONVERGED_DIFF:
                    stable += block
                    if stable > CONVERGED_SAMPLES:
                        if rank == 0 and log is not None:
                            print("\nConverged after %d simulations" % done)
                        break
                else:
                    stable = 0
            if log is not None and rank == 0:
                el = time.time() - t0
                log.write("%sdata saved after %d simulations, TTF: %ds, diff: %5.4f (%d)"
                          % ("\b" * 100, done, (el * samples / done) - el, diff, stable))
                log.flush()
            last = (tab_mid, scores)
        step = min(block, total - done)
        lo, hi = step * rank // world, step * (rank + 1) // world
        if device_samples:
            if hi > lo:
                cal.add_synthetic(seed or 0, done + lo, hi - lo)
        else:
            drawn = draw_samples(step, cal.L)     # every rank consumes the whole block of the stream
            if hi > lo:
                cal.add_samples(drawn[lo:hi])
        done += step
        blocks += 1
    hist = cal.reduced_hist(dist)
    tab, scores = gc_evalues_dense(hist, evalue_min)
    save(tab, scores)
    if scorer is None or isinstance(cal, Calibrator):
        cal.close()
    if stats is not None:
        stats.update(blocks=blocks, hist=hist, ms_kernel=getattr(cal, "ms_kernel", 0.0), n_ok=getattr(cal, "n_ok", None))
    return tab, scores, done
