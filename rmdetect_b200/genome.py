"""Genome-scale scans: one long sequence, optionally both strands, cut into pieces per GPU.

BASELINE.json configs[4]: a 3 Gb sequence, both strands, all models, sharded by chunk + halo over the
GPUs of one box.  What the reference does with such an input (paths relative to the reference root):

* lib/sequences.py:103-111,138-143  --both-strands appends the reverse complement as a second sequence
  (key + "--"), so a strand is just another sequence;
* lib/scanner.py:74-75, lib/gcx.py:4-13  the GC table is taken over the WHOLE sequence, before the first
  window: a strand that is scanned in pieces needs its letter counts first (the one exchange before the
  scan, SURVEY.md §8(e): an all-reduce of four counters);
* lib/scanner.py:80-111  windows at k * step, the last one re-anchored at len - W: a piece that starts on a
  window start and is scanned as a sequence of its own has exactly the parent's windows;
* lib/scanner.py:216-244  the duplicate rule looks at later overlapping windows: every piece carries
  ceil(W / step) halo windows that are scanned only for that decision (dist.halo_windows).

No window list is materialised (a 3 Gb strand has 4 * 10^7 windows): the geometry is arithmetic.  The
sequence and its stand-in bpp matrices are generated on the device (rmd_synth_*; ViennaRNA is external to
the path); hits are reduced to counts and order-independent checksums, or handed to `on_hits`.
"""
from __future__ import annotations

import math
import time

import numpy as np

from .dist import halo_windows
from .gcx import gc_table
from .scanner import placements_in_window

MASK64 = (1 << 64) - 1


def strand_windows(n, win_len, win_step):
    """Number of windows of a sequence of n letters (reference lib/scanner.py:80-111)."""
    if n <= 0:
        return 0
    if win_len == 0 or win_step == 0 or n <= win_len:
        return 1
    return 1 + -(-(n - win_len) // win_step)


def window_start(k, n_win, n, win_len, win_step):
    return max(0, n - win_len) if k == n_win - 1 else k * win_step


def rank_pieces(n, win_len, win_step, rank, world, piece_windows):
    """Pieces of one strand that `rank` scans: (a, b, bh, lo, hi) = own windows [a, b), scanned windows
    [a, bh) (halo for the duplicate rule), letters [lo, hi) of the strand."""
    n_win = strand_windows(n, win_len, win_step)
    halo = halo_windows(win_len, win_step)
    first, last = n_win * rank // world, n_win * (rank + 1) // world
    out = []
    for a in range(first, last, piece_windows):
        b = min(a + piece_windows, last)
        bh = min(n_win, b + halo)
        lo = window_start(a, n_win, n, win_len, win_step)
        hi = n if bh == n_win else window_start(bh - 1, n_win, n, win_len, win_step) + win_len
        out.append((a, b, bh, lo, hi))
    return out


def hit_checksum(strand, start, p0, p1, model, leaf, score):
    """Order-independent checksum of a hit list: wrapping sum of mixed (strand, absolute positions, model,
    gap configuration) keys and of the scores' bit patterns."""
    with np.errstate(over="ignore"):
        key = (start.astype(np.uint64) + p0.astype(np.uint64)) * np.uint64(0x9E3779B97F4A7C15)
        key ^= (start.astype(np.uint64) + p1.astype(np.uint64)) * np.uint64(0xC2B2AE3D27D4EB4F)
        key ^= (model.astype(np.uint64) << np.uint64(40)) ^ (leaf.astype(np.uint64) << np.uint64(48)) ^ np.uint64(strand << 60)
        key = (key ^ (key >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
        total = int(key.sum(dtype=np.uint64)) + int(np.ascontiguousarray(score).view(np.uint64).sum(dtype=np.uint64))
    return total & MASK64


def strand_gc(counts_fwd, reverse):
    """GC dict of a strand from the forward strand's (A, C, G, U) counts (reference lib/gcx.py:4-13)."""
    c = [int(v) for v in counts_fwd]
    if reverse:
        c = c[::-1]                                   # A<->U, C<->G
    n = sum(c)
    return {b: math.log(float(v) / n) for b, v in zip("ACGU", c) if v}


def scan_synthetic_genome(engine, total_nt, both_strands, seed, win_len, win_step, min_bpp_mean, cutoff,
                          rank=0, world=1, dist=None, piece_windows=266_667, on_hits=None, min_score=None,
                          forward_counts=None, table=None):
    """Scan `rank`'s share of a synthetic genome of `total_nt` letters per strand.

    Returns a dict with whole-job totals (after the closing all-reduce when `dist` is given): tests_all,
    tests_pairs, hits, checksum, plus this rank's timings.  `on_hits(strand, lo, hits)` receives each piece's
    own hits (window starts relative to `lo`) when the caller wants the records themselves.  `min_score` moves
    the command line's score filter (reference rmdetect.py:44, default 5.0 there) in front of the download.
    `forward_counts` (A, C, G, U of the whole forward strand) replaces the counting pass and its all-reduce.
    `table` (a `hittable.HitTable`) collects every piece's own hits on the device, in strand coordinates, under the keys
    "genome" and "genome--" (lib/sequences.py:138-143): the list rmfilter / rmcluster would work on never leaves the GPU.
    """
    import torch
    ctx = engine.ctx
    n_models = len(engine.models)
    # ---- the exchange before: letter counts of the whole forward strand ------------------------------------
    if dist is not None and world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    if forward_counts is None:
        lo, hi = total_nt * rank // world, total_nt * (rank + 1) // world
        ctx.synth_strand(total_nt, False)
        counts = ctx.synth_counts(hi - lo, lo, seed)
        if dist is not None and world > 1:
            t = torch.from_numpy(counts.copy()).to(dev)
            dist.all_reduce(t)                         # 4 x int64
            counts = t.cpu().numpy()
    else:
        counts = np.asarray(forward_counts, dtype=np.int64)
    if int(counts.sum()) != total_nt:
        raise ValueError("letter counts %s do not add up to the strand's %d letters (ranks of a multi-GPU scan "
                         "need the all-reduce: pass `dist`)" % (counts.tolist(), total_nt))
    ctx.set_min_score(min_score)

    wlen = min(win_len, total_nt) if win_len else total_nt
    per_window = sum(placements_in_window(m, wlen) for m in engine.models)
    tests_all = tests_pairs = n_hits = n_raw = checksum = 0
    ms_gen = ms_scan = ms_dev = 0.0
    n_windows = n_pieces = n_entries = 0
    t0 = time.perf_counter()
    for strand in range(2 if both_strands else 1):
        gc = strand_gc(counts, strand == 1)
        cls = []
        for ev in engine.evalues:
            if ev is None or not ev.gc_classes:
                cls.append(-1)
            else:
                ev.select_gc(gc)
                cls.append(-1 if ev.selected is None else ev.selected)
        tab = gc_table(gc)
        if table is not None:
            table.key_id("genome--" if strand else "genome")
        ctx.synth_strand(total_nt, strand == 1)
        for a, b, bh, plo, phi in rank_pieces(total_nt, win_len, win_step, rank, world, piece_windows):
            n_win, n_bpp, ms = ctx.synth_job(phi - plo, plo, seed, seed, win_len, win_step, min_bpp_mean, cutoff, tab, cls)
            assert n_win == bh - a, (n_win, a, bh)
            r = ctx.scan_resident(0, n_win, copy=False)
            own = b - a
            h = r["hits"]
            n_own = int(np.searchsorted(h["window"], own))   # hits are in window order
            h = h[:n_own]                             # halo windows only served the duplicate rule
            if table is not None and n_own:
                table.append_scan(["genome--" if strand else "genome"], start_base=plo, first_n=n_own)
            n_hits += len(h)
            n_raw += int(np.asarray(r["raw_count"])[:own].sum())
            tests_pairs += int(np.asarray(r["gate_pass"])[:own].sum())
            tests_all += own * per_window
            if len(h):
                checksum = (checksum + hit_checksum(strand, h["start"] + plo, h["p0"], h["p1"], h["model"], h["leaf"],
                                                    h["score"])) & MASK64
                if on_hits is not None:
                    on_hits(strand, plo, h)
            ms_gen += ms; ms_scan += r["ms_scan"]; ms_dev += r["ms_total"]
            n_windows += own; n_pieces += 1; n_entries += n_bpp
    wall = time.perf_counter() - t0
    ctx.set_min_score(None)
    ctx.synth_strand(0, False)
    tot = np.array([tests_all, tests_pairs, n_hits, n_raw, n_windows, n_entries], dtype=np.int64)
    sums = [checksum >> 32, checksum & 0xffffffff]
    if dist is not None and world > 1:                # the exchange after: counters (hit records stay with their rank)
        t = torch.from_numpy(np.concatenate([tot, np.array(sums, dtype=np.int64)])).to(dev)
        dist.all_reduce(t)
        v = t.cpu().numpy()
        tot, sums = v[:6], [int(v[6]), int(v[7])]
    checksum = ((sums[0] << 32) + sums[1]) & MASK64
    return dict(tests_all=int(tot[0]), tests_pairs=int(tot[1]), hits=int(tot[2]), hits_before_dedupe=int(tot[3]),
                windows=int(tot[4]), bpp_entries=int(tot[5]), checksum=checksum, counts=[int(c) for c in counts],
                rank_wall_s=wall, rank_ms_generate=ms_gen, rank_ms_scan_kernel=ms_scan, rank_ms_device=ms_dev,
                rank_pieces=n_pieces, rank_windows=n_windows)


# ---- real input: FASTA chromosomes + a fold provider -------------------------------------------------------------
def read_fasta(path):
    """[(key, sequence)] of a FASTA file with the reference's rules (lib/file_io.py:5-24: key = first word after '>',
    lines concatenated; lib/sequences.py:145-150: upper case, T -> U).  Vectorised: a chromosome is hundreds of megabytes."""
    raw = np.fromfile(path, dtype=np.uint8)
    nl = np.flatnonzero(raw == 10)
    line_starts = np.concatenate([[0], nl + 1])
    line_starts = line_starts[line_starts < len(raw)]
    heads = line_starts[raw[line_starts] == ord(">")]
    # (a '>' preceded by blanks on its line is not recognised as a header; the reference strips every line first)
    out = []
    lut = np.arange(256, dtype=np.uint8)
    for c in range(ord("a"), ord("z") + 1):
        lut[c] = c - 32
    lut[ord("T")] = lut[ord("t")] = ord("U")
    for k, h in enumerate(heads):
        end_head = nl[np.searchsorted(nl, h)] if np.searchsorted(nl, h) < len(nl) else len(raw)
        words = raw[h + 1:end_head].tobytes().decode("latin-1").strip().split()
        key = words[0] if words else ""
        body = raw[end_head + 1:heads[k + 1] if k + 1 < len(heads) else len(raw)]
        body = body[(body != 10) & (body != 13) & (body != 32) & (body != 9)]
        out.append((key, lut[body].tobytes().decode("latin-1")))
    return out


def reverse_complement(seq):
    """lib/sequences.py:138-143: reversed, A<->U and C<->G, every other letter unchanged."""
    return seq[::-1].translate(str.maketrans("ACGU", "UGCA"))


def scan_genome(engine, sequences, fold, both_strands=False, win_len=150, win_step=75, min_bpp_mean=0.001,
                cutoff=math.log(0.1), constraint="", rank=0, world=1, dist=None, piece_windows=32768, min_score=None,
                table=None, on_hits=None):
    """Scan `rank`'s share of real sequences: [(key, sequence)] (e.g. `read_fasta`), bpp lists from `fold` (an object with
    `coo_many(windows, constraint)` / `coo` / `bpprobs`: fold.RNAfold, the reference's RNATools, synth.SynthFold).

    What `rmdetect.py` does with such an input (lib/sequences.py:103-111: with --both-strands every sequence is followed,
    after all forward ones, by its reverse complement under key + "--"; lib/scanner.py:73-113: GC table of the whole
    sequence, windows at k * step) is done piece by piece: every sequence is cut at window starts into contiguous ranges of
    windows, one share per rank (`rank_pieces`), each piece is scanned as a sequence of its own with the parent's GC table
    plus halo windows for the duplicate rule, and its own hits are kept in parent coordinates.  `table` (a `HitTable`)
    collects them on the device; `on_hits(key, hits)` receives the records (start in parent coordinates).  Returns totals
    over all ranks (after the closing all-reduce when `dist` is given)."""
    from .gcx import GC
    from ._native import HIT_DTYPE
    strands = [(k, s) for k, s in sequences]
    if both_strands:
        strands += [(k + "--", reverse_complement(s)) for k, s in sequences]
    engine.ctx.set_min_score(min_score)
    tests_all = tests_pairs = n_hits = n_windows = checksum = 0
    ms_dev = 0.0
    t0 = time.perf_counter()
    for si, (key, seq) in enumerate(strands):
        if len(seq) == 0:
            continue
        gc = GC.compute(seq)                                   # over the whole sequence, before the first window
        if table is not None:
            table.key_id(key)                                  # ids in strand order on every rank
        for a, b, bh, lo, hi in rank_pieces(len(seq), win_len, win_step, rank, world, piece_windows):
            job = engine.build_job([(key, seq[lo:hi], None)], win_len, win_step, fold, constraint, min_bpp_mean, cutoff, gc=gc)
            assert job["n_win"] == bh - a, (job["n_win"], a, bh)
            r = engine.ctx.scan(job, copy=False)
            own = b - a
            h = r["hits"]
            n_own = int(np.searchsorted(h["window"], own))     # hits are in window order; halo windows come last
            h = h[:n_own]
            tests_pairs += int(np.asarray(r["gate_pass"])[:own].sum())
            starts = job["starts"][0][:own]
            tests_all += sum(placements_in_window(m, min(win_len, hi - lo - int(s)) if win_len else hi - lo)
                             for s in starts.tolist() for m in engine.models)
            n_hits += n_own
            n_windows += own
            ms_dev += r["ms_total"]
            if n_own:
                checksum = (checksum + hit_checksum(si & 7, h["start"] + lo, h["p0"], h["p1"], h["model"], h["leaf"], h["score"])) & MASK64
                if table is not None:
                    table.append_scan([key], start_base=lo, first_n=n_own)
                if on_hits is not None:
                    hh = np.array(h, dtype=HIT_DTYPE, copy=True)
                    hh["start"] += lo
                    hh["window"] += a
                    hh["seq"] = si
                    on_hits(key, hh)
    wall = time.perf_counter() - t0
    engine.ctx.set_min_score(None)
    tot = np.array([tests_all, tests_pairs, n_hits, n_windows, checksum >> 32, checksum & 0xffffffff], dtype=np.int64)
    if dist is not None and world > 1:
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(tot).to(dev)
        dist.all_reduce(t)                                     # counters only: hit records stay with their rank
        tot = t.cpu().numpy()
    return dict(tests_all=int(tot[0]), tests_pairs=int(tot[1]), hits=int(tot[2]), windows=int(tot[3]),
                checksum=((int(tot[4]) << 32) + int(tot[5])) & MASK64, strands=len(strands), rank_wall_s=wall, rank_ms_device=ms_dev)
