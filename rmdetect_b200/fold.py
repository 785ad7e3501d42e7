"""Batched ViennaRNA driver: the base-pair-probability provider in front of the scan path.

Folding stays the reference's external dependency (ViennaRNA `RNAfold -p`, McCaskill partition function);
this module only runs it well.  What it replaces in the reference (paths relative to the reference root):

* lib/rnatools.py:31-39   `RNATools.bpprobs`: one `RNAfold -p` per window, path hard-coded to a developer's
  home directory (SURVEY.md defect D2);
* lib/rnatools.py:131-159 `vienna_tool`: fork + `rna.ps` in the CURRENT directory, removed with `os.system("rm")`
  — not re-entrant, so windows can only be folded one after the other;
* lib/rnatools.py:176-205 `parse_bpprobs`: reads `rna.ps` and waits for `% data start here`; current ViennaRNA
  releases write the `ubox` lines to `dot.ps` (defect D3).

Here every fold runs in a directory of its own, `workers` of them at a time (threads that wait on child
processes), and the `dot.ps` text goes through the library's parser (`rmd_parse_dot_ps`, csrc/ingest.cpp:
`p = sqrt_p ** 2.0` through libm's pow as CPython evaluates it, zeros dropped, a later line for the same pair
wins, sorted upper triangle) straight into the arrays the scan uploads.  In real scans folding is > 99 % of
the wall time, so this is where a user of the drop-in gets the host cores to work.

The provider interface is the one `Scanner.rna_tools` expects: `bpprobs(seq, constraint)` like the reference's
`RNATools`, `coo(seq)` for the coordinate list, and `coo_many(seqs)` which `ScanEngine.build_job` uses to fold
all windows of a job concurrently.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .bpprobs import BPProbs


class FoldError(RuntimeError):
    pass


class RNAfold:
    def __init__(self, exe=None, workers=None, args=("-p",), use_constraint=False, factory=BPProbs, tmp_root=None):
        """exe: path of the RNAfold binary (default: $RMDETECT_RNAFOLD, else `RNAfold` on PATH).
        args: its arguments; the reference passes only `-p` (lib/rnatools.py:34).
        use_constraint: pass `-C` and the constraint line.  The reference builds the input with the constraint but
        never passes `-C` (lib/rnatools.py:134 is commented out, SURVEY.md defect D7), so the line is inert there;
        False reproduces that by not sending it at all."""
        self.exe = exe or os.environ.get("RMDETECT_RNAFOLD") or shutil.which("RNAfold")
        self.workers = max(1, int(workers or os.cpu_count() or 1))
        self.args = list(args)
        self.use_constraint = bool(use_constraint)
        self.factory = factory
        self.tmp_root = tmp_root
        self.folds = 0

    # ---- one fold -----------------------------------------------------------------------------------------
    def _fold_text(self, seq, constraint=""):
        if not self.exe:
            raise FoldError("RNAfold not found: install ViennaRNA or set RMDETECT_RNAFOLD to the binary")
        lines = [seq]
        cmd = [self.exe] + self.args
        if self.use_constraint and constraint:
            cmd.append("-C")
            lines.append(constraint)
        with tempfile.TemporaryDirectory(prefix="rmfold_", dir=self.tmp_root) as work:
            try:
                proc = subprocess.run(cmd, input=("\n".join(lines) + "\n").encode(), cwd=work,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            except OSError as exc:
                raise FoldError("cannot run %s: %s" % (self.exe, exc)) from exc
            if proc.returncode != 0:
                raise FoldError("%s failed (%d): %s" % (self.exe, proc.returncode, proc.stderr.decode(errors="replace")[-300:]))
            path = os.path.join(work, "dot.ps")
            if not os.path.exists(path):             # a FASTA header in the input names the files <id>_dp.ps
                named = [f for f in os.listdir(work) if f.endswith("_dp.ps")]
                if not named:
                    raise FoldError("%s wrote no dot.ps (is `-p` among its arguments?)" % self.exe)
                path = os.path.join(work, named[0])
            with open(path, "rb") as fh:
                return fh.read()

    def coo(self, seq, constraint=""):
        """(i uint16[], j uint16[], p float64[]) of one window, 0-based, i < j, sorted by (i, j)."""
        from ._native import parse_dot_ps
        if len(seq) == 0:
            z = np.zeros(0, dtype=np.uint16)
            return z, z, np.zeros(0, dtype=np.float64)
        wseq, i, j, p, _q = parse_dot_ps(self._fold_text(seq, constraint))
        if wseq is not None and len(wseq) != len(seq):
            raise FoldError("dot.ps describes %d letters, the window has %d" % (len(wseq), len(seq)))
        self.folds += 1
        return i, j, p

    # ---- many folds ---------------------------------------------------------------------------------------
    def coo_many(self, seqs, constraint=""):
        """Coordinate lists of many windows, folded `workers` at a time; equal windows are folded once."""
        uniq = {}
        for s in seqs:
            uniq.setdefault(s, None)
        keys = list(uniq)
        if self.workers == 1 or len(keys) < 2:
            vals = [self.coo(s, constraint) for s in keys]
        else:
            with ThreadPoolExecutor(max_workers=min(self.workers, len(keys))) as pool:
                vals = list(pool.map(lambda s: self.coo(s, constraint), keys))
        for k, v in zip(keys, vals):
            uniq[k] = v
        return [uniq[s] for s in seqs]

    # ---- energies (lib/rnatools.py:41-52 `fold`, :160-174 `parse_fold`, :231-248 `parse_line`) ------------------
    def fold(self, seq, constraint="", ensemble=False):
        """(structure, mfe) or, with `ensemble`, (structure, free energy of the ensemble): `RNAfold --noPS [-p] [-C]`.
        The constraint is applied (`-C`) when the driver was built with use_constraint=True; otherwise it is inert, as
        in the reference as shipped (which sends the line but never `-C`: lib/rnatools.py:134, defect D7)."""
        import re
        if not self.exe:
            raise FoldError("RNAfold not found: install ViennaRNA or set RMDETECT_RNAFOLD to the binary")
        cmd = [self.exe, "--noPS"] + (["-p"] if ensemble else [])
        lines = [seq]
        if constraint and self.use_constraint:          # without -C RNAfold would read the line as another sequence
            lines.append(constraint)
            cmd.append("-C")
        with tempfile.TemporaryDirectory(prefix="rmfold_", dir=self.tmp_root) as work:
            try:
                proc = subprocess.run(cmd, input=("\n".join(lines) + "\n").encode(), cwd=work,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            except OSError as exc:
                raise FoldError("cannot run %s: %s" % (self.exe, exc)) from exc
        if proc.returncode != 0:
            raise FoldError("%s failed (%d): %s" % (self.exe, proc.returncode, proc.stderr.decode(errors="replace")[-300:]))
        out = proc.stdout.decode(errors="replace").split("\n")
        self.folds += 1
        m = re.match(r"([\.\(\)]+)( \(\s*(\-?\d+\.\d+)\))?", out[1].strip()) if len(out) > 1 else None
        if m is None:
            raise FoldError("cannot read RNAfold's structure line: %r" % (out[1] if len(out) > 1 else ""))
        struct, energy = m.group(1), float(m.group(3)) if m.group(3) is not None else 0.0
        if ensemble:
            m = re.match(r"([\.\(\)\{\}\,\|]+)( \[\s*(\-?\d+\.\d+)\])", out[2].strip()) if len(out) > 2 else None
            if m is None:
                raise FoldError("cannot read RNAfold's ensemble line: %r" % (out[2] if len(out) > 2 else ""))
            energy = float(m.group(3))
        return struct, energy

    # ---- the reference's call shape (lib/rnatools.py:31-39) -------------------------------------------------
    def bpprobs(self, seq, constraint=""):
        i, j, p = self.coo(seq, constraint)
        out = self.factory()
        for a, b, v in zip(i.tolist(), j.tolist(), p.tolist()):
            out.add(a, b, v)
        return out
