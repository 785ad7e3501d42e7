"""The flow around the scan that `rmdetect.py` wires up: callbacks, incremental saves, the final `.res` file.

`rmdetect.py` itself stays the reference's (its command line is part of the drop-in contract); this module restates
what that script does between `Scanner(config)` and the bytes on disk, so that the swapped-in `Scanner` can be
driven — and tested — exactly the way the script drives the reference's (paths relative to the reference root):

* rmdetect.py:237-249   builds the scanner, installs the six callbacks, calls `scan(seqs)`;
* rmdetect.py:86-93     `cback_before_scan`: a new result file with the header and no hits (mode "n");
* rmdetect.py:114-126   `cback_after_window`: more than 1000 new hits since the last save are appended (mode "a");
* rmdetect.py:129-138   `cback_after_sequence`: the sequence's remaining hits are appended;
* rmdetect.py:141-144   `cback_after_scan`: the whole file is rewritten from the complete list (mode "w");
* rmdetect.py:31-70     `cands_save`: `score > min_score` -> joint bpp -> `bpp > min_bpp` -> sort -> `Results.write`.

The joint bpp (two more folds per hit, lib/candidates.py:209-224,396-419) is `jointbpp.compute_bpp` when a fold
driver is given.  Without one every hit gets 1.0 — what HEAD computes, since its folds never receive the constraint
(SURVEY.md defect D7; every shipped `cluster_*.res` shows 1.00000E+00).
"""
from __future__ import annotations

from .candidates import filter_candidates
from .jointbpp import compute_bpp
from .results import write_results


class ScanSession:
    def __init__(self, scanner, seqs, out_file, min_score=5.0, min_bpp=0.001, fold=None, comments="",
                 seq_file=None, both_strands=False, n_sequences=None, from_stdin=False):
        """scanner: a `Scanner` with `config` and `rna_tools` set; seqs: the object `scan` takes (`sequence_list`).
        seq_file / both_strands / n_sequences feed the header (lib/results.py:24-34); by default they are read from
        `seqs` (`fname`, `both_strands`, `seq_count()`), as `Results.write` does."""
        self.scanner, self.seqs, self.out_file = scanner, seqs, out_file
        self.min_score, self.min_bpp, self.fold, self.comments = min_score, min_bpp, fold, comments
        self.seq_file = seq_file if seq_file is not None else getattr(seqs, "fname", None)
        self.both_strands = both_strands or bool(getattr(seqs, "both_strands", False))
        self.n_sequences = n_sequences if n_sequences is not None else (
            seqs.seq_count() if hasattr(seqs, "seq_count") else len(seqs.sequence_list))
        self.from_stdin = from_stdin or bool(getattr(seqs, "from_stdin", False))
        self.last_cand = 0
        self.saves = []                           # (mode, hits written): what the reference's stderr chatter would show

    # rmdetect.py:31-70
    def save(self, cands, tries, mode):
        cands = filter_candidates(cands, min_score=self.min_score)
        if self.fold is not None:
            by_key = {q.key: q.seq for q in self.seqs.sequence_list}
            cands = compute_bpp(cands, by_key.__getitem__, self.fold, constraint=self.scanner.config.constraint)
        else:
            for c in cands:
                if c.bpp is None:
                    c.bpp = 1.0
        cands = filter_candidates(cands, min_bpp=self.min_bpp)
        cands.sort()
        write_results(self.out_file, cands, mode, tries,
                      seq_file=None if self.from_stdin else self.seq_file,
                      seq=self.seqs.sequence_list[0].seq if self.from_stdin else None,
                      both_strands=self.both_strands, n_sequences=self.n_sequences,
                      constraint=self.scanner.config.constraint, comments="" if mode == "a" else self.comments)
        self.saves.append((mode, len(cands)))
        return cands

    def run(self):
        """rmdetect.py:237-249: returns (candidates, tries) of `scan`; the result file is complete afterwards."""
        sc = self.scanner

        def before_scan(total):
            if self.out_file is not None:
                self.save([], [0, 0], "n")

        def before_sequence(total, n, key):
            self.last_cand = 0

        def after_window(cands, tries):
            if len(cands) - self.last_cand > 1000 and self.out_file is not None:
                self.save(cands[self.last_cand:], tries, "a")
                self.last_cand = len(cands)

        def after_sequence(cands, tries):
            if self.out_file is not None:
                self.save(cands[self.last_cand:], tries, "a")

        def after_scan(cands, tries):
            self.final = self.save(cands, tries, "w")

        sc.cback_before_scan, sc.cback_before_sequence = before_scan, before_sequence
        sc.cback_before_window = lambda total, n: 0
        sc.cback_after_window, sc.cback_after_sequence, sc.cback_after_scan = after_window, after_sequence, after_scan
        return sc.scan(self.seqs)
