#!/usr/bin/env python
"""Wall time of the drop-in `Scanner.scan` on the reference's alignments, one GPU job per sequence (round 1) against one job
per alignment (round 2): lys.stk (49 sequences), rnaselci.stk (26), arich_test.stk (100), stand-in bpp provider.

  python tools/scanner_batching.py > profiles/rNN_scanner_batching.txt
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import Cfg, SeqList, fresh_models, golden            # noqa: E402
from rmdetect_b200.scanner import Scanner                         # noqa: E402
from rmdetect_b200.synth import SynthFold                         # noqa: E402


class CachedFold:
    """The provider's own time is not the subject: every window's list is computed once, before the timed scans."""

    def __init__(self, inner):
        self.inner, self.cache = inner, {}

    def coo(self, win):
        if win not in self.cache:
            self.cache[win] = self.inner.coo(win)
        return self.cache[win]

    def coo_many(self, wins, constraint=""):
        return [self.coo(w) for w in wins]


if __name__ == "__main__":
    for name, models in (("lys.stk", ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]), ("rnaselci.stk", ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]),
                         ("arich_test.stk", ["ARICH_1.0"])):
        triples = golden("seqsets")[name]
        ms, evs = fresh_models(models)
        fold = CachedFold(SynthFold(7))
        out = {}
        for label, cap in (("one job per sequence", 1), ("one job per alignment", 1 << 19)):
            sc = Scanner(Cfg(ms, evs))
            sc.rna_tools = fold
            sc.max_batch_windows = cap
            sc.scan(SeqList(triples))                                  # warm-up: library load, model upload, provider cache
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                cands, tries = sc.scan(SeqList(triples))
                best = min(best, time.perf_counter() - t0)
            out[label] = (best, len(cands), tries)
            sc._engine.close()
        (a, na, ta), (b, nb, tb) = out["one job per sequence"], out["one job per alignment"]
        assert (na, ta) == (nb, tb)
        print("%-16s %3d sequences, %d models: %7.2f ms one job per sequence -> %6.2f ms one job per alignment (%.1fx), %d hits, tries %s"
              % (name, len(triples), len(models), 1e3 * a, 1e3 * b, a / b, na, ta))
