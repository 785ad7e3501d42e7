#!/usr/bin/env python
"""cProfile of the drop-in `Scanner.scan` on lys.stk (49 sequences, four models, cached stand-in bpp lists): where the host
side of a small scan spends its time (development aid)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import Cfg, SeqList, fresh_models, golden            # noqa: E402
from rmdetect_b200.scanner import Scanner                         # noqa: E402
from rmdetect_b200.synth import SynthFold                         # noqa: E402
from scanner_batching import CachedFold                           # noqa: E402,F401  (prints its own table on import)

name, models = sys.argv[1] if len(sys.argv) > 1 else "lys.stk", ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]
triples = golden("seqsets")[name]
ms, evs = fresh_models(models)
sc = Scanner(Cfg(ms, evs))
sc.rna_tools = CachedFold(SynthFold(7))
sc.scan(SeqList(triples))
sc.scan(SeqList(triples))
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    cands, tries = sc.scan(SeqList(triples))
pr.disable()
print("%s: %.2f ms per scan under the profiler, %d hits" % (name, 1e3 * (time.perf_counter() - t0) / 5, len(cands)))
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
sc._engine.close()
