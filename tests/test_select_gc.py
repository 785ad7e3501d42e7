"""EValues.select_gc settles in bulk which classes can be nearest and lets the reference's own arithmetic decide among them
(lib/evalues.py:12-23): the choice must be the plain loop's for every GC table, with ties, missing and extra letters."""
import math
import random

from helpers import load_evalues
from rmdetect_b200.evalues import EValues


def plain(classes, gc):
    sel, min_d = None, None
    for i, c in enumerate(classes):
        x = 0.0
        for k, v in c.items():
            x += (v - math.e ** gc.get(k, 0)) ** 2
        d = math.sqrt(x)
        if min_d is None or d < min_d:                  # strict: the first minimum wins
            min_d, sel = d, i
    return sel


def tables(rng, n):
    for t in range(n):
        f = [rng.random() for _ in range(4)]
        s = sum(f)
        gc = {k: math.log(x / s) for k, x in zip("ACGU", f)}
        if t % 7 == 0:
            del gc["G"]                                 # a letter the sequence lacks counts as e**0 (the reference's quirk)
        if t % 11 == 0:
            gc["N"] = math.log(0.01)
        if t % 13 == 0:
            gc = {k: math.log(0.25) for k in "ACGU"}    # the centre: many classes at equal distance
        yield gc


def test_shipped_classes():
    ev = load_evalues("KT_1.0")
    for gc in tables(random.Random(1), 3000):
        ev._last_gc = None
        assert ev.select_gc(gc) == plain(ev.gc_classes, gc)


def test_duplicated_and_short_classes():
    src = load_evalues("CL_1.0")
    ev = EValues()
    ev.gc_classes = [dict(c) for c in src.gc_classes[:20]] + [dict(src.gc_classes[3])] + [{"A": 0.5, "C": 0.5}, {"U": 0.9}]
    ev.gc_values = [{} for _ in ev.gc_classes]
    for gc in tables(random.Random(2), 2000):
        ev._last_gc = None
        assert ev.select_gc(gc) == plain(ev.gc_classes, gc)
    ev.gc_classes = ev.gc_classes[5:9]                  # the list changes: the dense form follows
    ev._last_gc = None
    gc = {k: math.log(0.25) for k in "ACGU"}
    assert ev.select_gc(gc) == plain(ev.gc_classes, gc)
    empty = EValues()
    assert empty.select_gc(gc) is None
