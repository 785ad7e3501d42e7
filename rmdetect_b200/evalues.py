"""`.evalues` tables: parser, GC-class selection and the dense device form.

File format and lookup semantics are the reference's (lib/evalues.py:12-80):
`[GC_CLASSES]` rows then `[E_VALUES]` rows keyed by the literal score string.
`EValues.compute` looks a score up under the key ``"%.1f" % round(score, 1)``
(``"-0.0"`` folded to ``"0.0"``); a key that is not in the table gives 0.0, a
table that was never loaded gives 1.0.

The device cannot format strings, so the lookup is restated as an integer
bucket: bucket(x) = number of thresholds R_k <= |x|, where R_k is the smallest
double that Python's correctly rounded ``round(x, 1)`` sends to (k+1)/10 or
above (`round_thresholds`).  Row k of the dense table holds the value stored
under ``"%.1f" % (k / 10)``.
"""
from __future__ import annotations

import math
import os
from fractions import Fraction

import numpy as np


_CLASS_SETS = {}                  # class lists seen so far -> id (EValues.class_set_id)


class EValues:
    def __init__(self):
        self.gc_classes = []      # list of {letter: fraction}
        self.gc_values = []       # per class: {score string: e-value}
        self.selected = None

    def select_gc(self, gc):
        # reference lib/evalues.py:12-23: Euclidean distance over the class's own letters,
        # a letter absent from `gc` counts as e**0 = 1; strict '<' so the first minimum wins.
        # Same floating-point operations in the same order as the reference; only the four `e ** gc[k]` are taken
        # once per call instead of once per class (same function, same argument: same bits).
        items = self._class_items()
        last = getattr(self, "_last_gc", None)
        key = tuple(gc.items())
        if last is not None and last[0] == key and last[1] is items:
            self.selected = last[2]
            return self.selected
        # Which classes can be the minimum is settled in bulk (numpy, rounding differences of a few ulp); the reference's own
        # arithmetic — same operations, same order, `** 2` through pow, strict '<' in class order — then decides among the
        # classes within 1e-9 of the smallest distance, normally one.  Only the four `e ** gc[k]` are taken once per call
        # instead of once per class (same function, same argument: same bits).
        power = {}
        rows = range(len(items))
        dense = self._class_dense(items)
        if dense is not None:
            letters, values, present = dense
            e = np.array([math.e ** gc.get(k, 0) for k in letters], dtype=np.float64)
            approx = np.sqrt((np.where(present, values - e, 0.0) ** 2).sum(axis=1))
            if np.all(np.isfinite(approx)):
                rows = np.flatnonzero(approx <= approx.min() * (1.0 + 1e-9) + 1e-300).tolist()
        min_d = None
        for i in rows:
            x = 0.0
            for k, v in items[i]:
                e_k = power.get(k)
                if e_k is None:
                    e_k = power[k] = math.e ** gc.get(k, 0)
                x += (v - e_k) ** 2
            d = math.sqrt(x)
            if min_d is None or d < min_d:
                min_d = d
                self.selected = i
        self._last_gc = (key, items, self.selected)
        return self.selected

    def _class_dense(self, items):
        """(letters, values [class][letter], present [class][letter]) of the classes, rebuilt when they change; None when
        there is no class."""
        cached = getattr(self, "_dense", None)
        if cached is None or cached[0] is not items:
            if not items:
                return None
            letters = []
            for pairs in items:
                for k, _ in pairs:
                    if k not in letters:
                        letters.append(k)
            at = {k: n for n, k in enumerate(letters)}
            values = np.zeros((len(items), len(letters)))
            present = np.zeros((len(items), len(letters)), dtype=bool)
            for i, pairs in enumerate(items):
                for k, v in pairs:
                    values[i, at[k]] = v
                    present[i, at[k]] = True
            cached = (items, (tuple(letters), values, present))
            self._dense = cached
        return cached[1]

    def _class_items(self):
        """gc_classes as a tuple of ((letter, value), ...) in dict order, rebuilt when the list changes."""
        cached = getattr(self, "_items", None)
        if cached is None or cached[0] is not self.gc_classes or cached[1] != len(self.gc_classes):
            items = tuple(tuple(c.items()) for c in self.gc_classes)
            cached = (self.gc_classes, len(self.gc_classes), items, _CLASS_SETS.setdefault(items, len(_CLASS_SETS)))
            self._items = cached
        return cached[2]

    def class_set_id(self):
        """A small integer that is equal for tables calibrated on the same GC classes (same letters, values and order): what
        decides whether two models select the same class for a sequence.  The deep comparison is made once per table."""
        self._class_items()
        return self._items[3]

    def compute(self, score):
        # reference lib/evalues.py:25-36 (host restatement, used by the drop-in API and tests)
        if self.selected is None:
            return 1.0
        key = "%.1f" % round(score, 1)
        if key == "-0.0":
            key = "0.0"
        return self.gc_values[self.selected].get(key, 0.0)

    def dense(self):
        """float64[n_classes][n_rows]; row k = value under the key "%.1f" % (k/10)."""
        n_class = len(self.gc_classes)
        if n_class == 0:
            return np.zeros((0, 0), dtype=np.float64)
        top = -1
        for tab in self.gc_values:
            for key in tab:
                k = _bucket_of_key(key)
                if k is not None:
                    top = max(top, k)
        out = np.zeros((n_class, top + 1), dtype=np.float64)
        for c, tab in enumerate(self.gc_values):
            for key, val in tab.items():
                k = _bucket_of_key(key)
                if k is not None:
                    out[c, k] = val
        return out


def _bucket_of_key(key):
    """Bucket index of a table key the lookup can actually produce, else None."""
    try:
        val = float(key)
    except ValueError:
        return None
    k = int(round(val * 10))
    if k < 0 or ("%.1f" % (k / 10.0)) != key:
        return None
    return k


def parse_evalues(fname) -> EValues:
    """reference lib/evalues.py:38-80; a missing file yields an empty table (E-value 1.0)."""
    result = EValues()
    if not os.path.isfile(fname):
        return result
    state = 0
    with open(fname) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            if "GC_CLASSES" in line:
                state = 1
            elif "E_VALUES" in line:
                state = 2
            elif state == 1:
                tok = line.split()
                result.gc_classes.append({tok[i]: float(tok[i + 1]) for i in range(1, len(tok), 2)})
                result.gc_values.append({})
            elif state == 2:
                tok = line.split()
                for i in range(1, len(tok)):
                    result.gc_values[i - 1][tok[0]] = float(tok[i])
    return result


def round_thresholds(n):
    """R_0..R_{n-1}: x >= R_k  <=>  round(x, 1) >= (k+1)/10, for doubles x >= 0.

    Python rounds the exact binary value of x to one decimal, ties to even
    (Objects/floatobject.c double_round via correctly rounded dtoa).  The real
    midpoint m_k = (2k+1)/20 is a double only at x.25 / x.75; there the tie goes
    to the even tenth.  Computed with exact rationals.
    """
    out = np.empty(n, dtype=np.float64)
    for k in range(n):
        mid = Fraction(2 * k + 1, 20)
        d = float(mid)                       # nearest double to the midpoint
        fd = Fraction(d)
        if fd == mid:
            goes_up = (k + 1) % 2 == 0       # tie -> even tenth
            out[k] = d if goes_up else math.nextafter(d, math.inf)
        elif fd > mid:
            out[k] = d
        else:
            out[k] = math.nextafter(d, math.inf)
    return out


def bucket(score, thresholds):
    """Host twin of the device bucket routine: (bucket index, in_table_domain)."""
    a = abs(score)
    k = int(np.searchsorted(thresholds, a, side="right"))
    if score < 0 and k > 0:
        return k, False            # key "-0.1", … never matches a table row
    return k, True
