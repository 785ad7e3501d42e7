"""Background letter frequencies (reference lib/gcx.py).

`GC.compute` gives every distinct character of the *whole* sequence its own
log-frequency; the scorer uses it both for the random model and for
GC-sentinel CPT rows.  The device sees the same numbers as a 16-entry table
indexed by letter code (A C G U gap other1..other11).
"""
from __future__ import annotations

import math

import numpy as np

BASES = "ACGU"
N_CODES = 16           # 4 bases + gap + up to 11 other letters (the IUPAC ambiguity set)
CODE_GAP = 4
FIRST_OTHER = 5


class GC:
    @staticmethod
    def compute(sequence):
        # reference lib/gcx.py:4-13 — counts as floats, log(count / len)
        counts = {}
        for c in sequence:
            counts[c] = counts.get(c, 0.0) + 1.0
        n = len(sequence)
        return {nt: math.log(cnt / n) for nt, cnt in counts.items()}

    @staticmethod
    def neutral():
        p = math.log(0.25)
        return {"A": p, "C": p, "G": p, "U": p}

    @staticmethod
    def specific(gc):
        pgc = math.log(gc / 2.0)
        pau = math.log((1.0 - gc) / 2.0)
        return {"A": pau, "C": pgc, "G": pgc, "U": pau}

    @staticmethod
    def get_classes(resolution=10, min_p=10, max_p=90):
        # reference lib/gcx.py:25-48: all (a,c,g,u) percent tuples on the grid summing to 100,
        # a outermost … u innermost
        grid = range(min_p, max_p + 1, resolution)
        out = []
        for a in grid:
            for c in grid:
                for g in grid:
                    for u in grid:
                        if a + c + g + u == 100:
                            out.append({"A": a / 100, "C": c / 100, "G": g / 100, "U": u / 100})
        return out


def encode_sequence(seq: str):
    """Letter codes for one sequence.

    Returns (codes uint8[n], others str) where codes 0..3 are A C G U and code
    5+k is the k-th distinct non-ACGU character in order of first appearance.
    """
    raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
    lut = np.full(256, 255, dtype=np.uint8)
    for k, b in enumerate(BASES):
        lut[ord(b)] = k
    codes = lut[raw]
    others = ""
    if (codes == 255).any():
        odd = raw[codes == 255]
        _, first = np.unique(odd, return_index=True)
        for ch in odd[np.sort(first)]:
            if len(others) >= N_CODES - FIRST_OTHER:
                raise ValueError("more than %d distinct non-ACGU characters in a sequence"
                                 % (N_CODES - FIRST_OTHER))
            lut[ch] = FIRST_OTHER + len(others)
            others += chr(ch)
        codes = lut[raw]
    return codes, others


def gc_table(gc: dict, others: str = ""):
    """16-entry fp64 table for the device.

    Slot 4 (gap) holds log 0.25, the reference's fallback for a letter missing
    from the GC dict (lib/node.py:223); so do the four bases when the caller's
    dict lacks them (`scan(seqs, gc=...)` with a partial dict).
    """
    fallback = math.log(0.25)
    tab = np.full(N_CODES, fallback, dtype=np.float64)
    for k, b in enumerate(BASES):
        tab[k] = gc.get(b, fallback)
    for k, ch in enumerate(others):
        tab[FIRST_OTHER + k] = gc.get(ch, fallback)
    return tab
