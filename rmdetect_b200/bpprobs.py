"""Base-pair-probability matrices at the boundary of the scan path.

`BPProbs` keeps the reference's container contract (lib/bpprobs.py:6-18):
symmetric, zeros never stored, a later `add` of the same pair overwrites.  The
kernels take one window's matrix as a coordinate list of its upper triangle
sorted by (i, j); `to_coo` produces that from a `BPProbs`, from the reference's
own class (same `.matrix` dict-of-dicts), or from anything exposing
`coo() -> (i, j, p)`.
"""
from __future__ import annotations

import numpy as np


class BPProbs:
    def __init__(self):
        self.bases = {}
        self.matrix = {}

    def add(self, b1, b2, p):
        if p != 0.0:
            self.matrix.setdefault(b1, {})[b2] = p
            self.matrix.setdefault(b2, {})[b1] = p
            self.bases[b1] = self.bases.get(b1, 0.0) + p
            self.bases[b2] = self.bases.get(b2, 0.0) + p

    def prob_pair(self, b1, b2):
        return self.matrix.get(b1, {}).get(b2, 0.0)


def to_coo(bpp, win_len):
    """(i uint16[], j uint16[], p float64[]) with i < j < win_len, sorted by (i, j).

    Entries the gate can never read (an index outside the window, or the
    diagonal, which the symmetric store keeps once) are handled as follows:
    out-of-window entries are dropped — no placement reaches them because
    pointer + offset < window length (reference lib/scanner.py:173) — and a
    diagonal entry (i == i) cannot be produced by a fold and is rejected.
    """
    if hasattr(bpp, "coo"):
        i, j, p = bpp.coo()
        if (isinstance(i, np.ndarray) and i.dtype == np.uint16 and isinstance(j, np.ndarray) and j.dtype == np.uint16
                and isinstance(p, np.ndarray) and p.dtype == np.float64 and len(i) == len(j) == len(p)):
            # a provider that already hands over the kernel's form (fold.RNAfold's parser, synth.SynthFold): one check
            # that the list is a strictly increasing upper triangle inside the window with no zeros
            key = i.astype(np.int32) * 65536 + j
            if len(i) == 0 or ((i < j).all() and int(j.max()) < win_len and (p != 0.0).all() and (len(i) == 1 or (key[1:] > key[:-1]).all())):
                return i, j, p
        i = np.asarray(i, dtype=np.int64)
        j = np.asarray(j, dtype=np.int64)
        p = np.asarray(p, dtype=np.float64)
    else:
        ii, jj, pp = [], [], []
        for b1, row in bpp.matrix.items():
            for b2, p in row.items():
                if b1 < b2:
                    ii.append(b1); jj.append(b2); pp.append(p)
                elif b1 == b2:
                    raise ValueError("bpp matrix has a diagonal entry at %d" % b1)
        i = np.asarray(ii, dtype=np.int64)
        j = np.asarray(jj, dtype=np.int64)
        p = np.asarray(pp, dtype=np.float64)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    if (lo == hi).any():
        raise ValueError("bpp matrix has a diagonal entry")
    keep = (lo >= 0) & (hi < win_len) & (p != 0.0)
    lo, hi, p = lo[keep], hi[keep], p[keep]
    order = np.lexsort((hi, lo))
    lo, hi, p = lo[order], hi[order], p[order]
    if len(lo) > 1 and ((lo[1:] == lo[:-1]) & (hi[1:] == hi[:-1])).any():
        raise ValueError("bpp coordinate list repeats a pair")
    return lo.astype(np.uint16), hi.astype(np.uint16), p
