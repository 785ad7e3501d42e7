"""Read base-pair probabilities out of a ViennaRNA `dot.ps` file.

`RNAfold -p` writes one line `i j sqrt(p) ubox` per pair of the ensemble
(1-based, i < j); `lbox` lines describe the MFE structure and are ignored.
The reference parses the same lines (lib/rnatools.py:176-205, `p = sqrt_p ** 2.0`)
but opens the wrong file for current ViennaRNA releases (SURVEY.md defect D3);
this reader takes the file that actually holds the data.  Returns the sorted
upper-triangle coordinate list the kernels consume.
"""
from __future__ import annotations

import numpy as np


def read_dot_ps(path):
    """(sequence or None, i uint16[], j uint16[], p float64[]) — 0-based, sorted by (i, j)."""
    ii, jj, pp = [], [], []
    seq, in_seq = None, False
    with open(path) as fh:
        for line in fh:
            s = line.strip()
            if s.startswith("/sequence"):
                in_seq, seq = True, ""
                continue
            if in_seq:
                if s.startswith(")"):
                    in_seq = False
                else:
                    seq += s.rstrip("\\")
                continue
            if s.endswith("ubox") and not s.startswith("%") and not s.startswith("/"):
                tok = s.split()
                if len(tok) == 4:
                    p = float(tok[2]) ** 2.0
                    if p != 0.0:
                        ii.append(int(tok[0]) - 1)
                        jj.append(int(tok[1]) - 1)
                        pp.append(p)
    i = np.asarray(ii, dtype=np.int64)
    j = np.asarray(jj, dtype=np.int64)
    p = np.asarray(pp, dtype=np.float64)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    # a later line for the same pair overwrites an earlier one (BPProbs.add semantics)
    order = np.lexsort((np.arange(len(lo)), hi, lo))
    lo, hi, p = lo[order], hi[order], p[order]
    if len(lo):
        last = np.ones(len(lo), dtype=bool)
        last[:-1] = (lo[1:] != lo[:-1]) | (hi[1:] != hi[:-1])
        lo, hi, p = lo[last], hi[last], p[last]
    return seq, lo.astype(np.uint16), hi.astype(np.uint16), p


def read_dot_ps_native(path):
    """Same result as read_dot_ps through the library's parser (rmd_parse_dot_ps), plus the compact
    transport words q (0 where an entry has no compact form): (sequence, i, j, p, q)."""
    from ._native import parse_dot_ps
    with open(path, "rb") as fh:
        return parse_dot_ps(fh.read())
