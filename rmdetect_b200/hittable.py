"""Hit lists that stay on the GPU: rmfilter's selections and rmcluster's bookkeeping on a device-resident table.

Mirrors the reference's `Candidates` static methods (lib/candidates.py:78-207) on `HitTable`: `filter`, `top_model`,
`top_seq`, `top_seq_model`, `get_model`, `nooverlap`, plus `sort` (`Candidate.__cmp__`, :380-394).  The rows come from
the last scan without leaving the device (`append_scan`) or from `Candidate` objects loaded from a `.res` file
(`append_candidates`); only what survives is downloaded (`rows`, `candidates`).  A genome-scale scan keeps 10^8 hits
(DESIGN.md §7): these never become Python objects.

The work is done by the library (include/rmdetect_b200.h: rmd_table_*); nothing here computes on the host beyond
naming sequences: key strings are ranked once (`str` order, what `__cmp__` compares) and handed over as integers.
"""
from __future__ import annotations

import ctypes as C
import math
import struct

import numpy as np

from ._native import ROW_DTYPE, _ptr
from .candidates import Candidate

LETTER_CODE = {"A": 0, "C": 1, "G": 2, "U": 3, ".": 4}
CODE_LETTER = "ACGU.N"


def pack_letters(chains_seqs):
    """Chain letters in node order, three bits each (rmd_hit_row.letters)."""
    v = 0
    k = 0
    for chain in chains_seqs:
        for ch in chain:
            v |= LETTER_CODE.get(ch, 5) << (3 * k)
            k += 1
    if k > 32:
        raise ValueError("a hit row holds the letters of at most 32 model positions")
    return [(v >> 0) & 0xffffffff, (v >> 32) & 0xffffffff, (v >> 64) & 0xffffffff]


def unpack_letters(words, n):
    v = int(words[0]) | (int(words[1]) << 32) | (int(words[2]) << 64)
    return [(v >> (3 * k)) & 7 for k in range(n)]


def evalue_bound(min_evalue):
    """Smallest double e with `-math.log(e) <= min_evalue`: the reference keeps a hit when its E-value is not 0 and
    `-math.log(evalue) > min_evalue` (lib/candidates.py:90), i.e. when evalue < this bound (log is monotone).  Found by
    bisection over the bit patterns of the positive doubles with the same libm the reference calls."""
    lo, hi = 1, struct.unpack("<q", struct.pack("<d", float("inf")))[0]          # smallest denormal .. +inf
    as_double = lambda b: struct.unpack("<d", struct.pack("<q", b))[0]
    ok = lambda b: b >= hi or -math.log(as_double(b)) <= min_evalue
    while lo < hi:
        mid = (lo + hi) // 2
        if ok(mid):
            hi = mid
        else:
            lo = mid + 1
    return as_double(lo)


class HitTable:
    """The device-resident hit table of one `ScanEngine` (its context, models and gap configurations)."""

    def __init__(self, engine):
        self.engine = engine
        self.ctx = engine.ctx
        self.lib = engine.ctx.lib
        self.keys = []                    # sequence id -> key string
        self._key_id = {}
        self.clear()

    # -- rows in ----------------------------------------------------------------------------------------------------
    def clear(self):
        self.ctx._check(self.lib.rmd_table_clear(self.ctx.h), "rmd_table_clear")
        self.keys, self._key_id = [], {}

    def __len__(self):
        n = C.c_int64(0)
        self.ctx._check(self.lib.rmd_table_rows(self.ctx.h, C.byref(n)), "rmd_table_rows")
        return n.value

    def key_id(self, key):
        if key not in self._key_id:
            self._key_id[key] = len(self.keys)
            self.keys.append(key)
        return self._key_id[key]

    def append_scan(self, keys, start_base=0, first_n=-1):
        """The hits the engine's last scan kept become rows, on the device.  keys: the names of the scanned job's
        sequences; start_base: position of the job's first letter in its sequence (a piece of a chromosome);
        first_n: only that many of the hits (a piece's halo windows come last and are not its own)."""
        ids = [self.key_id(k) for k in keys]
        if ids != list(range(ids[0], ids[0] + len(ids))):
            raise ValueError("append_scan: the job's sequences need consecutive ids (new names, or the same names in the same order)")
        n = C.c_int64(0)
        self.ctx._check(self.lib.rmd_table_append_scan(self.ctx.h, ids[0], int(start_base), int(first_n), C.byref(n)), "rmd_table_append_scan")
        return n.value

    def row_of(self, cand, ident=0):
        """rmd_hit_row of a `Candidate` (loaded from a `.res` file): pointers and gap configuration are recovered from
        its positions (lib/candidates.py:508-514 in reverse)."""
        m = next(i for i, mm in enumerate(self.engine.models) if mm is cand.model or mm.full_name() == cand.model.full_name())
        start = int(cand.coords[0])
        ptr = [int(ch[0]) - start for ch in cand.chains_pos]
        leaf = None
        for li, offs in enumerate(self.engine.flats[m].leaf_offsets):
            want = [[None if o is None else start + ptr[c] + o for o in offs[c]] for c in range(2)]
            if want == [list(ch) for ch in cand.chains_pos]:
                leaf = li
                break
        if leaf is None:
            raise ValueError("hit %s: positions match no gap configuration of %s" % (cand.key, cand.model.full_name()))
        row = np.zeros(1, dtype=ROW_DTYPE)[0]
        row["start"], row["key"], row["wlen"] = start, self.key_id(cand.key), int(cand.coords[1]) - start
        row["p0"], row["p1"], row["leaf"], row["model"] = ptr[0], ptr[1], leaf, m
        row["score"], row["evalue"] = cand.score, cand.evalue
        row["bpp"] = float("nan") if cand.bpp is None else cand.bpp
        row["letters"] = pack_letters(cand.chains_seqs)
        row["id"] = ident
        return row

    def append_candidates(self, cands):
        base = len(self)
        rows = np.zeros(len(cands), dtype=ROW_DTYPE)
        for k, c in enumerate(cands):
            rows[k] = self.row_of(c, base + k)
        n = C.c_int64(0)
        self.ctx._check(self.lib.rmd_table_append_rows(self.ctx.h, _ptr(rows), len(rows), C.byref(n)), "rmd_table_append_rows")
        return n.value

    # -- rows out ---------------------------------------------------------------------------------------------------
    def rows(self, first=0, n=None):
        total = len(self)
        n = total - first if n is None else n
        out = np.zeros(max(n, 1), dtype=ROW_DTYPE)
        self.ctx._check(self.lib.rmd_table_download(self.ctx.h, _ptr(out), first, n), "rmd_table_download")
        return out[:n]

    def candidates(self, cols_of=None, factory=Candidate):
        """`Candidate` objects of the rows (get_solution + set_data_build + update_chains_pos/cols); cols_of(key) gives
        the sequence's alignment column index or None."""
        out = []
        for r in self.rows():
            m = int(r["model"])
            model = self.engine.models[m]
            offs = self.engine.flats[m].leaf_offsets[int(r["leaf"])]
            start, ptr = int(r["start"]), (int(r["p0"]), int(r["p1"]))
            key = self.keys[int(r["key"])]
            cols = cols_of(key) if cols_of is not None else None
            codes = unpack_letters(r["letters"], sum(len(o) for o in offs))
            cand = factory(model, key, (start, start + int(r["wlen"])))
            k = 0
            for c in range(2):
                cand.chains_seqs.append([CODE_LETTER[codes[k + i]] for i in range(len(offs[c]))])
                cand.chains_pos.append([None if o is None else start + ptr[c] + o for o in offs[c]])
                k += len(offs[c])
            cand.chains_cols = [[None if p is None else (p if cols is None else cols[p]) for p in ch] for ch in cand.chains_pos]
            cand.score, cand.evalue = float(r["score"]), float(r["evalue"])
            cand.bpp = None if math.isnan(r["bpp"]) else float(r["bpp"])
            out.append(cand)
        return out

    # -- the reference's list operations ------------------------------------------------------------------------------
    def _ranks(self):
        order = {k: r for r, k in enumerate(sorted(set(self.keys)))}
        return np.ascontiguousarray([order[k] for k in self.keys] or [0], dtype=np.int32)

    def filter(self, min_score=None, min_bpp=None, min_evalue=None, sort=True):
        """`Candidates.filter` (lib/candidates.py:78-97): strict thresholds, then the sort."""
        d = lambda v: None if v is None else C.byref(C.c_double(v))
        n = C.c_int64(0)
        self.ctx._check(self.lib.rmd_table_filter(self.ctx.h, d(min_score), d(min_bpp),
                                                  d(None if min_evalue is None else evalue_bound(min_evalue)), -1, C.byref(n)), "rmd_table_filter")
        if sort:
            self.sort()
        return n.value

    def sort(self):
        """`cands.sort()` with `Candidate.__cmp__` (:380-394): key, E-value ascending, score descending; stable."""
        r = self._ranks()
        self.ctx._check(self.lib.rmd_table_sort(self.ctx.h, _ptr(r), len(r)), "rmd_table_sort")

    def first_unsorted(self):
        """The first row that compares smaller than its predecessor (`Candidate.__cmp__`), or -1: the table is sorted."""
        r = self._ranks()
        bad = C.c_int64(-1)
        self.ctx._check(self.lib.rmd_table_is_sorted(self.ctx.h, _ptr(r), len(r), C.byref(bad)), "rmd_table_is_sorted")
        return bad.value

    def _top(self, group, n_top):
        r = self._ranks()
        n = C.c_int64(0)
        self.ctx._check(self.lib.rmd_table_top(self.ctx.h, group, int(n_top), _ptr(r), len(r), C.byref(n)), "rmd_table_top")
        return n.value

    def top_model(self, n):
        return self._top(0, n)          # :99-116

    def top_seq(self, n):
        return self._top(1, n)          # :118-135

    def top_seq_model(self, n):
        return self._top(2, n)          # :137-154

    def get_model(self, model_name):
        """:156-169: sort, then the hits of one model."""
        self.sort()
        m = next((i for i, mm in enumerate(self.engine.models) if mm.full_name().upper() == model_name.upper()), None)
        n = C.c_int64(0)
        if m is None:
            m = len(self.engine.models)          # no such model: nothing stays
            self.ctx._check(self.lib.rmd_table_filter(self.ctx.h, C.byref(C.c_double(float("inf"))), None, None, -1, C.byref(n)), "rmd_table_filter")
            return n.value
        self.ctx._check(self.lib.rmd_table_filter(self.ctx.h, None, None, None, m, C.byref(n)), "rmd_table_filter")
        return n.value

    def nooverlap(self):
        """:171-207, in the table's present order."""
        n = C.c_int64(0)
        self.ctx._check(self.lib.rmd_table_nooverlap(self.ctx.h, C.byref(n)), "rmd_table_nooverlap")
        return n.value
