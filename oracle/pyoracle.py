"""ctypes front end of the CPU oracle (oracle/rmd_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  Nothing under rmdetect_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librmd_oracle.so")
MAXC, MAXL = 4, 64


def build(force=False):
    src = os.path.join(HERE, "rmd_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"], env={k: v for k, v in os.environ.items() if k != "CC"})
    return LIB_PATH


class GcDict(C.Structure):
    _fields_ = [("v", C.c_double * 256), ("has", C.c_ubyte * 256)]

    @classmethod
    def from_dict(cls, gc):
        out = cls()
        for ch, val in gc.items():
            out.v[ord(ch)] = val
            out.has[ord(ch)] = 1
        return out

    def to_dict(self):
        return {chr(c): self.v[c] for c in range(256) if self.has[c]}


class Solution(C.Structure):
    _fields_ = [("ok", C.c_int), ("cfg", C.c_int), ("fault", C.c_int),
                ("prob_model", C.c_double), ("prob_rand", C.c_double),
                ("seqs", (C.c_char * MAXL) * MAXC), ("pos", (C.c_int * MAXL) * MAXC)]


class Hit(C.Structure):
    _fields_ = [("seq", C.c_int32), ("model", C.c_int32),
                ("coord0", C.c_int64), ("coord1", C.c_int64),
                ("p", C.c_int32 * MAXC), ("cfg", C.c_int32),
                ("prob_model", C.c_double), ("prob_rand", C.c_double),
                ("score", C.c_double), ("evalue", C.c_double),
                ("seqs", (C.c_char * MAXL) * MAXC), ("pos", (C.c_int64 * MAXL) * MAXC)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_model_load.restype = C.c_void_p
        L.orc_model_load.argtypes = [C.c_char_p]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_tree_size.argtypes = [C.c_void_p]
        L.orc_gc_compute.argtypes = [C.c_char_p, C.c_int64, C.POINTER(GcDict)]
        L.orc_eval.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(GcDict),
                               C.c_double, C.c_double, C.POINTER(Solution)]
        L.orc_evalues_load.restype = C.c_void_p
        L.orc_evalues_load.argtypes = [C.c_char_p]
        L.orc_evalues_free.argtypes = [C.c_void_p]
        L.orc_evalues_nclass.argtypes = [C.c_void_p]
        L.orc_evalues_nrows.argtypes = [C.c_void_p]
        L.orc_select_gc.argtypes = [C.c_void_p, C.POINTER(GcDict)]
        L.orc_set_selected.argtypes = [C.c_void_p, C.c_int]
        L.orc_evalue.restype = C.c_double
        L.orc_evalue.argtypes = [C.c_void_p, C.c_double]
        L.orc_windows.restype = C.c_int64
        L.orc_windows.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.orc_scan_sequence.restype = C.c_int64
        L.orc_scan_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int64,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_double, C.c_double,
                                        C.POINTER(C.POINTER(Hit)), C.POINTER(C.c_int64), C.c_void_p,
                                        C.POINTER(C.c_int), C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_window_key.restype = C.c_uint64
        L.orc_window_key.argtypes = [C.c_char_p, C.c_int, C.c_uint64]
        L.orc_synth_bpp.restype = C.c_int64
        L.orc_synth_bpp.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_bench_scan.restype = C.c_int64
        L.orc_bench_scan.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int64, C.c_int, C.c_int,
                                     C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.POINTER(C.c_int64)]
        L.orc_bench_scan2.restype = C.c_int64
        L.orc_bench_scan2.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int64, C.c_int, C.c_int,
                                      C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                      C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.orc_calibrate.restype = C.c_int64
        L.orc_calibrate.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                    C.c_double, C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        assert L.orc_sizeof_hit() == C.sizeof(Hit), (L.orc_sizeof_hit(), C.sizeof(Hit))
        assert L.orc_sizeof_solution() == C.sizeof(Solution)
        assert L.orc_sizeof_gcdict() == C.sizeof(GcDict)
        _lib = L
    return _lib


def dump_model(model) -> str:
    """Text form of a reference-shaped Model object for `orc_model_load`.

    Mirrors the reference's attributes (nodes/probs, ORDER, OFFSETS, pairs_list); floats travel
    as C99 hex so no digit is lost.
    """
    nch = model.chains_count
    out = ["MODEL %s %s" % (model.name, model.version),
           "CHAINS %d %s" % (nch, " ".join(str(model.chains_length[c]) for c in range(nch))),
           "SYMMETRIC %d" % int(bool(model.symmetric)),
           "SEPMIN %d" % model.sep_min,
           "NODES %d" % len(model.nodes)]
    for n in model.nodes:
        conds = [] if n.conds is None else list(n.conds)
        out.append("N %d %d %d %d %s %d" % (n.id, n.chain, n.ndx, len(conds),
                                            " ".join(map(str, conds)), len(n.probs)))
        for key, row in n.probs.items():
            if row == "GC":
                out.append("R %s GC" % key)
            else:
                vals = [float(row[ch]).hex() if ch in row else "-" for ch in "ACGU."]
                out.append("R %s T %s" % (key, " ".join(vals)))
    out.append("ORDER " + " ".join(str(n.id) for n in model.order))
    out.append("OFFSETS %d" % len(model.OFFSETS))
    for ci, cfg in enumerate(model.OFFSETS):
        for ch in range(nch):
            out.append("O %d %d %s" % (ci, ch, " ".join(str(-1 if o is None else o) for o in cfg[ch])))
    plist = list(model.pairs_list)
    out.append("PAIRCFG %d %d" % (len(plist), sum(len(c) for c in plist)))
    for cfg in plist:
        out.append("P %d %s" % (len(cfg), " ".join("%d %d %d %d %d" % tuple(p) for p in cfg)))
    return "\n".join(out) + "\n"


class OracleModel:
    def __init__(self, model):
        self.model = model
        self.dump = dump_model(model).encode()
        self.h = lib().orc_model_load(self.dump)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_model_free(self.h)
            self.h = None

    def eval(self, seq, pointers, gc, unknown_prob, cutoff):
        """(ok, seqs, poss, prob_model, prob_rand) like eval + get_solution of the reference."""
        sol = Solution()
        ptr = (C.c_int * MAXC)(*list(pointers))
        g = GcDict.from_dict(gc)
        b = seq.encode("latin-1")
        ok = lib().orc_eval(self.h, b, len(b), ptr, C.byref(g), unknown_prob, cutoff, C.byref(sol))
        if sol.fault:
            raise IndexError("the reference would raise for this placement")
        if not ok:
            return False, None, None, None, None
        nch = self.model.chains_count
        seqs, poss = [], []
        for c in range(nch):
            L = self.model.chains_length[c]
            seqs.append(sol.seqs[c].raw[:L].decode())
            poss.append([None if sol.pos[c][k] < 0 else sol.pos[c][k] for k in range(L)])
        return True, seqs, poss, sol.prob_model, sol.prob_rand


class OracleEValues:
    def __init__(self, path):
        self.h = lib().orc_evalues_load(os.fsencode(path))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_evalues_free(self.h)
            self.h = None

    def select_gc(self, gc):
        return lib().orc_select_gc(self.h, C.byref(GcDict.from_dict(gc)))

    def set_selected(self, cls):
        lib().orc_set_selected(self.h, -1 if cls is None else cls)

    def compute(self, score):
        return lib().orc_evalue(self.h, float(score))


def gc_compute(seq):
    g = GcDict()
    b = seq.encode("latin-1")
    lib().orc_gc_compute(b, len(b), C.byref(g))
    return g.to_dict()


def windows(seqlen, win_len, win_step):
    n = lib().orc_windows(seqlen, win_len, win_step, None, 0)
    starts = np.zeros(n, dtype=np.int64)
    lib().orc_windows(seqlen, win_len, win_step, starts.ctypes.data, n)
    return starts


def synth_bpp(seq, seed):
    b = seq.encode("latin-1")
    cap = max(16, len(b) * len(b) // 2)
    i = np.zeros(cap, dtype=np.uint16)
    j = np.zeros(cap, dtype=np.uint16)
    p = np.zeros(cap, dtype=np.float64)
    n = lib().orc_synth_bpp(b, len(b), seed, i.ctypes.data, j.ctypes.data, p.ctypes.data, cap)
    return i[:n].copy(), j[:n].copy(), p[:n].copy()


def scan_sequence(omodels, oevalues, seq_id, seq, win_len, win_step, bpp_list,
                  min_bpp_mean, unknown_prob, cutoff, gc=None):
    """Scan one sequence.  `bpp_list[k]` = (i, j, p) arrays of window k (one entry in whole-sequence mode).

    Returns (hits: list of dict, tries [all, pairs], tries per window int64[nwin][2]).
    """
    nm = len(omodels)
    marr = (C.c_void_p * nm)(*[m.h for m in omodels])
    earr = (C.c_void_p * nm)(*[e.h for e in oevalues]) if oevalues is not None else None
    off = np.zeros(len(bpp_list) + 1, dtype=np.int64)
    for k, (i, _, _) in enumerate(bpp_list):
        off[k + 1] = off[k] + len(i)
    cat = lambda idx, dt: (np.concatenate([np.asarray(b[idx], dtype=dt) for b in bpp_list])
                           if bpp_list else np.zeros(0, dtype=dt))
    bi, bj, bp = cat(0, np.uint16), cat(1, np.uint16), cat(2, np.float64)
    bi, bj, bp = (np.ascontiguousarray(np.append(a, a.dtype.type(0))) for a in (bi, bj, bp))
    out = C.POINTER(Hit)()
    tries = (C.c_int64 * 2)()
    tpw = np.zeros((max(1, len(bpp_list)), 2), dtype=np.int64)
    fault = C.c_int(0)
    b = seq.encode("latin-1")
    g = GcDict.from_dict(gc) if gc is not None else None
    n = lib().orc_scan_sequence(marr, earr, nm, seq_id, b, len(b), win_len, win_step,
                                off.ctypes.data, bi.ctypes.data, bj.ctypes.data, bp.ctypes.data,
                                min_bpp_mean, unknown_prob, cutoff, C.byref(out), tries, tpw.ctypes.data,
                                C.byref(fault), C.addressof(g) if g is not None else None)
    if fault.value:
        lib().orc_free(out)
        raise IndexError("the reference would raise on this input")
    hits = []
    for k in range(n):
        h = out[k]
        m = omodels[h.model].model
        seqs, pos = [], []
        for c in range(m.chains_count):
            L = m.chains_length[c]
            seqs.append(h.seqs[c].raw[:L].decode())
            pos.append([None if h.pos[c][q] < 0 else h.pos[c][q] for q in range(L)])
        hits.append(dict(seq=h.seq, model=h.model, coords=[h.coord0, h.coord1], p=[h.p[0], h.p[1]], cfg=h.cfg,
                         prob_model=h.prob_model, prob_rand=h.prob_rand, score=h.score, evalue=h.evalue,
                         seqs=seqs, pos=pos))
    lib().orc_free(out)
    return hits, [tries[0], tries[1]], tpw


def bench_scan(models, seq, win_len, win_step, w0, w1, seed, nthreads, min_bpp_mean, unknown_prob, cutoff):
    dumps = [dump_model(m).encode() for m in models]
    arr = (C.c_char_p * len(dumps))(*dumps)
    tries = (C.c_int64 * 2)()
    b = seq.encode("latin-1")
    nh = lib().orc_bench_scan(len(dumps), arr, b, len(b), win_len, win_step, w0, w1, seed, nthreads,
                              min_bpp_mean, unknown_prob, cutoff, tries)
    return nh, [tries[0], tries[1]]


def bench_scan2(models, seq, win_len, win_step, w0, w1, seed, nthreads, min_bpp_mean, unknown_prob, cutoff, counts=None):
    """Windows [w0, w1) of `seq` with the stand-in bpp twin on `nthreads` cores.  `counts` = (A, C, G, U) letter
    counts of the whole sequence the sample was cut from (its GC table); returns dict(hits, tries, checksum,
    s_generate, s_scan): bpp generation and scan are timed apart (omp_get_wtime inside the C driver)."""
    dumps = [dump_model(m).encode() for m in models]
    arr = (C.c_char_p * len(dumps))(*dumps)
    tries = (C.c_int64 * 2)()
    secs = (C.c_double * 2)()
    chk = C.c_uint64(0)
    cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.int64)
    b = seq.encode("latin-1")
    nh = lib().orc_bench_scan2(len(dumps), arr, b, len(b), win_len, win_step, w0, w1, seed, nthreads,
                               None if cnt is None else cnt.ctypes.data, min_bpp_mean, unknown_prob, cutoff,
                               tries, C.byref(chk), secs)
    return dict(hits=nh, tries=[tries[0], tries[1]], checksum=chk.value, s_generate=secs[0], s_scan=secs[1])


def calibrate(omodel, samples, classes, unknown_prob, cutoff, nbins=512, hist=None):
    """Accumulate rmcalibrate histograms for `samples` (list of equal-length strings)."""
    L = len(samples[0])
    blob = "".join(samples).encode("latin-1")
    cls = np.ascontiguousarray([[c["A"], c["C"], c["G"], c["U"]] for c in classes], dtype=np.float64)
    if hist is None:
        hist = np.zeros((len(classes), nbins), dtype=np.float64)
    ovf = C.c_int(0)
    nok = lib().orc_calibrate(omodel.h, blob, len(samples), L, cls.ctypes.data, len(classes),
                              unknown_prob, cutoff, hist.ctypes.data, nbins, C.byref(ovf))
    if ovf.value:
        raise OverflowError("score beyond the histogram range")
    return hist, nok
