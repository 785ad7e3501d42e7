"""ctypes binding of librmdetect_b200.so (include/rmdetect_b200.h).

The library is built in-tree by `make -C rmdetect_b200/csrc` (or
`__graft_entry__.build()`).  If it is missing, or no CUDA device is usable,
every entry point raises: there is no CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .modelflat import PAIR_DTYPE, TREE_DTYPE, FlatModel

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RMD_LIB") or os.path.join(HERE, "librmdetect_b200.so")   # RMD_LIB: A/B runs of two builds
N_CODES = 16
FAST_WINDOW = 160


class NativeError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [("chain_len", C.c_int32 * 2), ("symmetric", C.c_int32),
                ("n_tree", C.c_int32), ("n_leaves", C.c_int32), ("n_cpt_rows", C.c_int32),
                ("n_gate_cfg", C.c_int32), ("n_gate_pairs", C.c_int32), ("n_first", C.c_int32),
                ("brute_candidates", C.c_int32),
                ("tree", C.c_void_p), ("leaf_dist", C.c_void_p), ("cpt", C.c_void_p),
                ("gate_pairs", C.c_void_p), ("gate_cfg", C.c_void_p), ("first_pairs", C.c_void_p),
                ("leaf_offsets", C.c_void_p)]


class ScanInput(C.Structure):
    _fields_ = [("n_seq", C.c_int32), ("seq_off", C.c_void_p), ("codes", C.c_void_p),
                ("gc_log", C.c_void_p), ("gc_class", C.c_void_p),
                ("win_len", C.c_int32), ("win_step", C.c_int32), ("n_win", C.c_int64),
                ("bpp_off", C.c_void_p), ("bpp_i", C.c_void_p), ("bpp_j", C.c_void_p), ("bpp_p", C.c_void_p),
                ("min_bpp_mean", C.c_double), ("cutoff", C.c_double),
                ("bpp_ij", C.c_void_p), ("bpp_q", C.c_void_p)]


class ScanResult(C.Structure):
    _fields_ = [("n_hits", C.c_int64), ("hits", C.c_void_p), ("n_raw", C.c_int64),
                ("tries", C.c_int64 * 2), ("gate_pass", C.c_void_p), ("raw_count", C.c_void_p),
                ("ms_scan", C.c_float), ("ms_post", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
                ("ms_total", C.c_float), ("n_launches", C.c_int32), ("hit_bytes", C.c_int32)]


HIT_DTYPE = np.dtype([("window", "<i8"), ("start", "<i8"), ("seq", "<i4"), ("wlen", "<i4"),
                      ("p0", "<i4"), ("p1", "<i4"), ("model", "<i2"), ("leaf", "<i2"), ("keep", "<i4"),
                      ("prob_model", "<f8"), ("prob_rand", "<f8"), ("score", "<f8"), ("evalue", "<f8")])
assert HIT_DTYPE.itemsize == 72
HIT32_DTYPE = np.dtype([("window", "<u4"), ("seq", "<u4"), ("p0", "<u2"), ("p1", "<u2"), ("model", "u1"), ("keep", "u1"),
                        ("leaf", "<u2"), ("score", "<f8"), ("evalue", "<f8")])        # rmd_hit32: the download form
assert HIT32_DTYPE.itemsize == 32
ROW_DTYPE = np.dtype([("start", "<i8"), ("key", "<i4"), ("wlen", "<u2"), ("p0", "<u2"), ("p1", "<u2"), ("leaf", "<u2"),
                      ("model", "u1"), ("flags", "u1"), ("pad_", "<u2"), ("score", "<f8"), ("evalue", "<f8"), ("bpp", "<f8"),
                      ("letters", "<u4", (3,)), ("id", "<u4")])                        # rmd_hit_row: the device-resident hit table
assert ROW_DTYPE.itemsize == 64
CL_MAX_PAIRS = 16

EXPORTS = ["rmd_create", "rmd_destroy", "rmd_last_error", "rmd_version", "rmd_set_models", "rmd_set_evalues",
           "rmd_upload", "rmd_gc_counts", "rmd_set_gc", "rmd_set_return_dropped", "rmd_set_min_score", "rmd_set_hit_format", "rmd_set_screen_policy", "rmd_scan_resident", "rmd_scan", "rmd_eval",
           "rmd_calibrate", "rmd_calibrate_synth", "rmd_calib_open", "rmd_calib_add", "rmd_calib_device_hist", "rmd_calib_read", "rmd_calib_read_rows", "rmd_cpt_counts", "rmd_synth_job", "rmd_synth_strand", "rmd_synth_counts", "rmd_download_codes", "rmd_download_bpp",
           "rmd_download_job", "rmd_compact_bpp", "rmd_parse_dot_ps", "rmd_host_alloc", "rmd_host_free",
           "rmd_table_clear", "rmd_table_rows", "rmd_table_append_scan", "rmd_table_append_rows", "rmd_table_download",
           "rmd_table_filter", "rmd_table_sort", "rmd_table_is_sorted", "rmd_table_top", "rmd_table_nooverlap", "rmd_table_cluster_index",
           "rmd_table_cluster_info", "rmd_table_cluster_metrics"]

_lib = None


def load_library():
    """dlopen the CUDA library and declare signatures.  Raises NativeError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError("%s is missing: build it with `make -C rmdetect_b200/csrc` "
                          "(python -c 'import __graft_entry__ as g; g.build()')" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.rmd_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.rmd_destroy.argtypes = [vp]
    L.rmd_destroy.restype = None
    L.rmd_last_error.argtypes = [vp]
    L.rmd_last_error.restype = C.c_char_p
    L.rmd_version.restype = C.c_char_p
    L.rmd_set_models.argtypes = [vp, C.POINTER(ModelDesc), C.c_int]
    L.rmd_set_evalues.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp]
    L.rmd_upload.argtypes = [vp, C.POINTER(ScanInput)]
    L.rmd_gc_counts.argtypes = [vp, vp]
    L.rmd_set_gc.argtypes = [vp, vp, vp]
    L.rmd_set_return_dropped.argtypes = [vp, C.c_int]
    L.rmd_set_min_score.argtypes = [vp, C.c_int, f64]
    L.rmd_set_hit_format.argtypes = [vp, C.c_int]
    L.rmd_set_screen_policy.argtypes = [vp, i64]
    L.rmd_scan_resident.argtypes = [vp, i64, i64, C.POINTER(ScanResult)]
    L.rmd_scan.argtypes = [vp, C.POINTER(ScanInput), C.POINTER(ScanResult)]
    L.rmd_eval.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, i64, f64, vp, vp, vp]
    L.rmd_calibrate.argtypes = [vp, C.c_int, vp, i64, i32, vp, i32, f64, vp, i32,
                                C.POINTER(i64), C.POINTER(i32), C.POINTER(C.c_float)]
    L.rmd_calibrate_synth.argtypes = [vp, C.c_int, C.c_uint64, i64, i64, i32, vp, i32, f64, vp, i32,
                                      C.POINTER(i64), C.POINTER(i32), C.POINTER(C.c_float)]
    L.rmd_calib_open.argtypes = [vp, C.c_int, i32, vp, i32, f64, i32]
    L.rmd_calib_add.argtypes = [vp, vp, C.c_uint64, i64, i64]
    L.rmd_calib_device_hist.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.rmd_calib_read.argtypes = [vp, vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(C.c_float)]
    L.rmd_calib_read_rows.argtypes = [vp, i32, i32, vp, C.POINTER(vp), C.POINTER(i32)]
    L.rmd_cpt_counts.argtypes = [vp, vp, i64, i32, i32, vp, i32, vp, i32, vp, vp]
    L.rmd_synth_job.argtypes = [vp, i64, i64, C.c_uint64, C.c_uint64, i32, i32, f64, f64, vp, vp,
                                C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_float)]
    L.rmd_synth_strand.argtypes = [vp, i64, i32]
    L.rmd_synth_counts.argtypes = [vp, i64, i64, C.c_uint64, vp]
    L.rmd_download_codes.argtypes = [vp, vp, i64]
    L.rmd_download_bpp.argtypes = [vp, i64, vp, vp, vp, i64, C.POINTER(i64)]
    L.rmd_download_job.argtypes = [vp, vp, vp, vp, vp]
    L.rmd_compact_bpp.argtypes = [vp, vp, vp, i64, vp, vp, C.POINTER(i64)]
    L.rmd_parse_dot_ps.argtypes = [C.c_char_p, i64, vp, vp, vp, vp, i64, C.POINTER(i64), C.c_char_p, i64]
    pi64, pf64 = C.POINTER(i64), C.POINTER(f64)
    L.rmd_table_clear.argtypes = [vp]
    L.rmd_table_rows.argtypes = [vp, pi64]
    L.rmd_table_append_scan.argtypes = [vp, i32, i64, i64, pi64]
    L.rmd_table_append_rows.argtypes = [vp, vp, i64, pi64]
    L.rmd_table_download.argtypes = [vp, vp, i64, i64]
    L.rmd_table_filter.argtypes = [vp, pf64, pf64, pf64, i32, pi64]
    L.rmd_table_sort.argtypes = [vp, vp, i32]
    L.rmd_table_is_sorted.argtypes = [vp, vp, i32, pi64]
    L.rmd_table_top.argtypes = [vp, i32, i32, vp, i32, pi64]
    L.rmd_table_nooverlap.argtypes = [vp, pi64]
    L.rmd_table_cluster_index.argtypes = [vp, vp, vp, i32, vp, pi64]
    L.rmd_table_cluster_info.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rmd_table_cluster_metrics.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.rmd_host_alloc.argtypes = [i64]
    L.rmd_host_alloc.restype = vp
    L.rmd_host_free.argtypes = [vp]
    L.rmd_host_free.restype = None
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One GPU context: models, E-value tables, a resident job, scans."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.rmd_create(device, C.byref(h))
        if rc != 0:
            raise NativeError("rmd_create(%d): %s" % (device, self.lib.rmd_last_error(None).decode()))
        self.h = h
        self.device = device
        self.n_models = 0
        self._keep = []          # numpy arrays the library only needs during a call, kept for safety

    def close(self):
        if getattr(self, "h", None):
            self.lib.rmd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise NativeError("%s failed (%d): %s" % (what, rc, self.lib.rmd_last_error(self.h).decode()))

    # -- models ------------------------------------------------------------------------------
    def set_models(self, flats):
        descs = (ModelDesc * len(flats))()
        keep = []
        for d, fm in zip(descs, flats):
            assert isinstance(fm, FlatModel)
            tree = np.ascontiguousarray(fm.tree, dtype=TREE_DTYPE)
            ld = np.ascontiguousarray(fm.leaf_dist, dtype=np.int32)
            cpt = np.ascontiguousarray(fm.cpt, dtype=np.float64)
            gp = np.ascontiguousarray(fm.gate_pairs, dtype=PAIR_DTYPE)
            gcfg = np.ascontiguousarray(fm.gate_cfg, dtype=np.int32)
            fp = np.ascontiguousarray(fm.first_pairs, dtype=np.int8)
            L = fm.chain_len[0] + fm.chain_len[1]
            lo = np.full((fm.n_leaves, L), -1, dtype=np.int8)
            for li, cfg in enumerate(fm.leaf_offsets):
                flat = list(cfg[0]) + list(cfg[1])
                lo[li] = [-1 if o is None else o for o in flat]
            keep += [tree, ld, cpt, gp, gcfg, fp, lo]
            d.chain_len[0], d.chain_len[1] = fm.chain_len
            d.symmetric = int(fm.symmetric)
            d.n_tree, d.n_leaves, d.n_cpt_rows = len(tree), fm.n_leaves, cpt.shape[0]
            d.n_gate_cfg, d.n_gate_pairs, d.n_first = len(gcfg) - 1, len(gp), len(fp)
            d.brute_candidates = int(fm.brute_candidates)
            d.tree, d.leaf_dist, d.cpt = tree.ctypes.data, ld.ctypes.data, cpt.ctypes.data
            d.gate_pairs, d.gate_cfg, d.first_pairs = gp.ctypes.data, gcfg.ctypes.data, fp.ctypes.data
            d.leaf_offsets = lo.ctypes.data
        self._check(self.lib.rmd_set_models(self.h, descs, len(flats)), "rmd_set_models")
        self.n_models = len(flats)
        del keep

    def set_evalues(self, model, table, thresholds):
        if table is None or table.size == 0:
            self._check(self.lib.rmd_set_evalues(self.h, model, None, 0, 0, None), "rmd_set_evalues")
            return
        table = np.ascontiguousarray(table, dtype=np.float64)
        thr = np.ascontiguousarray(thresholds, dtype=np.float64)
        assert thr.shape[0] == table.shape[1]
        self._check(self.lib.rmd_set_evalues(self.h, model, table.ctypes.data, table.shape[0], table.shape[1],
                                             thr.ctypes.data), "rmd_set_evalues")

    # -- jobs --------------------------------------------------------------------------------
    @staticmethod
    def _scan_input(job):
        si = ScanInput()
        si.n_seq = len(job["seq_off"]) - 1
        si.seq_off, si.codes = _ptr(job["seq_off"]), _ptr(job["codes"])
        si.gc_log, si.gc_class = _ptr(job["gc_log"]), _ptr(job.get("gc_class"))
        si.win_len, si.win_step, si.n_win = job["win_len"], job["win_step"], job["n_win"]
        si.bpp_off, si.bpp_i, si.bpp_j, si.bpp_p = (_ptr(job["bpp_off"]), _ptr(job["bpp_i"]), _ptr(job["bpp_j"]),
                                                    _ptr(job["bpp_p"]))
        si.min_bpp_mean, si.cutoff = job["min_bpp_mean"], job["cutoff"]
        si.bpp_ij, si.bpp_q = _ptr(job.get("bpp_ij")), _ptr(job.get("bpp_q"))       # compact transport (optional)
        return si

    def upload(self, job):
        si = self._scan_input(job)
        self._check(self.lib.rmd_upload(self.h, C.byref(si)), "rmd_upload")

    def _result(self, res, n_win, copy):
        """Wrap the library's result.  The arrays are views of page-locked buffers the context owns and are
        overwritten by its next scan call; `copy=True` detaches them."""
        def view(ptr, count, dtype):
            if count == 0 or not ptr:
                return np.empty(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            arr = np.frombuffer(buf, dtype=dtype, count=count)
            return arr.copy() if copy else arr
        ng = n_win * self.n_models
        hits = view(res.hits, res.n_hits, HIT32_DTYPE if res.hit_bytes == 32 else HIT_DTYPE)
        gate = view(res.gate_pass, ng, np.uint32).reshape(n_win, self.n_models)
        raw = view(res.raw_count, ng, np.uint32).reshape(n_win, self.n_models)
        return dict(hits=hits, tries=[res.tries[0], res.tries[1]], gate_pass=gate, raw_count=raw, n_raw=res.n_raw,
                    ms_scan=res.ms_scan, ms_post=res.ms_post, ms_h2d=res.ms_h2d, ms_d2h=res.ms_d2h,
                    ms_total=res.ms_total, n_launches=res.n_launches)

    def scan_resident(self, w_begin, w_end, copy=True):
        res = ScanResult()
        self._check(self.lib.rmd_scan_resident(self.h, w_begin, w_end, C.byref(res)), "rmd_scan_resident")
        return self._result(res, w_end - w_begin, copy)

    def scan(self, job, copy=True):
        si = self._scan_input(job)
        res = ScanResult()
        self._check(self.lib.rmd_scan(self.h, C.byref(si), C.byref(res)), "rmd_scan")
        return self._result(res, job["n_win"], copy)

    def gc_counts(self, n_seq):
        out = np.zeros((n_seq, N_CODES), dtype=np.int64)
        self._check(self.lib.rmd_gc_counts(self.h, out.ctypes.data), "rmd_gc_counts")
        return out

    def set_gc(self, gc_log, gc_class=None):
        gc_log = np.ascontiguousarray(gc_log, dtype=np.float64)
        gcc = None if gc_class is None else np.ascontiguousarray(gc_class, dtype=np.int32)
        self._check(self.lib.rmd_set_gc(self.h, gc_log.ctypes.data, _ptr(gcc)), "rmd_set_gc")

    def eval(self, model, codes, seq_off, gc_log, p0, p1, cutoff):
        n = len(p0)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        gc_log = np.ascontiguousarray(gc_log, dtype=np.float64)
        p0 = np.ascontiguousarray(p0, dtype=np.int32)
        p1 = np.ascontiguousarray(p1, dtype=np.int32)
        leaf = np.zeros(n, dtype=np.int32)
        pm = np.zeros(n, dtype=np.float64)
        pr = np.zeros(n, dtype=np.float64)
        self._check(self.lib.rmd_eval(self.h, model, codes.ctypes.data, seq_off.ctypes.data, gc_log.ctypes.data,
                                      p0.ctypes.data, p1.ctypes.data, n, cutoff, leaf.ctypes.data, pm.ctypes.data,
                                      pr.ctypes.data), "rmd_eval")
        return leaf, pm, pr

    def calibrate(self, model, samples, class_log, cutoff, hist, seed=None, first=0, n=None, L=None):
        class_log = np.ascontiguousarray(class_log, dtype=np.float64)
        assert hist.dtype == np.float64 and hist.flags.c_contiguous and hist.shape[0] == class_log.shape[0]
        n_ok, ovf, ms = C.c_int64(0), C.c_int32(0), C.c_float(0)
        if samples is not None:
            samples = np.ascontiguousarray(samples, dtype=np.uint8)
            n, L = samples.shape
            rc = self.lib.rmd_calibrate(self.h, model, samples.ctypes.data, n, L, class_log.ctypes.data,
                                        class_log.shape[0], cutoff, hist.ctypes.data, hist.shape[1],
                                        C.byref(n_ok), C.byref(ovf), C.byref(ms))
        else:
            rc = self.lib.rmd_calibrate_synth(self.h, model, seed, first, n, L, class_log.ctypes.data,
                                              class_log.shape[0], cutoff, hist.ctypes.data, hist.shape[1],
                                              C.byref(n_ok), C.byref(ovf), C.byref(ms))
        self._check(rc, "rmd_calibrate")
        if ovf.value:
            raise NativeError("calibration score beyond the histogram range (%d bins)" % hist.shape[1])
        return n_ok.value, ms.value

    # calibration session: histograms resident on the device (include/rmdetect_b200.h: rmd_calib_*)
    def calib_open(self, model, L, class_log, cutoff, n_bins):
        class_log = np.ascontiguousarray(class_log, dtype=np.float64)
        self._check(self.lib.rmd_calib_open(self.h, model, L, class_log.ctypes.data, class_log.shape[0], cutoff, n_bins),
                    "rmd_calib_open")
        self._cal_shape = (class_log.shape[0], n_bins)

    def calib_add(self, samples=None, seed=0, first=0, n=0):
        if samples is not None:
            samples = np.ascontiguousarray(samples, dtype=np.uint8)
            n = samples.shape[0]
        self._check(self.lib.rmd_calib_add(self.h, _ptr(samples), seed, first, n), "rmd_calib_add")

    def calib_device_hist(self):
        """(device address, element count) of the session's fp64 histogram, after the enqueued batches have finished."""
        p, n = C.c_void_p(), C.c_int64(0)
        self._check(self.lib.rmd_calib_device_hist(self.h, C.byref(p), C.byref(n)), "rmd_calib_device_hist")
        return p.value, n.value

    def calib_read(self, want_hist=True, out=None):
        """(histogram, n_ok, device ms).  `out`: a float64 array of the session's shape that receives the histogram (a caller
        that reads after every block keeps one buffer instead of mapping 676 KB of fresh pages per call)."""
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == self._cal_shape
        hist = (out if out is not None else np.zeros(self._cal_shape, dtype=np.float64)) if want_hist else None
        n_ok, ovf, ms = C.c_int64(0), C.c_int32(0), C.c_float(0)
        self._check(self.lib.rmd_calib_read(self.h, _ptr(hist), C.byref(n_ok), C.byref(ovf), C.byref(ms)), "rmd_calib_read")
        if ovf.value:
            raise NativeError("calibration score beyond the histogram range (%d bins)" % self._cal_shape[1])
        return hist, n_ok.value, ms.value

    def calib_read_rows(self, first_class, n_rows, want_rows=True):
        """(rows [n_rows][n_bins] of the session's histogram or None, device address of the first row, (first, last) bucket
        occupied in any class) after the enqueued batches."""
        rows = np.empty((n_rows, self._cal_shape[1]), dtype=np.float64) if want_rows else None
        p, cols = C.c_void_p(), (C.c_int32 * 2)()
        self._check(self.lib.rmd_calib_read_rows(self.h, first_class, n_rows, _ptr(rows), C.byref(p), cols), "rmd_calib_read_rows")
        return rows, p.value, (cols[0], cols[1])

    def cpt_counts(self, letters, K, tables, run, n_runs):
        """(counts [n_runs][n_tables][K^3] uint32, first_row [n_tables][K^3] int32) — include/rmdetect_b200.h: rmd_cpt_counts."""
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        tables = np.ascontiguousarray(tables, dtype=np.int32).reshape(-1, 3)
        run = np.ascontiguousarray(run, dtype=np.int32)
        k3 = K ** 3
        counts = np.zeros((n_runs, len(tables), k3), dtype=np.uint32)
        first = np.zeros((len(tables), k3), dtype=np.int32)
        self._check(self.lib.rmd_cpt_counts(self.h, letters.ctypes.data, letters.shape[0], letters.shape[1], K, tables.ctypes.data,
                                            len(tables), run.ctypes.data, n_runs, counts.ctypes.data, first.ctypes.data), "rmd_cpt_counts")
        return counts, first

    def synth_job(self, n, first, seq_seed, bpp_seed, win_len, win_step, min_bpp_mean, cutoff, gc_log, gc_class=None):
        gc_log = np.ascontiguousarray(gc_log, dtype=np.float64).reshape(1, N_CODES)
        gcc = None if gc_class is None else np.ascontiguousarray(gc_class, dtype=np.int32)
        n_win, n_bpp, ms = C.c_int64(0), C.c_int64(0), C.c_float(0)
        self._check(self.lib.rmd_synth_job(self.h, n, first, seq_seed, bpp_seed, win_len, win_step, min_bpp_mean,
                                           cutoff, gc_log.ctypes.data, _ptr(gcc), C.byref(n_win), C.byref(n_bpp),
                                           C.byref(ms)), "rmd_synth_job")
        return n_win.value, n_bpp.value, ms.value

    def set_screen_policy(self, min_windows):
        """Sequences with at least `min_windows` windows get prefix-screen tables (no effect on results); < 0: none."""
        self._check(self.lib.rmd_set_screen_policy(self.h, int(min_windows)), "rmd_set_screen_policy")

    def set_return_dropped(self, enable):
        """Scans also return the hits the duplicate rule removed (keep = 0), in place, as the reference's per-window
        callbacks see them (lib/scanner.py:99-109 runs before :53)."""
        self._check(self.lib.rmd_set_return_dropped(self.h, 1 if enable else 0), "rmd_set_return_dropped")

    def set_hit_format(self, compact):
        """compact=True: scans return HIT32_DTYPE records (window, seq, p0, p1, model, keep, leaf, score, evalue: 32
        bytes instead of 72 on the way back); `expand_hits` rebuilds start / wlen from the job's geometry."""
        self._check(self.lib.rmd_set_hit_format(self.h, 1 if compact else 0), "rmd_set_hit_format")

    def set_min_score(self, min_score):
        """Drop hits with score <= min_score on the device (reference rmdetect.py:44); None switches the filter off."""
        self._check(self.lib.rmd_set_min_score(self.h, 0 if min_score is None else 1, 0.0 if min_score is None else float(min_score)),
                    "rmd_set_min_score")

    def synth_strand(self, total, reverse):
        """Strand of the synthetic stream that synth_job / synth_counts read from now on (reference
        lib/sequences.py:138-143 for the reverse one)."""
        self._check(self.lib.rmd_synth_strand(self.h, total, 1 if reverse else 0), "rmd_synth_strand")

    def synth_counts(self, n, first, seq_seed):
        out = np.zeros(4, dtype=np.int64)
        self._check(self.lib.rmd_synth_counts(self.h, n, first, seq_seed, out.ctypes.data), "rmd_synth_counts")
        return out

    def download_codes(self, n):
        out = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.rmd_download_codes(self.h, out.ctypes.data, n), "rmd_download_codes")
        return out

    def download_bpp(self, window, cap=16384):
        i = np.zeros(cap, dtype=np.uint16)
        j = np.zeros(cap, dtype=np.uint16)
        p = np.zeros(cap, dtype=np.float64)
        n = C.c_int64(0)
        self._check(self.lib.rmd_download_bpp(self.h, window, i.ctypes.data, j.ctypes.data, p.ctypes.data, cap,
                                              C.byref(n)), "rmd_download_bpp")
        return i[:n.value], j[:n.value], p[:n.value]

    def download_job(self, n_win, n_bpp, pinned=False):
        """(bpp_off, i, j, p) of the resident job as host arrays (page-locked when `pinned`)."""
        off = np.zeros(n_win + 1, dtype=np.int64)
        if pinned:
            i = pinned_array(n_bpp, np.uint16)
            j = pinned_array(n_bpp, np.uint16)
            p = pinned_array(n_bpp, np.float64)
        else:
            i = np.zeros(max(n_bpp, 1), dtype=np.uint16)
            j = np.zeros(max(n_bpp, 1), dtype=np.uint16)
            p = np.zeros(max(n_bpp, 1), dtype=np.float64)
        self._check(self.lib.rmd_download_job(self.h, off.ctypes.data, i.ctypes.data, j.ctypes.data, p.ctypes.data),
                    "rmd_download_job")
        return off, i, j, p


def expand_hits(h32, seq_off, win_len, win_step):
    """HIT_DTYPE records from the 32-byte download form: `start` and `wlen` follow from the window index and the job's
    geometry (reference lib/scanner.py:80-111); prob_model / prob_rand are not carried (NaN)."""
    seq_off = np.asarray(seq_off, dtype=np.int64)
    lens = np.diff(seq_off)
    if win_len == 0 or win_step == 0:
        n_win = np.ones(len(lens), dtype=np.int64)
    else:
        n_win = np.where(lens <= win_len, 1, 1 + -(-(lens - win_len) // max(win_step, 1)))
    base = np.concatenate([[0], np.cumsum(n_win)])
    out = np.zeros(len(h32), dtype=HIT_DTYPE)
    seq = h32["seq"].astype(np.int64)
    k = h32["window"].astype(np.int64) - base[seq]
    if win_len == 0 or win_step == 0:
        start = np.zeros(len(h32), dtype=np.int64)
        wlen = lens[seq]
    else:
        last = k == n_win[seq] - 1
        start = np.where(last, np.maximum(lens[seq] - win_len, 0), k * win_step)
        wlen = np.minimum(lens[seq] - start, win_len)
    out["window"], out["start"], out["seq"], out["wlen"] = h32["window"], start, seq, wlen
    for f in ("p0", "p1", "model", "leaf", "keep", "score", "evalue"):
        out[f] = h32[f]
    out["prob_model"] = out["prob_rand"] = np.nan
    return out


def compact_bpp(i, j, p, pinned=False):
    """(ij uint16[], q uint32[]) — the 6-byte transport form of a coordinate list (include/rmdetect_b200.h:
    rmd_compact_bpp), or None when some entry has no such form."""
    lib = load_library()
    i = np.ascontiguousarray(i, dtype=np.uint16)
    j = np.ascontiguousarray(j, dtype=np.uint16)
    p = np.ascontiguousarray(p, dtype=np.float64)
    n = len(p)
    ij = pinned_array(n, np.uint16) if pinned else np.zeros(max(n, 1), dtype=np.uint16)
    q = pinned_array(n, np.uint32) if pinned else np.zeros(max(n, 1), dtype=np.uint32)
    bad = C.c_int64(-1)
    rc = lib.rmd_compact_bpp(i.ctypes.data, j.ctypes.data, p.ctypes.data, n, ij.ctypes.data, q.ctypes.data, C.byref(bad))
    if rc != 0:
        return None
    return ij[:n] if not pinned else ij, q[:n] if not pinned else q


def parse_dot_ps(text):
    """(sequence, i uint16[], j uint16[], p float64[], q uint32[]) from the text of a ViennaRNA dot.ps, by the
    library's native parser (rmd_parse_dot_ps; replaces lib/rnatools.py:176-205 of the reference)."""
    lib = load_library()
    raw = text if isinstance(text, bytes) else text.encode("ascii", "replace")
    cap = raw.count(b"ubox") + 1
    i = np.zeros(cap, dtype=np.uint16)
    j = np.zeros(cap, dtype=np.uint16)
    p = np.zeros(cap, dtype=np.float64)
    q = np.zeros(cap, dtype=np.uint32)
    n = C.c_int64(0)
    seq = C.create_string_buffer(len(raw) + 1)
    rc = lib.rmd_parse_dot_ps(raw, len(raw), i.ctypes.data, j.ctypes.data, p.ctypes.data, q.ctypes.data, cap,
                              C.byref(n), seq, len(raw) + 1)
    if rc != 0:
        raise NativeError("rmd_parse_dot_ps failed (%d)" % rc)
    k = n.value
    return seq.value.decode("ascii") or None, i[:k], j[:k], p[:k], q[:k]


class _Pinned:
    def __init__(self, nbytes):
        self.lib = load_library()
        self.ptr = self.lib.rmd_host_alloc(nbytes)
        if not self.ptr:
            raise NativeError("rmd_host_alloc(%d) failed" % nbytes)

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.rmd_host_free(self.ptr)
            self.ptr = None


class _PinnedArray(np.ndarray):
    """ndarray view that keeps its page-locked allocation alive."""
    _owner = None


def pinned_array(n, dtype):
    """numpy array over page-locked host memory (released when the array is collected)."""
    dtype = np.dtype(dtype)
    n = max(int(n), 1)
    owner = _Pinned(n * dtype.itemsize)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(owner.ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n).view(_PinnedArray)
    arr._owner = owner
    return arr
