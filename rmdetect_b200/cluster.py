"""rmcluster's algorithm on a device-resident hit table (reference lib/cluster.py).

`ClusterAlgorithm.go_cluster` of the reference (lib/cluster.py:185-240) does three things:

1. the first cluster list: one cluster per distinct (model, first column of chain 0, first column of chain 1), in
   the order the hit list meets those keys (:185-203) — done on the device, `rmd_table_cluster_index`;
2. merging: clusters of one model whose rows and columns differ by at most `dlimit` are merged pairwise, in list
   order, until a pass merges nothing (:205-226, `Cluster.merge` :38-56) — order-dependent bookkeeping over cluster
   SUMMARIES (a few per alignment column), replayed here on the host statement by statement, including the
   "most representative position" update that, with every count equal to 1, simply takes the last position in dict
   order;
3. the metrics of every cluster (`Cluster.calc_metrics`, :58-83): count, occurrence, mean score and bpp, mutual
   information and entropy of the model's base pairs (:98-166).  The sums over member hits and the letter-pair counts
   come from the device (`rmd_table_cluster_metrics`, in the member order the reference's lists have); the handful
   of logarithms per cluster are taken here with the reference's own expressions.

The reference evaluates its `filter_code` but returns True whatever it says (:81-83), so every cluster is retained;
`apply_filter=True` makes the thresholds effective.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._native import CL_MAX_PAIRS, _ptr

CANONICAL = ("AU", "CG", "GC", "GU", "UA", "UG")
_IDX = {"A": 0, "C": 1, "G": 2, "U": 3}


class Cluster:
    def __init__(self, model_name, model_index, row, col, members):
        self.model_name, self.model_index = model_name, model_index
        self.row, self.col = row, col
        self.coords = {(row, col): 1}
        self.members = members            # table rows, in the order of the reference's cluster.cands
        self.count, self.occur, self.score, self.bpp, self.mi, self.h = 0, 0.0, 0.0, 0.0, 0.0, 0.0

    def merge(self, other, dlimit):
        # lib/cluster.py:38-56
        if self.model_name == other.model_name and abs(self.row - other.row) <= dlimit and abs(self.col - other.col) <= dlimit:
            self.members = np.concatenate([self.members, other.members])
            for k, v in other.coords.items():
                self.coords[k] = self.coords.get(k, 0) + v
            max_v = max(self.coords.values())
            for (row, col), v in self.coords.items():
                if v == max_v:
                    self.row, self.col = row, col
            return True
        return False

    def __str__(self):
        # lib/cluster.py:85-86
        return ("model: %s count: %6d occur_(%%): %6.2f score: %7.3f bpp: %5.3f MI: %7.3f H: %7.3f cols: %4d %4d"
                % (self.model_name.ljust(10), self.count, self.occur * 100.0, self.score, self.bpp, self.mi, self.h, self.row, self.col))


def _mi(counts, pair_count):
    """lib/cluster.py:98-135 from the letter-pair counts [pair][x][y]."""
    total = 0.0
    for i in range(pair_count):
        xy = {a + b: float(counts[i][_IDX[a] * 4 + _IDX[b]]) for a, b in CANONICAL}
        n = 0.0
        dx, dy = {}, {}
        for k, v in xy.items():
            if v > 0.0:
                n += v
                dx[k[0]] = dx.get(k[0], 0.0) + v
                dy[k[1]] = dy.get(k[1], 0.0) + v
        mi = 0.0
        for x in "ACGU":
            for y in "ACGU":
                f = xy.get(x + y, 0.0)
                if f > 0.0 and n > 0.0:
                    p_xy, p_x, p_y = f / n, dx.get(x, 0.0) / n, dy.get(y, 0.0) / n
                    mi += p_xy * math.log(p_xy / (p_x * p_y), 2)
        total += mi
    return total / (2.0 * float(pair_count))


def _h(counts, pair_count, n_seqs):
    """lib/cluster.py:137-166."""
    max_h = math.log(6.0, 2)
    total = 0.0
    for i in range(pair_count):
        xy = {a + b: float(counts[i][_IDX[a] * 4 + _IDX[b]]) for a, b in CANONICAL}
        n = 0.0
        for v in xy.values():
            n += v
        h = 0.0
        for k in CANONICAL:
            f = xy[k]
            if f > 0.0 and n > 0.0:
                p = f / n
                h += -p * math.log(p, 2)
        h = h * (n / float(n_seqs))
        total += h
    return total / (max_h * float(pair_count))


class ClusterAlgorithm:
    def __init__(self, table, seq_count, min_count=3, min_occur=0.05, min_score=0.0, min_bpp=0.0, min_mi=0.0, min_h=0.0,
                 apply_filter=False):
        self.table, self.seq_count = table, seq_count
        self.thresholds = (min_count, min_occur, min_score, min_bpp, min_mi, min_h)
        self.apply_filter = apply_filter
        self.clusters = []

    def _index(self, cols_by_key):
        T = self.table
        n = len(T)
        cluster_of = np.zeros(max(n, 1), dtype=np.int32)
        nc = C.c_int64(0)
        if cols_by_key is None:
            rc = T.lib.rmd_table_cluster_index(T.ctx.h, None, None, 0, _ptr(cluster_of), C.byref(nc))
        else:
            parts = [np.ascontiguousarray(cols_by_key[k], dtype=np.int32) for k in T.keys]
            off = np.zeros(len(parts) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(p) for p in parts])
            cols = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(1, np.int32), dtype=np.int32)
            rc = T.lib.rmd_table_cluster_index(T.ctx.h, _ptr(off), _ptr(cols), len(parts), _ptr(cluster_of), C.byref(nc))
        T.ctx._check(rc, "rmd_table_cluster_index")
        k = nc.value
        row, col = np.zeros(max(k, 1), np.int64), np.zeros(max(k, 1), np.int64)
        model, count = np.zeros(max(k, 1), np.int32), np.zeros(max(k, 1), np.int32)
        T.ctx._check(T.lib.rmd_table_cluster_info(T.ctx.h, _ptr(row), _ptr(col), _ptr(model), _ptr(count), None), "rmd_table_cluster_info")
        return cluster_of[:n], row[:k], col[:k], model[:k], count[:k]

    def go_cluster(self, dlimit, cols_by_key=None):
        """cols_by_key: {key: alignment column index} when the hits' sequences came from an alignment
        (lib/candidates.py:489-506: a hit's columns are col_index[position]); None: columns are positions."""
        T = self.table
        models = T.engine.models
        cluster_of, row, col, model, count = self._index(cols_by_key)
        order = np.argsort(cluster_of, kind="stable")             # members of a cluster together, in table order
        ends = np.cumsum(count)
        self.clusters = [Cluster(models[model[c]].full_name(), int(model[c]), int(row[c]), int(col[c]),
                                 order[ends[c] - count[c]:ends[c]].astype(np.int32)) for c in range(len(count))]
        merge = True
        while merge:                                              # lib/cluster.py:205-226
            merge = False
            cl = self.clusters
            for i in range(len(cl)):
                if cl[i] is not None:
                    for j in range(i + 1, len(cl)):
                        if cl[j] is not None and cl[i].merge(cl[j], dlimit):
                            merge = True
                            cl[j] = None
            self.clusters = [c for c in cl if c is not None]
        self._metrics()
        if self.apply_filter:
            mc, mo, ms, mb, mm, mh = self.thresholds
            self.clusters = [c for c in self.clusters if c.count >= mc and c.occur >= mo and c.score >= ms and c.bpp >= mb
                             and c.mi >= mm and c.h >= mh]
        return self.clusters

    def _metrics(self):
        T = self.table
        models = T.engine.models
        k = len(self.clusters)
        if k == 0:
            return
        members = np.ascontiguousarray(np.concatenate([c.members for c in self.clusters]), dtype=np.int32)
        off = np.zeros(k + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(c.members) for c in self.clusters])
        pair_ids = np.zeros((len(models), CL_MAX_PAIRS, 2), dtype=np.uint8)
        n_pairs = np.zeros(len(models), dtype=np.int32)
        for m, mod in enumerate(models):
            if len(mod.pairing) > CL_MAX_PAIRS:
                raise ValueError("model %s has more than %d PAIRS lines" % (mod.full_name(), CL_MAX_PAIRS))
            n_pairs[m] = len(mod.pairing)
            for q, pr in enumerate(mod.pairing):
                pair_ids[m, q] = (pr.id1, pr.id2)
        sums = np.zeros((k, 2), dtype=np.float64)
        counts = np.zeros((k, CL_MAX_PAIRS, 16), dtype=np.uint32)
        n_keys = np.zeros(k, dtype=np.uint32)
        T.ctx._check(T.lib.rmd_table_cluster_metrics(T.ctx.h, _ptr(members), _ptr(off), k, _ptr(pair_ids), _ptr(n_pairs),
                                                     _ptr(sums), _ptr(counts), _ptr(n_keys)), "rmd_table_cluster_metrics")
        for c, cl in enumerate(self.clusters):                    # lib/cluster.py:58-83
            cl.count = len(cl.members)
            cl.occur = float(int(n_keys[c])) / float(self.seq_count)
            cl.score = float(sums[c, 0]) / float(cl.count)
            cl.bpp = float(sums[c, 1]) / float(cl.count)
            np_ = int(n_pairs[cl.model_index])
            cl.mi = _mi(counts[c], np_)
            cl.h = _h(counts[c], np_, cl.count)

    def member_rows(self, cluster):
        """The cluster's hits as table rows, in the reference's `cluster.cands` order."""
        rows = self.table.rows()
        return rows[cluster.members]
