"""The device-resident hit table (rmd_table_*) against the reference's own list operations.

* rmfilter's selections on the 120 shipped lysine hits: the rows left on the device equal what the reference's
  `Candidates.top_model / top_seq / top_seq_model / get_model / nooverlap` return (tests/golden/filters.json.gz);
* rmcluster on the same hits for several column limits and list orders: clusters, members in order, metrics and the
  printed summary equal `lib/cluster.py`'s (tests/golden/clusters.json.gz, written by tests/golden/make_clusters.py);
* tables of 3 000 and 200 000 synthetic rows (ties, dense overlaps) against the host restatement in rmdetect_b200/candidates.py,
  which test_filters.py pins on the reference;
* rows appended straight from a scan equal the Python `Candidate` objects of the same scan.
"""
import math
import os

import numpy as np
import pytest

from helpers import LN_UNKNOWN, MODEL_ORDER, REF_DATA, Cfg, SeqList, fresh_models, golden, load_evalues, load_model

pytestmark = pytest.mark.gpu


def _engine():
    from rmdetect_b200.scanner import ScanEngine
    models = [load_model(m) for m in MODEL_ORDER]
    return ScanEngine(models, [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)


def _shipped_hits(name="filters"):
    from rmdetect_b200 import candidates as C
    models = [load_model(m) for m in MODEL_ORDER]
    g = golden(name)
    by_file, out = {}, []
    for fname, line in g["input"]:
        if fname not in by_file:
            with open(os.path.join(REF_DATA, fname)) as fh:
                by_file[fname] = fh.read().split("\n")
        out.append(C.decode_candidate(by_file[fname][line], models))
    return g, out


def test_selections_on_the_device_equal_the_reference():
    from rmdetect_b200.hittable import HitTable
    g, cands = _shipped_hits()
    eng = _engine()
    T = HitTable(eng)
    ident = g["input"]

    def run(op, order=None):
        T.clear()
        T.append_candidates(cands if order is None else [cands[k] for k in order])
        op(T)
        ids = T.rows()["id"]
        return [ident[k if order is None else order[k]] for k in ids]

    assert run(lambda t: t.top_model(3)) == g["top_model_3"]
    assert run(lambda t: t.top_seq(1)) == g["top_seq_1"]
    assert run(lambda t: t.top_seq_model(2)) == g["top_seq_model_2"]
    assert run(lambda t: t.get_model("kt_1.0")) == g["get_model_kt"]
    assert run(lambda t: t.nooverlap()) == g["nooverlap"]
    order = sorted(range(len(cands)), key=lambda k: cands[k]._sort_key())       # stable, as list.sort
    assert run(lambda t: t.nooverlap(), order) == g["nooverlap_sorted"]
    assert run(lambda t: (t.sort(), t.nooverlap())) == g["nooverlap_sorted"]
    # rows back to Candidate objects: the same .res lines
    T.clear()
    T.append_candidates(cands)
    back = T.candidates()
    for a, b in zip(cands, back):
        b.chains_cols = a.chains_cols                     # columns are the alignment's, not kept in a row
        assert str(a) == str(b)
    eng.close()


def test_clusters_equal_the_reference():
    from rmdetect_b200.cluster import ClusterAlgorithm
    from rmdetect_b200.hittable import HitTable
    g, cands = _shipped_hits("clusters")
    eng = _engine()
    T = HitTable(eng)
    ident = g["input"]
    # the alignment's column index, as far as these hits show it: column of every first position
    cols = {}
    for c in cands:
        arr = cols.setdefault(c.key, np.arange(4096, dtype=np.int32))
        for ch_pos, ch_col in zip(c.chains_pos, c.chains_cols):
            arr[ch_pos[0]] = ch_col[0]
    for run in g["runs"]:
        order = {"file": list(range(len(cands))), "reversed": list(range(len(cands)))[::-1],
                 "sorted": sorted(range(len(cands)), key=lambda k: cands[k]._sort_key())}[run["order"]]
        T.clear()
        T.append_candidates([cands[k] for k in order])
        algo = ClusterAlgorithm(T, g["seq_count"])
        clusters = algo.go_cluster(run["dlimit"], cols_by_key=cols)
        assert len(clusters) == len(run["clusters"]), (run["order"], run["dlimit"])
        for mine, ref in zip(clusters, run["clusters"]):
            assert [ident[order[k]] for k in mine.members] == ref["cands"]
            assert (mine.model_name, mine.row, mine.col, mine.count) == (ref["model"], ref["row"], ref["col"], ref["count"])
            assert (mine.occur, mine.score, mine.bpp, mine.mi, mine.h) == (ref["occur"], ref["score"], ref["bpp"], ref["mi"], ref["h"])
            assert str(mine) == ref["text"]
    eng.close()


def _synthetic_rows(n, seed, n_keys=7, span=4000):
    """Rows with many ties and dense overlaps: few distinct scores / E-values, positions crowded into `span` letters."""
    from rmdetect_b200._native import ROW_DTYPE
    rng = np.random.default_rng(seed)
    rows = np.zeros(n, dtype=ROW_DTYPE)
    rows["key"] = rng.integers(0, n_keys, n)
    rows["model"] = rng.integers(0, 4, n)
    leaves = np.array([16, 1, 2, 8])[rows["model"]]                 # CL, TGA, GB, KT gap configurations
    rows["leaf"] = (rng.integers(0, 1 << 30, n) % leaves).astype(np.uint16)
    rows["start"] = rng.integers(0, span, n)
    rows["wlen"] = 150
    rows["p0"] = rng.integers(0, 60, n)
    rows["p1"] = rng.integers(60, 130, n)
    rows["score"] = rng.integers(50, 90, n) / 8.0
    rows["evalue"] = 10.0 ** -rng.integers(0, 6, n).astype(np.float64)
    rows["bpp"] = np.where(rng.random(n) < 0.5, 1.0, 0.0)
    rows["id"] = np.arange(n, dtype=np.uint32)
    return rows


class _Row:
    """A row as the host restatement's Candidate (rmdetect_b200/candidates.py), positions from the model's offsets."""
    __slots__ = ("key", "model", "coords", "chains_pos", "score", "evalue", "bpp", "id")

    def __init__(self, r, keys, models, flats):
        m = int(r["model"])
        self.key, self.model = keys[int(r["key"])], models[m]
        s = int(r["start"])
        self.coords = (s, s + int(r["wlen"]))
        offs = flats[m].leaf_offsets[int(r["leaf"])]
        ptr = (int(r["p0"]), int(r["p1"]))
        self.chains_pos = [[None if o is None else s + ptr[c] + o for o in offs[c]] for c in range(2)]
        self.score, self.evalue, self.bpp, self.id = float(r["score"]), float(r["evalue"]), float(r["bpp"]), int(r["id"])

    def _sort_key(self):
        return (self.key, self.evalue, -self.score)

    def __lt__(self, other):
        return self._sort_key() < other._sort_key()


@pytest.mark.parametrize("n,span", [(3000, 600), (200000, 2000000)])
def test_synthetic_table_against_the_host_restatement(n, span):
    import ctypes as C
    from rmdetect_b200 import candidates as host
    from rmdetect_b200._native import _ptr
    from rmdetect_b200.hittable import HitTable
    eng = _engine()
    T = HitTable(eng)
    keys = ["seq%c" % c for c in "gbadfce"]                       # names whose order differs from their ids
    rows = _synthetic_rows(n, 11 + n, len(keys), span)
    objs = [_Row(r, keys, eng.models, eng.flats) for r in rows]

    def load():
        T.clear()
        for k in keys:
            T.key_id(k)
        cnt = C.c_int64(0)
        T.ctx._check(T.lib.rmd_table_append_rows(T.ctx.h, _ptr(rows), len(rows), C.byref(cnt)), "rmd_table_append_rows")

    ids = lambda lst: [o.id for o in lst]
    load()
    assert T.first_unsorted() == next(k for k in range(1, n) if objs[k] < objs[k - 1])
    T.sort()
    assert T.rows()["id"].tolist() == ids(sorted(objs)) and T.first_unsorted() == -1
    load(); T.top_model(50)
    assert T.rows()["id"].tolist() == ids(host.top_model(objs, 50))
    load(); T.top_seq(17)
    assert T.rows()["id"].tolist() == ids(host.top_seq(objs, 17))
    load(); T.top_seq_model(5)
    assert T.rows()["id"].tolist() == ids(host.top_seq_model(objs, 5))
    load(); T.get_model("gb_1.0")
    assert T.rows()["id"].tolist() == ids(host.get_model(objs, "gb_1.0"))
    load(); T.filter(min_score=8.0, min_bpp=0.5, min_evalue=3.0)
    assert T.rows()["id"].tolist() == ids(host.filter_candidates(objs, 8.0, 0.5, 3.0))
    if n <= 5000:                                                   # the restatement's nooverlap is quadratic
        load(); T.nooverlap()
        assert T.rows()["id"].tolist() == ids(host.nooverlap(objs))
        load(); T.sort(); T.nooverlap()
        assert T.rows()["id"].tolist() == ids(host.nooverlap(sorted(objs)))
        left = len(T)
        assert T.nooverlap() == left                                # idempotent: no two survivors are duplicates
    else:                                                           # sparse positions: the same answer group by group
        load(); T.nooverlap()
        left = set(T.rows()["id"].tolist())
        by_key = {}
        for o in objs:
            by_key.setdefault(o.key, []).append(o)
        want = set()
        for lst in by_key.values():                                 # hits of different sequences never interact
            lst_sorted = sorted(range(len(lst)), key=lambda k: lst[k].chains_pos[0][0])
            group, reach = [], -1
            for k in lst_sorted + [None]:
                if k is None or (group and lst[k].chains_pos[0][0] > reach):
                    want.update(o.id for o in host.nooverlap(sorted(group, key=lambda o: o.id)))
                    group, reach = [], -1
                if k is not None:
                    group.append(lst[k])
                    reach = max(reach, lst[k].chains_pos[0][-1])
        assert left == want
    eng.close()


def test_evalue_bound_is_the_reference_comparison():
    from rmdetect_b200.hittable import evalue_bound
    for m in (0.0, 3.0, 4.605170185988091, 11.5, -2.0):
        e = evalue_bound(m)
        below = np.nextafter(e, 0.0)
        assert -math.log(e) <= m and not (-math.log(below) <= m)


def test_rows_appended_from_a_scan_equal_its_candidates():
    from rmdetect_b200.hittable import HitTable
    from rmdetect_b200.scanner import Scanner
    from rmdetect_b200.synth import SynthFold
    triples = golden("seqsets")["lys.stk"][:12]
    models, evs = fresh_models(MODEL_ORDER)
    sc = Scanner(Cfg(models, evs))
    sc.rna_tools = SynthFold(5)
    cands, tries = sc.scan(SeqList(triples))
    eng = sc._engine
    T = HitTable(eng)
    T.append_scan([t[0] for t in triples])
    assert len(T) == len(cands) > 0
    cols = {t[0]: t[2] for t in triples}
    back = T.candidates(cols_of=cols.get)
    for c in back:
        c.bpp = 1.0
    for c in cands:
        c.bpp = 1.0
    assert sorted(str(c) for c in back) == sorted(str(c) for c in cands)
    # the CLI's flow on the device: score filter, sort, then the best hit per sequence
    from rmdetect_b200 import candidates as host
    want = host.top_seq(host.filter_candidates(cands, min_score=5.0), 1)
    T.filter(min_score=5.0)
    T.top_seq(1)
    got = T.candidates(cols_of=cols.get)
    for c in got:
        c.bpp = 1.0
    assert [str(c) for c in got] == [str(c) for c in want]
    eng.close()
