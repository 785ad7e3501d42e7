"""Calibration bookkeeping (rmdetect_b200/calibrate.py) against the reference's own rmcalibrate.py output.

tests/golden/calibrate.json.gz holds the histograms and the `.evalues` text the reference wrote for seeded runs,
calibrate_ckpt.json.gz its table after 10 000 samples and the convergence measure between the two tables
(make_golden.py: calibrate_vectors, calibrate_checkpoints).  The sample scoring itself is the GPU's (test_gpu_parity.py);
here the oracle's `orc_calibrate` stands in for it so that the loop — checkpoints, stop rule, file — and its
multi-rank form (world-size-2 gloo, histograms all-reduced per block) run on the CPU.
"""
import os
import random

import numpy as np
import pytest

from helpers import LN_CUTOFF, LN_UNKNOWN, golden, load_model
from rmdetect_b200 import calibrate as cal
from rmdetect_b200.gcx import GC

MODELS = ["KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0"]


def dense(pairs_per_class):
    h = np.zeros((len(pairs_per_class), cal.N_BINS))
    for ci, pairs in enumerate(pairs_per_class):
        for k, v in pairs:
            h[ci, int(round(k * 10))] = v
    return h


@pytest.mark.parametrize("name", MODELS)
def test_table_text_equals_the_reference_file(name):
    g = golden("calibrate")[name]
    L = len(load_model(name).nodes)
    classes = GC.get_classes(5, 15, 85)
    hist = dense(g["hists"])
    gce, scores = cal.gc_evalues(cal.hist_dicts(hist), 1.0 / float(4 ** L))
    assert cal.table_text(classes, gce, scores) == g["text"]
    tab, scores_d = cal.gc_evalues_dense(hist, 1.0 / float(4 ** L))
    assert scores_d == scores
    for ci, row in enumerate(gce):                                  # the dense form is the dict form, bit for bit
        assert [row[s] for s in scores] == tab[ci].tolist()
    assert cal.table_text_dense(classes, tab, scores) == g["text"]
    # structure: first row 1.0 (give or take the floor and a top row the float grid drops), columns non-increasing, floor 4^-L
    assert np.allclose(tab[:, 0], 1.0, rtol=0, atol=1e-4) and np.all(np.diff(tab, axis=1) <= 0) and tab.min() >= 1.0 / float(4 ** L)


@pytest.mark.parametrize("name", MODELS)
def test_convergence_measure_equals_the_reference(name):
    g, full = golden("calibrate_ckpt")[name], golden("calibrate")[name]
    L = len(load_model(name).nodes)
    ha, hb = dense(g["hists_a"]), dense(full["hists"])
    ga, sa = cal.gc_evalues(cal.hist_dicts(ha), 1.0 / float(4 ** L))
    gb, sb = cal.gc_evalues(cal.hist_dicts(hb), 1.0 / float(4 ** L))
    assert sb == g["scores"]
    mid = len(gb) // 2
    assert sorted([k, v] for k, v in ga[mid].items()) == g["middle_a"]
    assert sorted([k, v] for k, v in gb[mid].items()) == g["middle_b"]
    assert cal.diff_gc_evalues(ga, gb, sb) == g["diff"]
    ta, sa_d = cal.gc_evalues_dense(ha, 1.0 / float(4 ** L))
    tb, sb_d = cal.gc_evalues_dense(hb, 1.0 / float(4 ** L))
    assert cal.diff_dense(ta, sa_d, tb, sb_d) == g["diff"]


@pytest.mark.parametrize("n,L,seed", [(1, 1, 0), (7, 16, 1), (1000, 20, 20261019), (10000, 8, 5), (0, 5, 1), (2500, 14, 123)])
def test_bulk_sample_draw_is_the_randint_stream(n, L, seed):
    """draw_samples takes the Mersenne Twister's outputs in bulk: the same letters as `random.randint(0, 3)` call by call
    (rmcalibrate.py:16-22), and the generator is left in the same state, so later draws agree too."""
    a = cal.draw_samples_loop(n, L, seed)
    after_a = [random.random(), random.randint(0, 10 ** 9), random.getrandbits(70)]
    b = cal.draw_samples(n, L, seed)
    after_b = [random.random(), random.randint(0, 10 ** 9), random.getrandbits(70)]
    assert a.shape == b.shape == (n, L) and b.dtype == np.uint8 and np.array_equal(a, b) and after_a == after_b


def test_bulk_sample_draw_continues_a_stream_in_use():
    random.seed(42)
    x1 = cal.draw_samples_loop(500, 16); g1 = random.gauss(0, 1); y1 = cal.draw_samples_loop(300, 16); s1 = random.getstate()
    random.seed(42)
    x2 = cal.draw_samples(500, 16); g2 = random.gauss(0, 1); y2 = cal.draw_samples(300, 16); s2 = random.getstate()
    assert np.array_equal(x1, x2) and g1 == g2 and np.array_equal(y1, y2) and s1 == s2


def test_table_text_row_by_row_equals_number_by_number():
    """table_text_dense formats a score row with one call; the bytes are those of the reference's number-by-number loop
    (rmcalibrate.py:101-131), also for values that print with three-digit exponents and for a class without any hit."""
    rng = np.random.default_rng(3)
    h = np.zeros((165, cal.N_BINS))
    h[:, :150] = rng.random((165, 150)) * 10.0 ** rng.integers(-300, 3, (165, 150))
    h[7, :] = 0.0
    h[7, 0] = 1.0
    h[3, 170] = 1e-9
    classes = GC.get_classes(5, 15, 85)
    tab, scores = cal.gc_evalues_dense(h, 4.0 ** -16)
    want = cal.table_text(classes, [dict(zip(scores, row.tolist())) for row in tab], scores)
    assert cal.table_text_dense(classes, tab, scores) == want and len(scores) == 171


def _diff_loop(tab1, scores1, tab2, scores2, gc):
    """rmcalibrate.py:134-157 row by row (what diff_dense has to equal bit for bit)."""
    at1 = {s: k for k, s in enumerate(scores1)}
    diff = atotal = 0.0
    for i in range(1, len(scores2)):
        a, b = scores2[i - 1], scores2[i]
        if a in at1 and b in at1:
            ya1, yb1 = float(tab1[gc, at1[a]]), float(tab1[gc, at1[b]])
            ya2, yb2 = float(tab2[gc, i - 1]), float(tab2[gc, i])
            x = b - a
            a1 = x * (yb1 + abs(yb1 - ya1) / 2.0)
            a2 = x * (yb2 + abs(yb2 - ya2) / 2.0)
            diff += abs(a1 - a2)
            atotal += a1
    return diff / atotal


def test_convergence_measure_in_bulk_equals_the_row_by_row_loop():
    """diff_dense forms its terms with numpy and sums them with a running sum: the loop's value exactly, for grids that
    start and end at different scores (rows the older table lacks are skipped), with cached and with plain score lists."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        n1, n2 = int(rng.integers(20, 200)), int(rng.integers(20, 200))
        f1, f2 = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        h1, h2 = np.zeros((3, cal.N_BINS)), np.zeros((3, cal.N_BINS))
        h1[:, f1:f1 + n1] = rng.random((3, n1)) * 10.0 ** rng.integers(-8, 3, (3, n1))
        h2[:, f2:f2 + n2] = rng.random((3, n2)) * 10.0 ** rng.integers(-8, 3, (3, n2))
        t1, s1 = cal.gc_evalues_dense(h1, 4.0 ** -16)
        t2, s2 = cal.gc_evalues_dense(h2, 4.0 ** -16)
        assert s1 == cal.score_grid(h1) and list(s1.idx) == [int(round(v * 10)) for v in s1]
        want = _diff_loop(t1, s1, t2, s2, 1)
        assert cal.diff_dense(t1, s1, t2, s2) == want == cal.diff_dense(t1, list(s1), t2, list(s2), gc=1)
    with pytest.raises(ZeroDivisionError):                        # no common row pair: the reference divides 0.0 by 0.0 too
        h1, h2 = np.zeros((1, cal.N_BINS)), np.zeros((1, cal.N_BINS))
        h1[0, 0:3] = 1.0
        h2[0, 40:43] = 1.0
        t1, s1 = cal.gc_evalues_dense(h1, 1e-9)
        t2, s2 = cal.gc_evalues_dense(h2, 1e-9)
        cal.diff_dense(t1, s1, t2, s2)


class OracleScorer:
    """CPU stand-in for calibrate.Calibrator: the oracle's restatement of rmcalibrate.py:265-295."""

    def __init__(self, name):
        from oracle import pyoracle
        self.py = pyoracle
        self.om = pyoracle.OracleModel(load_model(name))
        self.L = len(load_model(name).nodes)
        self.classes = GC.get_classes(5, 15, 85)
        self.hist = np.zeros((len(self.classes), cal.N_BINS))

    def add_samples(self, samples):
        strings = ["".join("ACGU"[c] for c in row) for row in samples]
        _, nok = self.py.calibrate(self.om, strings, self.classes, LN_UNKNOWN, LN_CUTOFF, cal.N_BINS, self.hist)
        return nok

    def reduced_hist(self, dist=None):
        self.total = cal.all_reduce_hist(self.hist, dist)
        return self.total

    def add_synthetic(self, seed, first, n):
        from rmdetect_b200.synth import synth_calibration_samples
        return self.add_samples(synth_calibration_samples(seed, first, n, self.L))


class RowScorer(OracleScorer):
    """The same with the one-row checkpoint read of calibrate.Calibrator.reduced_row (rmd_calib_read_rows)."""

    def reduced_row(self, cls, dist=None):
        occupied = np.flatnonzero(self.hist.any(axis=0))
        cols = (int(occupied[0]), int(occupied[-1])) if len(occupied) else (-1, -1)
        self.rows_read = getattr(self, "rows_read", 0) + 1
        return cal.all_reduce_hist(self.hist[cls:cls + 1], dist), cal.all_reduce_cols(cols, dist)


def _row_rank_main(rank, world, port, name, samples, seed, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = {}
        tab, scores, done = cal.calibrate_model(load_model(name), None, samples=samples, seed=seed, log=None, dist=dist,
                                                scorer=RowScorer(name), device_samples=True, save_every_block=False, stats=st)
        np.save(os.path.join(out_dir, "rowtab_r%d.npy" % rank), tab)
        np.save(os.path.join(out_dir, "rowdone_r%d.npy" % rank), np.array([done, st["blocks"]]))
    finally:
        dist.destroy_process_group()


def test_checkpoints_that_read_one_row_decide_as_the_full_table_does(tmp_path):
    """Throughput mode (save_every_block=False): a checkpoint reads the middle class's row and the occupied bucket range
    instead of the whole histogram — same score rows, same diff, same stop, same final table; on one rank and on two."""
    import torch.multiprocessing as mp
    name, samples, seed = "TGA_1.0", 60000, 7
    full, rows = OracleScorer(name), RowScorer(name)
    tab_f, scores_f, done_f = cal.calibrate_model(load_model(name), None, samples=samples, seed=seed, log=None, scorer=full,
                                                  device_samples=True, save_every_block=False)
    tab_r, scores_r, done_r = cal.calibrate_model(load_model(name), None, samples=samples, seed=seed, log=None, scorer=rows,
                                                  device_samples=True, save_every_block=False)
    assert rows.rows_read >= 5 and done_f == done_r and scores_f == scores_r and np.array_equal(tab_f, tab_r)
    # the middle row's table from the row alone equals that row of the full table (the score grid spans every class)
    h = full.hist
    mid = h.shape[0] // 2
    occupied = np.flatnonzero(h.any(axis=0))
    t_row, s_row = cal.gc_evalues_dense(h[mid:mid + 1], 4.0 ** -8, cols=(int(occupied[0]), int(occupied[-1])))
    t_all, s_all = cal.gc_evalues_dense(h, 4.0 ** -8)
    assert s_row == s_all and np.array_equal(t_row[0], t_all[mid])
    assert np.flatnonzero(h[mid])[-1] < occupied[-1]                # ... which the row by itself would not tell
    port = 29500 + random.Random(os.getpid() + 17).randrange(2000)
    mp.spawn(_row_rank_main, args=(2, port, name, samples, seed, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "rowtab_r0.npy"), np.load(tmp_path / "rowtab_r1.npy")
    assert np.array_equal(t0, t1) and np.allclose(t0, tab_f, rtol=1e-12, atol=0)
    assert np.array_equal(np.load(tmp_path / "rowdone_r0.npy"), np.load(tmp_path / "rowdone_r1.npy"))


@pytest.mark.parametrize("name", ["KT_1.0", "TGA_1.0"])
def test_calibration_loop_writes_the_reference_file(name, tmp_path):
    """calibrate_model with the reference's seeded sample stream: checkpoints every 10 000 samples, the stop rule,
    min(samples, 4^L) (TGA: 65 536 of the 70 000 asked for) — the final file is the reference's, byte for byte."""
    g = golden("calibrate")[name]
    out = tmp_path / "x.evalues"
    tab, scores, done = cal.calibrate_model(load_model(name), str(out), samples=g["samples"], seed=g["seed"], log=None,
                                            scorer=OracleScorer(name))
    assert done == min(g["samples"], 4 ** len(load_model(name).nodes))
    assert out.read_text() == g["text"]


def test_a_million_samples_requested_stop_where_the_reference_stops(tmp_path):
    """BASELINE configs[3] at its full size through the oracle (the GPU twin of this test is in test_gpu_round2.py): KT_1.0,
    --samples 1000000, seeded stream; the reference's own run (tests/golden/make_calibrate_full.py, 98 s) converges after
    370 000 simulations — the loop here stops there too and has written the same bytes."""
    g = golden("calibrate_full")["KT_1.0"]
    out = tmp_path / "kt_full.evalues"
    _, _, done = cal.calibrate_model(load_model("KT_1.0"), str(out), samples=g["samples"], seed=g["seed"], log=None,
                                     scorer=OracleScorer("KT_1.0"))
    assert done == g["converged_after"] == 370000
    assert out.read_text() == g["text"]


def _rank_main(rank, world, port, name, samples, seed, device_samples, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = OracleScorer(name)
        tab, scores, done = cal.calibrate_model(load_model(name), os.path.join(out_dir, "w%d.evalues" % world), samples=samples,
                                                seed=seed, log=None, dist=dist, scorer=sc, device_samples=device_samples)
        np.save(os.path.join(out_dir, "hist_r%d.npy" % rank), sc.total)
        np.save(os.path.join(out_dir, "tab_r%d.npy" % rank), tab)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("device_samples", [False, True])
def test_two_ranks_reduce_to_the_single_rank_histogram(device_samples, tmp_path):
    """World size 2 over gloo: each rank scores half of every block, the blocks' histograms are all-reduced; the final
    histogram equals the single-rank one (same buckets; sums agree to rounding) on both ranks, and so does the file."""
    import torch.multiprocessing as mp
    name, samples, seed = "GB_1.0", 30000, 4242
    port = 29500 + random.Random(os.getpid() + device_samples).randrange(2000)
    mp.spawn(_rank_main, args=(2, port, name, samples, seed, device_samples, str(tmp_path)), nprocs=2, join=True)
    one = OracleScorer(name)
    tab1, scores1, done1 = cal.calibrate_model(load_model(name), str(tmp_path / "w1.evalues"), samples=samples, seed=seed,
                                               log=None, scorer=one, device_samples=device_samples)
    h0, h1 = np.load(tmp_path / "hist_r0.npy"), np.load(tmp_path / "hist_r1.npy")
    assert np.array_equal(h0, h1)                                   # the reduced histogram is the same everywhere
    assert np.array_equal(h0 != 0, one.total != 0)
    assert np.allclose(h0, one.total, rtol=1e-12, atol=0)
    assert np.array_equal(np.load(tmp_path / "tab_r0.npy"), np.load(tmp_path / "tab_r1.npy"))
    assert (tmp_path / "w2.evalues").read_text() == (tmp_path / "w1.evalues").read_text()
    assert done1 == samples
