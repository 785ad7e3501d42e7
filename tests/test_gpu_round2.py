"""GPU parity at the level of whole result files, full-size inputs and the calibration loop.

* the complete example inputs of BASELINE configs[0] and [1] (lys.seq, lys.stk 49 sequences, rnaselci.stk 26,
  arich_test.stk 100) from `Scanner.scan` to the BYTES of the `.res` file, against what the reference's own flow
  (Scanner.scan -> Candidates.filter -> sort -> Results.write) wrote (tests/golden/res_files.json.gz);
* BASELINE configs[2] at full size (10 Mb): tests_all, tests_pairs, the number of hits before duplicate removal and
  an order-independent checksum of positions, gap configurations and score bits, against the oracle on all cores;
* rmd_calibrate_synth against the oracle fed the host twin of the device's sample draw; the calibration loop on the
  GPU against the reference's `.evalues` file.
"""
import math
import os

import numpy as np
import pytest

from helpers import LN_CUTOFF, LN_UNKNOWN, MODEL_ORDER, Cfg, SeqList, fresh_models, golden, load_evalues, load_model

pytestmark = pytest.mark.gpu

RES_CASES = ["lys_seq_kt", "lys_seq_both_all", "rnasel_seq_all", "rnasel_stk_all", "lys_stk_all", "arich_stk"]


@pytest.mark.parametrize("name", RES_CASES)
def test_result_file_bytes_equal_the_reference(name, tmp_path):
    from rmdetect_b200.pipeline import ScanSession
    from rmdetect_b200.scanner import Scanner
    from rmdetect_b200.synth import SynthFold
    case = next(c for c in golden("res_files") if c["name"] == name)
    triples = golden("seqsets")[case["seqset"]]
    models, evs = fresh_models(case["models"])
    sc = Scanner(Cfg(models, evs))
    sc.rna_tools = SynthFold(case["seed"])
    out = tmp_path / (name + ".res")
    session = ScanSession(sc, SeqList(triples), str(out), min_score=5.0, min_bpp=0.001, seq_file=case["seq_file"],
                          both_strands=case["both_strands"], n_sequences=case["n_sequences"])
    cands, tries = session.run()
    sc._engine.close()
    assert tries == case["tries"]
    keys = [t[0] for t in triples]
    names = [m.full_name() for m in models]
    got = [[names.index(c.model.full_name()), keys.index(c.key), c.coords[0], c.chains_pos[0][0], c.chains_pos[1][0],
            c.score, c.evalue] for c in cands]
    assert got == case["all"]                                   # every hit of the scan, exact floats
    assert out.read_text() == case["res"]                       # what rmout / rmfilter / rmcluster will read
    assert session.saves[0] == ("n", 0) and session.saves[-1][0] == "w"


@pytest.mark.parametrize("name", ["lys_seq_both_all", "rnasel_stk_all", "arich_stk"])
def test_result_file_from_the_input_file_without_a_reference_checkout(name, tmp_path):
    """The standalone flow: the example input is read by rmdetect_b200.sequences.Sequences (strands,
    gaps, column index, the header's sequence count), scanned, filtered, sorted and written — the bytes of the reference's
    own result file."""
    from rmdetect_b200.pipeline import ScanSession
    from rmdetect_b200.scanner import Scanner
    from rmdetect_b200.sequences import Sequences
    from rmdetect_b200.synth import SynthFold
    case = next(c for c in golden("res_files") if c["name"] == name)
    here = os.path.dirname(os.path.abspath(__file__))
    seqs = Sequences(fname=os.path.join(here, "golden", "ref_data", "examples", os.path.basename(case["seq_file"])),
                     both_strands=case["both_strands"])
    models, evs = fresh_models(case["models"])
    sc = Scanner(Cfg(models, evs))
    sc.rna_tools = SynthFold(case["seed"])
    out = tmp_path / (name + ".res")
    session = ScanSession(sc, seqs, str(out), min_score=5.0, min_bpp=0.001, seq_file=case["seq_file"])
    cands, tries = session.run()
    sc._engine.close()
    assert tries == case["tries"] and session.n_sequences == case["n_sequences"]
    assert out.read_text() == case["res"]


def test_ten_megabases_against_the_oracle():
    """BASELINE configs[2]: 10 000 000 letters, 133 333 windows, 9.46e9 placements, four models."""
    from oracle import pyoracle
    from rmdetect_b200.genome import hit_checksum
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import synth_sequence
    n, seed = 10_000_000, 20261019
    models = [load_model(m) for m in MODEL_ORDER]
    eng = ScanEngine(models, [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    neutral = gc_table({b: math.log(0.25) for b in "ACGU"})
    n_win, n_bpp, _ = eng.ctx.synth_job(n, 0, seed, seed, 150, 75, 0.001, LN_CUTOFF, neutral)
    counts = eng.ctx.gc_counts(1)[0][:4]
    gc = {b: math.log(float(counts[k]) / n) for k, b in enumerate("ACGU")}
    cls = []
    for ev in eng.evalues:
        ev.select_gc(gc)
        cls.append(ev.selected)
    eng.ctx.set_gc(gc_table(gc).reshape(1, -1), np.array([cls], dtype=np.int32))
    eng.ctx.set_return_dropped(True)
    res = eng.ctx.scan_resident(0, n_win)
    h = res["hits"]
    assert len(h) == res["n_raw"]
    want = pyoracle.bench_scan2(models, synth_sequence(n, seed), 150, 75, 0, n_win, seed, os.cpu_count() or 1, 0.001,
                                LN_UNKNOWN, LN_CUTOFF, counts=counts)
    assert res["tries"] == want["tries"] and want["tries"][0] == 9458509687
    assert res["n_raw"] == want["hits"]
    assert hit_checksum(0, h["start"], h["p0"], h["p1"], h["model"], h["leaf"], h["score"]) == want["checksum"]
    eng.close()


@pytest.mark.parametrize("name", ["KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0"])
def test_device_drawn_calibration_samples_against_the_oracle(name):
    """rmd_calibrate_synth (the variant bench.py times): its samples are those of the host twin, and its histograms
    are the oracle's for those samples — same buckets, weights to 1e-9 (different summation order)."""
    from oracle import pyoracle
    from rmdetect_b200.calibrate import N_BINS, class_logs
    from rmdetect_b200.gcx import GC
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import synth_calibration_samples
    m = load_model(name)
    L = len(m.nodes)
    seed, first, n = 20261019, 123457, 40000
    classes = GC.get_classes(5, 15, 85)
    cl = class_logs(classes)
    eng = ScanEngine([m], [None], LN_UNKNOWN)
    dev = np.zeros((len(classes), N_BINS))
    ok_a, _ = eng.ctx.calibrate(0, None, cl, LN_CUTOFF, dev, seed=seed, first=first, n=25000, L=L)      # two calls: `first`
    ok_b, _ = eng.ctx.calibrate(0, None, cl, LN_CUTOFF, dev, seed=seed, first=first + 25000, n=n - 25000, L=L)
    samples = synth_calibration_samples(seed, first, n, L)
    host = np.zeros_like(dev)
    ok_h, _ = eng.ctx.calibrate(0, samples, cl, LN_CUTOFF, host)
    eng.close()
    assert ok_a + ok_b == ok_h
    assert np.array_equal(dev != 0, host != 0) and np.allclose(dev, host, rtol=1e-12, atol=0)
    want, ok_o = pyoracle.calibrate(pyoracle.OracleModel(m), ["".join("ACGU"[c] for c in row) for row in samples], classes,
                                    LN_UNKNOWN, LN_CUTOFF, N_BINS)
    assert ok_o == ok_h
    assert np.array_equal(dev != 0, want != 0)                  # exactly the reference's score buckets
    assert np.allclose(dev, want, rtol=1e-9, atol=0)


LONG_MODEL = """NAME    LONG 1.0
%(nodes)s
PROB    0   *   GC
PROB    1   *   A:0.30          G:0.70
PROB    2   *   A:0.55  C:0.15  G:0.15  U:0.15
PROB    3   *   GC
PROB    4   *   GC
PROB    5   *   A:0.25  C:0.25  G:0.25  U:0.25
PROB    6   *   GC
PROB    7   *           C:0.50          U:0.50
PROB    8   *   GC
PROB    9   *   GC
PROB    10  *   GC
PROB    11  *   A:0.40  C:0.10  G:0.40  U:0.10
PROB    12  *   GC
PROB    13  *   GC
PROB    14  *   GC
PROB    15  *   GC
PROB    16  *   GC
PROB    17  A                           U:1.00
PROB    17  C                   G:1.00
PROB    17  G           C:0.60          U:0.40
PROB    17  U   A:0.60          G:0.40
PROB    18  *                   G:0.80  U:0.20
PROB    19  *   A:1.00
PROB    20  A                           U:1.00
PROB    20  C                   G:1.00
PROB    20  G           C:0.60          U:0.40
PROB    20  U   A:0.60          G:0.40

PAIRS   0   20  MUST
PAIRS   16  17  MUST

NO_PAIRING  1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 18 19

ORDER   %(order)s

SEPARATION  3   -
"""


def test_calibration_of_a_model_the_engine_cannot_hold(tmp_path):
    """A chain of 17 letters does not fit the scoring engine's 16-letter words: calibration walks the tree through
    eval_placement, one sample per thread, and (21 positions) counts the letter compositions in global memory.  Also a
    cutoff that fails the root's own test (lib/node.py:248): nobody has a solution, every weight goes to bucket 0.  Both
    against the oracle: same buckets, weights to 1e-9."""
    from oracle import pyoracle
    from rmdetect_b200.calibrate import N_BINS, class_logs
    from rmdetect_b200.gcx import GC
    from rmdetect_b200.modelfile import parse_model
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import synth_calibration_samples
    nodes = ["NODE    %d   0   %d   -" % (k, k) for k in range(17)]
    nodes += ["NODE    17  1   0   16", "NODE    18  1   1   -", "NODE    19  1   2   -", "NODE    20  1   3   0"]
    path = tmp_path / "long_1.model"
    path.write_text(LONG_MODEL % dict(nodes="\n".join(nodes), order=" ".join(str(k) for k in range(21))))
    classes = GC.get_classes(5, 15, 85)
    cl = class_logs(classes)
    for m, cutoff, n in ((parse_model(str(path)), LN_CUTOFF, 6000), (load_model("KT_1.0"), 0.25, 3000)):
        L = len(m.nodes)
        samples = synth_calibration_samples(7, 1000, n, L)
        eng = ScanEngine([m], [None], LN_UNKNOWN)
        dev = np.zeros((len(classes), N_BINS))
        ok_d, _ = eng.ctx.calibrate(0, None, cl, cutoff, dev, seed=7, first=1000, n=n, L=L)
        host = np.zeros_like(dev)
        ok_h, _ = eng.ctx.calibrate(0, samples, cl, cutoff, host)
        eng.close()
        want, ok_o = pyoracle.calibrate(pyoracle.OracleModel(m), ["".join("ACGU"[c] for c in row) for row in samples], classes,
                                        LN_UNKNOWN, cutoff, N_BINS)
        assert ok_d == ok_h == ok_o and (ok_o > 0) == (cutoff < 0)
        assert np.array_equal(dev != 0, host != 0) and np.allclose(dev, host, rtol=1e-12, atol=0)
        assert np.array_equal(dev != 0, want != 0)
        assert np.allclose(dev, want, rtol=1e-9, atol=0)


def test_calibration_of_every_model_of_one_context_with_other_cutoffs():
    """All shipped models resident in one context (the engine's nodes and tables are concatenated: a model's walk starts at
    its own first node), each calibrated with a cutoff of its own — 1e-12 lets a quarter of the samples reach a leaf (the list
    of solved samples is the main path then), 0.5 almost none: buckets and weights of the oracle."""
    from oracle import pyoracle
    from helpers import MODEL_FILES
    from rmdetect_b200.calibrate import N_BINS, class_logs
    from rmdetect_b200.gcx import GC
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import synth_calibration_samples
    names = sorted(MODEL_FILES)
    models = [load_model(n) for n in names]
    classes = GC.get_classes(5, 15, 85)
    cl = class_logs(classes)
    eng = ScanEngine(models, [None] * len(models), LN_UNKNOWN)
    solved = []
    for mi, (m, c) in enumerate(zip(models, [1e-8, 0.5, 1e-6, 0.1, 1e-12])):
        cutoff, L, n = math.log(c), len(m.nodes), 12000
        dev = np.zeros((len(classes), N_BINS))
        ok_d, _ = eng.ctx.calibrate(mi, None, cl, cutoff, dev, seed=99 + mi, first=5 * mi, n=n, L=L)
        samples = synth_calibration_samples(99 + mi, 5 * mi, n, L)
        want, ok_o = pyoracle.calibrate(pyoracle.OracleModel(m), ["".join("ACGU"[c] for c in row) for row in samples], classes,
                                        LN_UNKNOWN, cutoff, N_BINS)
        assert ok_d == ok_o, (names[mi], ok_d, ok_o)
        assert np.array_equal(dev != 0, want != 0), names[mi]
        assert np.allclose(dev, want, rtol=1e-9, atol=0), names[mi]
        solved.append(ok_o)
    eng.close()
    assert max(solved) > 1000 and min(solved) < 100, solved


def test_one_row_checkpoint_read_equals_the_histogram():
    """rmd_calib_read_rows: the middle class's row and the occupied bucket range of a device-resident session equal those of
    the downloaded histogram, batch after batch."""
    from rmdetect_b200.calibrate import Calibrator
    m = load_model("GB_1.0")
    c = Calibrator(m, LN_UNKNOWN, LN_CUTOFF)
    mid = len(c.classes) // 2
    for k in range(3):
        c.add_synthetic(11, 20000 * k, 20000)
        row, cols = c.reduced_row(mid)
        rows2, _, _ = c.ctx.calib_read_rows(3, 4)
        h = c.reduced_hist().copy()
        occupied = np.flatnonzero(h.any(axis=0))
        assert cols == (int(occupied[0]), int(occupied[-1]))
        assert np.array_equal(row[0], h[mid]) and np.array_equal(rows2, h[3:7])
    with pytest.raises(Exception):
        c.ctx.calib_read_rows(len(c.classes) - 1, 2)
    c.close()


def test_calibration_loop_on_the_gpu_writes_the_reference_file(tmp_path):
    """calibrate_model end to end on the device (seeded reference sample stream, checkpoints, stop rule)."""
    from rmdetect_b200.calibrate import calibrate_model
    for name in ("KT_1.0", "TGA_1.0"):
        g = golden("calibrate")[name]
        out = tmp_path / (name + ".evalues")
        _, _, done = calibrate_model(load_model(name), str(out), samples=g["samples"], seed=g["seed"], log=None)
        assert done == min(g["samples"], 4 ** len(load_model(name).nodes))
        assert out.read_text() == g["text"]


def test_a_million_samples_write_the_reference_file(tmp_path):
    """BASELINE configs[3] at its full size: KT_1.0, --samples 1000000, the reference's seeded sample stream, a table file at
    every checkpoint, the stop rule — the `.evalues` text the reference's own run leaves behind (five minutes of one core
    there: tests/golden/make_calibrate_full.py), byte for byte, and the same number of simulations."""
    import hashlib
    from rmdetect_b200.calibrate import calibrate_model
    g = golden("calibrate_full")["KT_1.0"]
    out = tmp_path / "kt_full.evalues"
    _, _, done = calibrate_model(load_model("KT_1.0"), str(out), samples=g["samples"], seed=g["seed"], log=None)
    text = out.read_text()
    assert done == (g["converged_after"] if g["converged_after"] is not None else g["samples"])
    assert hashlib.sha256(text.encode()).hexdigest() == g["sha256"] and text == g["text"]


def test_compact_hit_records_carry_the_same_hits():
    """rmd_set_hit_format(1): 32-byte records on the way back; expand_hits rebuilds start / wlen from the geometry."""
    from rmdetect_b200._native import HIT32_DTYPE, expand_hits
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    seqs = [("a", synth_sequence(2000, 3), None), ("b", synth_sequence(100, 4), None), ("c", synth_sequence(431, 5), None)]
    for W, step in ((150, 75), (64, 17), (0, 0)):
        job = eng.build_job(seqs, W, step, SynthFold(6), "", 0.001, LN_CUTOFF)
        for dropped in (False, True):
            eng.ctx.set_return_dropped(dropped)
            eng.ctx.set_hit_format(False)
            want = eng.scan_job(job)["hits"]
            eng.ctx.set_hit_format(True)
            got = eng.scan_job(job)["hits"]
            assert got.dtype == HIT32_DTYPE and len(got) == len(want) > 0
            full = expand_hits(got, job["seq_off"], W, step)
            for f in ("window", "start", "seq", "wlen", "p0", "p1", "model", "leaf", "keep", "score", "evalue"):
                assert np.array_equal(full[f], want[f]), f
    eng.close()
