"""GPU parity: the CUDA path through the C ABI against the reference's golden outputs and the oracle.

Bit-exact for hit sets, positions, gap configurations, tests_all / tests_pairs and (fp64, same
addition order) log-probabilities, scores and E-values.  Calibration histograms are sums of ~1e4-1e6
positive weights accumulated in a different order: relative tolerance 1e-9 (north star allows 1e-6).
"""
import math
import os
import random

import numpy as np
import pytest

from helpers import (LN_CUTOFF, LN_UNKNOWN, MODEL_FILES, MODEL_ORDER, REF_DATA, Cfg, SeqList, evalues_path,
                     fresh_models, golden, load_evalues, load_model)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine_all():
    from rmdetect_b200.scanner import ScanEngine
    names = sorted(MODEL_FILES)
    eng = ScanEngine([load_model(n) for n in names], [load_evalues(n) for n in names], LN_UNKNOWN)
    yield names, eng
    eng.close()


@pytest.mark.parametrize("name", sorted(MODEL_FILES))
def test_eval_vectors(engine_all, name):
    from rmdetect_b200.gcx import encode_sequence, gc_table
    names, eng = engine_all
    mi = names.index(name)
    cases = [c for c in golden("eval_vectors")[name] if c["unknown_prob"] == LN_UNKNOWN]
    assert len(cases) == 400
    for cutoff in sorted({c["cutoff"] for c in cases}):
        sub = [c for c in cases if c["cutoff"] == cutoff]
        codes, off, gcs = [], [0], []
        for c in sub:
            cd, others = encode_sequence(c["seq"])
            codes.append(cd)
            off.append(off[-1] + len(cd))
            gcs.append(gc_table(c["gc"], others))
        leaf, pm, pr = eng.ctx.eval(mi, np.concatenate(codes), off, np.array(gcs), [c["p"][0] for c in sub],
                                    [c["p"][1] for c in sub], cutoff)
        for c, lf, a, b in zip(sub, leaf, pm, pr):
            assert (lf >= 0) == c["ok"]
            if c["ok"]:
                assert a == c["prob_model"] and b == c["prob_rand"]
                offs = eng.flats[mi].leaf_offsets[lf]
                seqs = ["".join("." if o is None else c["seq"][c["p"][k] + o] for o in offs[k]) for k in range(2)]
                assert seqs == c["seqs"]


def run_case(case):
    from rmdetect_b200.scanner import Scanner
    from rmdetect_b200.synth import SynthFold
    models, evs = fresh_models(case["models"])
    cfg = Cfg(models, evs, case["win_len"], case["win_step"], case["min_bpp"], case["cutoff"], case["unknown"])
    sc = Scanner(cfg)
    sc.rna_tools = SynthFold(case["seed"])
    events = []
    sc.cback_before_scan = lambda n: events.append(["before_scan", n])
    sc.cback_before_sequence = lambda n, i, k: events.append(["before_sequence", n, i, k])
    sc.cback_before_window = lambda n, p: events.append(["before_window", n, p])
    sc.cback_after_window = lambda c, t: events.append(["after_window", len(c), list(t)])
    sc.cback_after_sequence = lambda c, t: events.append(["after_sequence", len(c), list(t)])
    sc.cback_after_scan = lambda c, t: events.append(["after_scan", len(c), list(t)])
    cands, tries = sc.scan(SeqList(case["seqs"]))
    sc._engine.close()
    return cands, tries, events


@pytest.mark.parametrize("idx", range(15))
def test_scan_cases_match_reference(idx):
    """Scanner.scan on the GPU == the reference's Scanner.scan (golden), field by field."""
    case = golden("scan_cases")[idx]
    cands, tries, events = run_case(case)
    assert tries == case["tries"]
    assert len(cands) == len(case["cands"])
    for got, want in zip(cands, case["cands"]):
        assert got.model.full_name() == want["model"] and got.key == want["key"]
        assert list(got.coords) == want["coords"]
        assert ["".join(s) for s in got.chains_seqs] == want["seqs"]
        assert got.chains_pos == want["pos"] and got.chains_cols == want["cols"]
        assert got.score == want["score"] and got.evalue == want["evalue"]
    if case["events"]:
        assert events == case["events"]


def test_real_bpp_windows_through_c_abi():
    """The two real RNAfold matrices shipped with the reference."""
    from rmdetect_b200.dotps import read_dot_ps
    from rmdetect_b200.scanner import ScanEngine

    class Fold:
        def __init__(self, table):
            self.table = table

        def coo(self, win):
            z = np.zeros(0, np.uint16)
            return self.table.get(win, (z, z, np.zeros(0)))

    key = "ABCZ01000009.1/166331-166512"
    seq = next(s for k, s, _ in golden("seqsets")["lys.stk"] if k == key)
    wseq, i, j, p = read_dot_ps(os.path.join(REF_DATA, "lys_dot.ps"))
    eng = ScanEngine([load_model(n) for n in MODEL_ORDER], [load_evalues(n) for n in MODEL_ORDER], LN_UNKNOWN)
    job = eng.build_job([(key, seq, None)], 150, 75, Fold({wseq: (i, j, p)}), "", 0.001, LN_CUTOFF)
    res = eng.scan_job(job)
    assert res["gate_pass"].sum(axis=1).tolist() == [0, 1806]
    assert res["tries"] == [2 * 70939, 1806]
    got = {(MODEL_ORDER[h["model"]], int(h["start"] + h["p0"]), int(h["start"] + h["p1"])): h for h in res["hits"]}
    kt = got[("KT_1.0", 164, 81)]
    assert ("%.5E" % kt["score"], "%.5E" % kt["evalue"]) == ("1.04455E+01", "3.04000E-05")
    tga = got[("TGA_1.0", 40, 56)]
    assert ("%.5E" % tga["score"], "%.5E" % tga["evalue"]) == ("8.21393E+00", "1.04900E-03")
    eng.close()

    wseq, i, j, p = read_dot_ps(os.path.join(REF_DATA, "arich", "dot.ps"))
    eng = ScanEngine([load_model("ARICH_1.0")], [load_evalues("ARICH_1.0")], LN_UNKNOWN)
    res = eng.scan_job(eng.build_job([("a", wseq, None)], 150, 75, Fold({wseq: (i, j, p)}), "", 0.001, LN_CUTOFF))
    assert res["tries"] == [9506, 452]
    strong = sorted(((int(h["p0"]), int(h["p1"]), "%.6G" % h["score"]) for h in res["hits"] if h["score"] > 5),
                    key=lambda t: -float(t[2]))
    assert strong == [(43, 69, "15.0855"), (43, 70, "11.6195"), (17, 70, "8.52062")]
    eng.close()


def _oracle_scan(names, seq, win_len, win_step, seed, min_bpp=0.001):
    from oracle import pyoracle
    from rmdetect_b200.synth import SynthFold
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
    fold = SynthFold(seed)
    starts = pyoracle.windows(len(seq), win_len, win_step)
    bpp = [pyoracle.synth_bpp(seq[s:s + win_len], seed) for s in starts]
    return pyoracle.scan_sequence(oms, oes, 0, seq, win_len, win_step, bpp, min_bpp, LN_UNKNOWN, LN_CUTOFF)


def test_synthetic_job_generated_on_device_matches_oracle():
    """Device-side genome + stand-in bpp are bit-identical to the host twins; the scan of a 30 kb
    synthetic chromosome matches the oracle hit for hit."""
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.scanner import ScanEngine, gc_of_codes
    from rmdetect_b200.synth import synth_bpp_coo, synth_codes, synth_sequence
    n, seed = 30000, 20261019
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    eng.ctx.set_screen_policy(0)                      # with prefix-screen tables (the default spares sequences this short)
    codes = synth_codes(n, seed)
    seq = synth_sequence(n, seed)
    gc = gc_of_codes(codes, "")
    cls = []
    for ev in eng.evalues:
        ev.select_gc(gc)
        cls.append(ev.selected)
    n_win, n_bpp, _ = eng.ctx.synth_job(n, 0, seed, seed, 150, 75, 0.001, LN_CUTOFF, gc_table(gc), cls)
    assert n_win == 399
    assert np.array_equal(eng.ctx.download_codes(n), codes)
    counts = eng.ctx.gc_counts(1)
    assert counts[0, :4].tolist() == np.bincount(codes, minlength=4).tolist() and counts[0, 4:].sum() == 0
    for w in (0, 1, 57, 398):
        s = w * 75 if w < 398 else n - 150
        a = synth_bpp_coo(seq[s:s + 150], seed)
        b = eng.ctx.download_bpp(w)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    res = eng.ctx.scan_resident(0, n_win)
    hits, tries, _ = _oracle_scan(MODEL_ORDER, seq, 150, 75, seed)
    assert res["tries"] == tries
    assert len(res["hits"]) == len(hits)
    for g, w in zip(res["hits"], hits):
        assert (g["model"], g["start"], g["start"] + g["wlen"], g["p0"], g["p1"], g["leaf"]) == \
               (w["model"], w["coords"][0], w["coords"][1], w["p"][0], w["p"][1], w["cfg"])
        assert (g["prob_model"], g["prob_rand"], g["score"], g["evalue"]) == \
               (w["prob_model"], w["prob_rand"], w["score"], w["evalue"])
    # splitting the window range must only lose the seam's duplicate removal
    a = eng.ctx.scan_resident(0, 200)
    b = eng.ctx.scan_resident(200, n_win)
    assert a["tries"][0] + b["tries"][0] == tries[0] and a["tries"][1] + b["tries"][1] == tries[1]
    assert a["n_raw"] + b["n_raw"] == res["n_raw"]
    eng.close()


def test_letters_outside_acgu_and_gc_counts():
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    s = list(synth_sequence(400, 5))
    for pos in (3, 150, 151, 152, 299):
        s[pos] = "N"
    s[200] = "Y"
    seq = "".join(s)
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    job = eng.build_job([("x", seq, None), ("y", seq[:90], None)], 150, 75, SynthFold(9), "", 0.001, LN_CUTOFF)
    eng.ctx.upload(job)
    counts = eng.ctx.gc_counts(2)
    assert counts[0].tolist()[:7] == [seq.count("A"), seq.count("C"), seq.count("G"), seq.count("U"), 0, 5, 1]
    assert counts[1, :4].sum() + counts[1, 5:].sum() == 90
    eng.close()


@pytest.mark.parametrize("name", ["KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0"])
def test_calibration_histograms(name):
    """rmd_calibrate on the reference's own seeded sample stream vs the reference's histograms."""
    from rmdetect_b200.calibrate import class_logs, draw_samples
    from rmdetect_b200.gcx import GC
    from rmdetect_b200.scanner import ScanEngine
    g = golden("calibrate")[name]
    m = load_model(name)
    L = len(m.nodes)
    n = min(g["samples"], 4 ** L)
    samples = draw_samples(n, L, g["seed"])
    eng = ScanEngine([m], [None], LN_UNKNOWN)
    classes = GC.get_classes(5, 15, 85)
    hist = np.zeros((len(classes), 512))
    n_ok, _ = eng.ctx.calibrate(0, samples, class_logs(classes), LN_CUTOFF, hist)
    for ci, want in enumerate(g["hists"]):
        got = {k / 10.0: v for k, v in enumerate(hist[ci]) if v != 0.0}
        assert sorted(got) == [k for k, _ in want]            # exactly the same score buckets
        for k, v in want:
            assert got[k] == pytest.approx(v, rel=1e-9)
    eng.close()


def test_one_megabase_counts_match_oracle():
    """1 Mb synthetic chromosome, all four models: tests_all, tests_pairs and the number of hits
    before duplicate removal equal the oracle's (13 332 windows, 9.46e8 placements)."""
    from oracle import pyoracle
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.scanner import ScanEngine, gc_of_codes
    from rmdetect_b200.synth import synth_codes, synth_sequence
    n, seed = 1_000_000, 77
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [None] * 4, LN_UNKNOWN)
    gc = gc_of_codes(synth_codes(n, seed), "")
    n_win, _, _ = eng.ctx.synth_job(n, 0, seed, seed, 150, 75, 0.001, LN_CUTOFF, gc_table(gc))
    res = eng.ctx.scan_resident(0, n_win)
    nh, tries = pyoracle.bench_scan([load_model(m) for m in MODEL_ORDER], synth_sequence(n, seed), 150, 75, 0, n_win,
                                    seed, os.cpu_count() or 1, 0.001, LN_UNKNOWN, LN_CUTOFF)
    assert res["tries"] == tries
    assert res["n_raw"] == nh
    assert tries[0] == n_win * 70939
    eng.close()


class _TableFold:
    """bpp provider for tests: window string -> (i, j, p) arrays."""

    def __init__(self, fn):
        self.fn = fn

    def coo(self, win):
        return self.fn(win)


def _scan_vs_oracle(names, seq, fold, min_bpp, win_len=150, win_step=75):
    from oracle import pyoracle
    from rmdetect_b200.scanner import ScanEngine
    eng = ScanEngine([load_model(n) for n in names], [load_evalues(n) for n in names], LN_UNKNOWN)
    res = eng.scan_job(eng.build_job([("s", seq, None)], win_len, win_step, fold, "", min_bpp, LN_CUTOFF))
    eng.close()
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
    starts = pyoracle.windows(len(seq), win_len, win_step)
    hits, tries, _ = pyoracle.scan_sequence(oms, oes, 0, seq, win_len, win_step, [fold.coo(seq[s:s + win_len]) for s in starts],
                                            min_bpp, LN_UNKNOWN, LN_CUTOFF)
    assert res["tries"] == tries
    assert len(res["hits"]) == len(hits)
    for g, w in zip(res["hits"], hits):
        assert (g["model"], g["start"], g["p0"], g["p1"], g["leaf"]) == (w["model"], w["coords"][0], w["p"][0], w["p"][1], w["cfg"])
        assert (g["prob_model"], g["prob_rand"], g["score"], g["evalue"]) == (w["prob_model"], w["prob_rand"], w["score"], w["evalue"])
    return res


@pytest.mark.parametrize("min_bpp", [0.0, -1.0])
def test_non_positive_threshold_takes_every_placement(min_bpp):
    """min_bpp_mean <= 0: the mean of no pairs (0.0) passes, every placement is scored (lib/scanner.py:143,211-214)."""
    from rmdetect_b200.synth import SynthFold, synth_sequence
    seq = synth_sequence(190, 31)
    res = _scan_vs_oracle(["KT_1.0", "TGA_1.0"], seq, SynthFold(4), min_bpp)
    assert res["tries"][0] == res["tries"][1]


def test_gate_mean_on_the_threshold():
    """Means that sit exactly on, one ulp under and one ulp over min_bpp_mean: the kernel's fast sum
    (value x multiplicity) must hand these to the reference-order sum."""
    from rmdetect_b200.synth import synth_bpp_coo, synth_sequence
    t = 0.001
    near = [t, np.nextafter(t, 0.0), np.nextafter(t, 1.0), 3 * t - 2 * np.nextafter(t, 1.0), 2 * t, t / 2, t / 3, 1e-5, 0.0015, 0.0005]

    def fold(win):
        i, j, p = synth_bpp_coo(win, 12)
        p = p.copy()
        rng = random.Random(len(win) * 7919 + sum(map(ord, win[:20])))
        for k in range(len(p)):
            p[k] = near[rng.randrange(len(near))]
        return i, j, p

    seq = synth_sequence(450, 8)
    _scan_vs_oracle(MODEL_ORDER, seq, _TableFold(fold), t)


def test_dense_matrix_window():
    """Every pair of the window stored (11 175 entries): list ranks beyond 8 bits, every trie node alive."""
    from rmdetect_b200.synth import synth_sequence

    def fold(win):
        n = len(win)
        i, j = np.triu_indices(n, 1)
        rng = np.random.default_rng(n)
        return i.astype(np.uint16), j.astype(np.uint16), rng.choice([2e-4, 8e-4, 1.2e-3, 5e-3], size=len(i))

    seq = synth_sequence(150, 3)
    res = _scan_vs_oracle(["CL_1.0", "GB_1.0"], seq, _TableFold(fold), 0.001)
    assert res["tries"][1] > 1000


def test_model_groups_and_two_node_tries():
    """Candidate generation by groups: gate tries of more than two nodes (CL, ARICH) take turns with the one resident
    duplicate bitmap, tries of two nodes (KT, GB, TGA) drop duplicates by the first-big-pair rule; dense matrices
    make every placement reachable from several entries."""
    from rmdetect_b200.synth import SynthFold, synth_sequence

    def fold(win):
        n = len(win)
        i, j = np.triu_indices(n, 1)
        rng = np.random.default_rng(n + 1)
        keep = rng.random(len(i)) < 0.35
        return i[keep].astype(np.uint16), j[keep].astype(np.uint16), rng.choice([3e-4, 9.99e-4, 1e-3, 2.5e-3, 0.3], size=int(keep.sum()))

    seq = synth_sequence(260, 12)
    names = ["CL_1.0", "KT_1.0", "ARICH_1.0", "GB_1.0", "CL_1.0", "TGA_1.0"]
    res = _scan_vs_oracle(names, seq, _TableFold(fold), 0.001)
    assert res["tries"][1] > 5000
    _scan_vs_oracle(["TGA_1.0", "GB_1.0", "KT_1.0"], seq, _TableFold(fold), 0.001)
    _scan_vs_oracle(names, synth_sequence(700, 13), SynthFold(5), 0.001)


@pytest.mark.parametrize("letters", ["ACGU", "AAACGUU", "GGGCCCAU", "ACGUN"])
def test_prefix_screens_do_not_change_results(letters):
    """Screens on (tables for every sequence), default (none for short sequences) and off: same hits as the oracle,
    for balanced and skewed GC tables and with a letter outside ACGU in the sequence."""
    from oracle import pyoracle
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold
    rng = random.Random(len(letters))
    seqs = [("a", "".join(rng.choice(letters) for _ in range(2400)), None),
            ("b", "".join(rng.choice("ACGU") for _ in range(700)), None)]
    names = MODEL_ORDER + ["ARICH_1.0"]
    fold = SynthFold(21)
    eng = ScanEngine([load_model(n) for n in names], [load_evalues(n) for n in names], LN_UNKNOWN)
    job = eng.build_job(seqs, 150, 75, fold, "", 0.001, LN_CUTOFF)
    runs = []
    for policy in (0, 1024, -1):
        eng.ctx.set_screen_policy(policy)
        r = eng.scan_job(job)
        runs.append((r["tries"], r["hits"].tobytes(), r["gate_pass"].tobytes(), r["raw_count"].tobytes()))
    eng.close()
    assert runs[0] == runs[1] == runs[2]
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
    hits = np.frombuffer(runs[0][1], dtype=r["hits"].dtype)
    at, tests = 0, [0, 0]
    for si, (_, seq, _) in enumerate(seqs):
        starts = pyoracle.windows(len(seq), 150, 75)
        want, tries, _ = pyoracle.scan_sequence(oms, oes, si, seq, 150, 75, [fold.coo(seq[s:s + 150]) for s in starts],
                                                0.001, LN_UNKNOWN, LN_CUTOFF)
        tests[0] += tries[0]; tests[1] += tries[1]
        for w in want:
            g = hits[at]; at += 1
            assert (g["seq"], g["model"], g["start"], g["p0"], g["p1"], g["leaf"]) == (si, w["model"], w["coords"][0], w["p"][0], w["p"][1], w["cfg"])
            assert (g["prob_model"], g["prob_rand"], g["score"], g["evalue"]) == (w["prob_model"], w["prob_rand"], w["score"], w["evalue"])
    assert at == len(hits) and runs[0][0] == tests


def test_full_size_chromosome_invariants():
    """BASELINE.json configs[2] at full size (10 Mb, 133 333 windows, 9.46e9 placements), where the oracle takes
    minutes: size-independent properties instead.  tests_all equals the odometer's closed form; the hit list does not
    depend on the prefix screens; cutting the window range changes nothing but the duplicate rule at the seams; the
    host-buffer call (compact and fp64 transport) returns the device-resident result."""
    from rmdetect_b200 import _native
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.scanner import ScanEngine, placements_in_window
    n, seed = 10_000_000, 20261019
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    gc = gc_table({b: math.log(0.25) for b in "ACGU"})
    n_win, n_bpp, _ = eng.ctx.synth_job(n, 0, seed, seed, 150, 75, 0.001, LN_CUTOFF, gc)
    assert n_win == 133333
    whole = eng.ctx.scan_resident(0, n_win)
    assert whole["tries"][0] == n_win * sum(placements_in_window(m, 150) for m in eng.models) == 9458509687
    assert whole["tries"][1] == int(whole["gate_pass"].sum()) and whole["n_raw"] == int(whole["raw_count"].sum())
    assert len(whole["hits"]) > 500000 and np.all(np.diff(whole["hits"]["window"]) >= 0)
    eng.ctx.set_screen_policy(-1)                                             # no prefix screens
    plain = eng.ctx.scan_resident(0, n_win)
    assert plain["tries"] == whole["tries"] and plain["hits"].tobytes() == whole["hits"].tobytes()
    eng.ctx.set_screen_policy(1024)
    cuts = [0, 40000, 40001, 99999, n_win]
    parts = [eng.ctx.scan_resident(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert [sum(p["tries"][k] for p in parts) for k in (0, 1)] == whole["tries"]
    assert sum(p["n_raw"] for p in parts) == whole["n_raw"]
    seams = sum(len(p["hits"]) for p in parts) - len(whole["hits"])           # duplicates across a cut survive
    assert 0 <= seams <= 3 * 40
    # the same job from host buffers, both transports
    off, bi, bj, bp = eng.ctx.download_job(n_win, n_bpp)
    job = dict(seq_off=np.array([0, n], dtype=np.int64), codes=eng.ctx.download_codes(n), gc_log=gc.reshape(1, -1),
               gc_class=np.full((1, 4), -1, dtype=np.int32), win_len=150, win_step=75, n_win=n_win, bpp_off=off,
               bpp_i=bi, bpp_j=bj, bpp_p=bp, min_bpp_mean=0.001, cutoff=LN_CUTOFF)
    ij, q = _native.compact_bpp(bi, bj, bp)
    for j in (job, dict(job, bpp_i=None, bpp_j=None, bpp_p=None, bpp_ij=ij, bpp_q=q)):
        r = eng.ctx.scan(j)
        assert r["tries"] == whole["tries"] and r["hits"].tobytes() == whole["hits"].tobytes()
    eng.close()


def test_sure_pass_rule_on_its_boundary():
    """A generating value of at least ceil(slots / mult) x min_bpp x (1 + 1e-12) skips the gate walk.  Worst case for
    that rule: every pair of the window present with a tiny value (all configurations run to their end, the count is
    the full number of MUST slots) and single spikes sitting exactly on, one ulp under and one ulp over the
    multiples of the threshold the rule uses (KT, GB, TGA: 2; CL: 15, 30, 60)."""
    from rmdetect_b200.synth import synth_sequence
    t = 0.001
    unit = t * (1.0 + 1e-12)
    spikes = []
    for need in (1, 2, 15, 30, 60):
        for v in (need * unit, need * t):
            spikes += [np.nextafter(v, 0.0), v, np.nextafter(v, 1.0)]

    def fold(win):
        n = len(win)
        i, j = np.triu_indices(n, 1)
        p = np.full(len(i), 1e-5)
        rng = np.random.default_rng(n + 7)
        at = rng.choice(len(i), size=12 * len(spikes), replace=False)
        p[at] = np.tile(np.array(spikes), 12)
        return i.astype(np.uint16), j.astype(np.uint16), p

    seq = synth_sequence(150, 41)
    res = _scan_vs_oracle(MODEL_ORDER, seq, _TableFold(fold), t)
    assert res["tries"][1] > 200


def test_malformed_bpp_list_is_rejected():
    """The kernel's own checks (the host layer sorts and validates; a C caller may not)."""
    from rmdetect_b200._native import NativeError
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    seq = synth_sequence(150, 3)

    def duplicate(job):
        job["bpp_i"][7], job["bpp_j"][7] = job["bpp_i"][6], job["bpp_j"][6]

    def negative(job):
        job["bpp_p"][5] = -job["bpp_p"][5]

    def lower_triangle(job):
        job["bpp_i"][9], job["bpp_j"][9] = job["bpp_j"][9], job["bpp_i"][9]

    for spoil in (duplicate, negative, lower_triangle):
        eng = ScanEngine([load_model(n) for n in MODEL_ORDER], [None] * 4, LN_UNKNOWN)
        job = eng.build_job([("s", seq, None)], 150, 75, SynthFold(1), "", 0.001, LN_CUTOFF)
        spoil(job)
        with pytest.raises(NativeError):
            eng.scan_job(job)
        eng.close()


def test_compact_transport_gives_identical_results():
    """rmd_scan with the 6-byte transport (ij + q words, values rebuilt on the device) == rmd_scan with fp64 lists."""
    from rmdetect_b200._native import compact_bpp
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    seqs = [("a", synth_sequence(3000, 41), None), ("b", synth_sequence(777, 42), None)]
    job = eng.build_job(seqs, 150, 75, SynthFold(7), "", 0.001, LN_CUTOFF)
    # make some values need the one-ulp correction (libm pow vs x * x)
    for k in range(0, len(job["bpp_p"]), 97):
        job["bpp_p"][k] = np.nextafter(job["bpp_p"][k], 1.0 if k % 2 else 0.0)
    want = eng.scan_job(job)
    ij, q = compact_bpp(job["bpp_i"], job["bpp_j"], job["bpp_p"])
    assert (q >> 30 != 1).sum() >= len(q) // 97
    cjob = dict(job, bpp_ij=ij, bpp_q=q)
    cjob["bpp_i"] = cjob["bpp_j"] = cjob["bpp_p"] = None
    got = eng.ctx.scan(cjob)
    assert got["tries"] == want["tries"] and len(want["hits"]) > 10
    assert got["hits"].tobytes() == want["hits"].tobytes()
    assert np.array_equal(got["gate_pass"], want["gate_pass"])
    # rmd_scan left the lists compact on the device (the scan kernel rebuilt the values itself): a resident rescan,
    # a rescan of a window range, and handing the lists back (which expands them) all see the caller's values
    again = eng.ctx.scan_resident(0, job["n_win"])
    assert again["hits"].tobytes() == want["hits"].tobytes() and again["tries"] == want["tries"]
    part = eng.ctx.scan_resident(5, 30)
    assert part["tries"][1] == int(want["gate_pass"][5:30].sum())
    for w in (0, 17, job["n_win"] - 1):
        a, b = int(job["bpp_off"][w]), int(job["bpp_off"][w + 1])
        i, j, p = eng.ctx.download_bpp(w)
        assert np.array_equal(i, job["bpp_i"][a:b]) and np.array_equal(j, job["bpp_j"][a:b]) and np.array_equal(p, job["bpp_p"][a:b])
    after = eng.ctx.scan_resident(0, job["n_win"])                 # now from the expanded lists
    assert after["hits"].tobytes() == want["hits"].tobytes()
    # the resident path too (rmd_upload + rmd_scan_resident)
    eng.ctx.upload(cjob)
    res = eng.ctx.scan_resident(0, job["n_win"])
    assert res["hits"].tobytes() == want["hits"].tobytes()
    eng.close()


def test_two_contexts_on_one_device_keep_their_own_models():
    """The model tables live in per-device __constant__ memory: a context must re-bind its own before every launch
    (an older context used after a newer one was created, a calibration context next to a scanner, a context that
    outlives the one created after it)."""
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    seq = synth_sequence(1200, 5)
    fold = SynthFold(3)
    a = ScanEngine([load_model(n) for n in MODEL_ORDER], [load_evalues(n) for n in MODEL_ORDER], LN_UNKNOWN)
    job_a = a.build_job([("s", seq, None)], 150, 75, fold, "", 0.001, LN_CUTOFF)
    want_a = a.scan_job(job_a)
    b = ScanEngine([load_model("ARICH_1.0"), load_model("KT_1.0")], [load_evalues("ARICH_1.0"), None], LN_UNKNOWN)
    job_b = b.build_job([("s", seq, None)], 150, 75, fold, "", 0.001, LN_CUTOFF)
    want_b = b.scan_job(job_b)
    assert want_a["tries"] != want_b["tries"]
    for _ in range(2):                                    # interleaved: each call finds the other context's tables bound
        got_a = a.scan_job(job_a)
        assert got_a["tries"] == want_a["tries"] and got_a["hits"].tobytes() == want_a["hits"].tobytes()
        got_b = b.ctx.scan_resident(0, job_b["n_win"])
        assert got_b["tries"] == want_b["tries"] and got_b["hits"].tobytes() == want_b["hits"].tobytes()
    c = ScanEngine([load_model("TGA_1.0")], [None], LN_UNKNOWN)   # a third one, destroyed while the others live on
    c.scan_job(c.build_job([("s", seq[:300], None)], 150, 75, fold, "", 0.001, LN_CUTOFF))
    c.close()
    got_a = a.ctx.scan_resident(0, job_a["n_win"])
    assert got_a["hits"].tobytes() == want_a["hits"].tobytes()
    b.close()
    got_a = a.scan_job(job_a)
    assert got_a["hits"].tobytes() == want_a["hits"].tobytes()
    a.close()


def test_failed_scan_leaves_no_resident_job_and_refused_models_change_nothing():
    from rmdetect_b200._native import NativeError
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    seq = synth_sequence(600, 9)
    eng = ScanEngine([load_model(n) for n in MODEL_ORDER], [None] * 4, LN_UNKNOWN)
    good = eng.build_job([("s", seq, None)], 150, 75, SynthFold(1), "", 0.001, LN_CUTOFF)
    want = eng.scan_job(good)
    bad = dict(good, bpp_p=good["bpp_p"].copy())
    bad["bpp_p"][5] = -1.0
    with pytest.raises(NativeError):
        eng.scan_job(bad)
    with pytest.raises(NativeError, match="no resident job"):      # not a rescan of half-staged lists
        eng.ctx.scan_resident(0, good["n_win"])
    # descriptors that are refused leave models and job as they were
    eng.ctx.upload(good)
    flats = list(eng.flats)
    import copy
    broken = copy.copy(flats[0])
    broken.gate_cfg = np.array(flats[0].gate_cfg, dtype=np.int32).copy()
    broken.gate_cfg[-1] += 5                                         # reaches past the gate pairs
    with pytest.raises(NativeError):
        eng.ctx.set_models([broken] + flats[1:])
    again = eng.ctx.scan_resident(0, good["n_win"])
    assert again["tries"] == want["tries"] and again["hits"].tobytes() == want["hits"].tobytes()
    eng.close()


def test_window_shorter_than_a_chain_with_the_gate_open():
    """[0, 0] is tried even when the window is shorter than a chain (lib/scanner.py:137-141); with min_bpp_mean <= 0 the
    reference would then index past the sequence (IndexError).  Here such a placement counts in tests_all / tests_pairs
    and yields no hit; no letter beyond the window is read."""
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold
    eng = ScanEngine([load_model(n) for n in MODEL_ORDER], [None] * 4, LN_UNKNOWN)
    # the second sequence would supply the letters a stray read would see: a perfect TGA/KT-looking neighbourhood
    job = eng.build_job([("tiny", "GGA", None), ("next", "GGACUGAAGCCGAUGAGGCGAAACCC" * 3, None)], 150, 75, SynthFold(2), "", 0.0, LN_CUTOFF)
    res = eng.scan_job(job)
    assert res["gate_pass"][0].tolist() == [1, 1, 1, 1]
    assert not np.any(res["hits"]["seq"] == 0)
    eng.close()


def test_tuning_switches_in_the_environment_are_ignored(monkeypatch):
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold, synth_sequence
    seq = synth_sequence(900, 2)

    def run():
        eng = ScanEngine([load_model(n) for n in MODEL_ORDER], [load_evalues(n) for n in MODEL_ORDER], LN_UNKNOWN)
        r = eng.scan_job(eng.build_job([("s", seq, None)], 150, 75, SynthFold(8), "", 0.001, LN_CUTOFF))
        eng.close()
        return r
    want = run()
    for k, v in (("RMD_DEBUG", "1"), ("RMD_TEAMS", "3"), ("RMD_NO_SCREEN", "1"), ("RMD_LOCAL_STACK", "1"),
                 ("RMD_NO_FUSED_EXPAND", "1"), ("RMD_CTAS_PER_SM", "2")):
        monkeypatch.setenv(k, v)
    got = run()
    assert len(want["hits"]) > 0 and got["hits"].tobytes() == want["hits"].tobytes() and got["n_launches"] == want["n_launches"]


@pytest.mark.parametrize("screens", [0, -1])
def test_windows_with_other_letters_take_the_shared_memory_kernel(screens):
    """N runs and scattered IUPAC codes in a 95 kb sequence (every such letter has its own log-frequency, lib/gcx.py:4-13,
    and no CPT entry, lib/node.py:215-223): the OTHER launch of scan_fast_kernel scans those windows; hit for hit, bit for
    bit the oracle's result, with the prefix screens on and off."""
    from oracle import pyoracle
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold
    rng = random.Random(77)
    letters = [rng.choice("ACGU") for _ in range(95000)]
    for _ in range(40):                                    # runs of N: 1 .. 200 letters
        at, ln = rng.randrange(len(letters)), rng.choice([1, 2, 5, 30, 200])
        letters[at:at + ln] = "N" * len(letters[at:at + ln])
    for _ in range(150):                                   # single ambiguity codes
        letters[rng.randrange(len(letters))] = rng.choice("RYKMSWN")
    seqs = [("chr", "".join(letters), None), ("short", "ACGUNACGGGCAUNNNNACGAUCGAUCGGCUAGCUAGGCUA" * 5, None)]
    names = MODEL_ORDER
    fold = SynthFold(31)
    eng = ScanEngine([load_model(n) for n in names], [load_evalues(n) for n in names], LN_UNKNOWN)
    eng.ctx.set_screen_policy(screens)
    job = eng.build_job(seqs, 150, 75, fold, "", 0.001, LN_CUTOFF)
    r = eng.scan_job(job)
    eng.close()
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
    hits = r["hits"]
    at, tests = 0, [0, 0]
    for si, (_, seq, _) in enumerate(seqs):
        starts = pyoracle.windows(len(seq), 150, 75)
        want, tries, _ = pyoracle.scan_sequence(oms, oes, si, seq, 150, 75, [fold.coo(seq[s:s + 150]) for s in starts],
                                                0.001, LN_UNKNOWN, LN_CUTOFF)
        tests[0] += tries[0]; tests[1] += tries[1]
        for w in want:
            g = hits[at]; at += 1
            assert (g["seq"], g["model"], g["start"], g["p0"], g["p1"], g["leaf"]) == (si, w["model"], w["coords"][0], w["p"][0], w["p"][1], w["cfg"])
            assert (g["prob_model"], g["prob_rand"], g["score"], g["evalue"]) == (w["prob_model"], w["prob_rand"], w["score"], w["evalue"])
    assert at == len(hits) > 100 and r["tries"] == tests


def test_windows_of_161_to_256_letters_take_the_shared_memory_kernel():
    """A larger --win-len and whole-sequence mode: windows up to 160 letters go to the common launch, 161..256 to the
    launch staged with a larger bit matrix, longer ones to the generic kernel, with and without letters other than
    ACGU; every hit equals the oracle's, bit for bit."""
    from oracle import pyoracle
    from rmdetect_b200.scanner import ScanEngine
    from rmdetect_b200.synth import SynthFold
    rng = random.Random(19)
    mk = lambda n, letters="ACGU": "".join(rng.choice(letters) for _ in range(n))
    names = MODEL_ORDER + ["ARICH_1.0"]
    fold = SynthFold(41)
    eng = ScanEngine([load_model(n) for n in names], [load_evalues(n) for n in names], LN_UNKNOWN)
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
    for W, step, seqs in ((230, 90, [("a", mk(2000), None), ("b", mk(140), None), ("c", mk(700, "ACGUN"), None), ("d", mk(256), None)]),
                          (0, 0, [("w%d" % n, mk(n, "ACGU" if n % 2 else "ACGUR"), None) for n in (120, 160, 161, 200, 255, 256, 257, 300)])):
        job = eng.build_job(seqs, W, step, fold, "", 0.001, LN_CUTOFF)
        r = eng.scan_job(job)
        hits = r["hits"]
        at, tests = 0, [0, 0]
        for si, (_, seq, _) in enumerate(seqs):
            Wo, so = (W, step) if W else (len(seq), len(seq))      # whole-sequence mode = one window that is the sequence
            starts = pyoracle.windows(len(seq), Wo, so)
            want, tries, _ = pyoracle.scan_sequence(oms, oes, si, seq, Wo, so, [fold.coo(seq[s:s + Wo]) for s in starts],
                                                    0.001, LN_UNKNOWN, LN_CUTOFF)
            tests[0] += tries[0]; tests[1] += tries[1]
            for w in want:
                g = hits[at]; at += 1
                assert (g["seq"], g["model"], g["start"], g["p0"], g["p1"], g["leaf"]) == (si, w["model"], w["coords"][0], w["p"][0], w["p"][1], w["cfg"])
                assert (g["prob_model"], g["prob_rand"], g["score"], g["evalue"]) == (w["prob_model"], w["prob_rand"], w["score"], w["evalue"])
        assert at == len(hits) > 20 and r["tries"] == tests
    eng.close()
