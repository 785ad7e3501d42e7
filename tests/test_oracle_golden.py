"""Pin the CPU oracle (oracle/rmd_oracle.c) to the reference.

Every expectation here was produced by the reference's own Python code
(tests/golden/make_golden.py) or ships with the reference (cluster_*.res,
arich.out, the two real dot.ps matrices).  The CUDA path is then compared with
this oracle in tests/test_gpu_parity.py.
"""
import math
import os
import random

import numpy as np
import pytest

from helpers import (LN_CUTOFF, LN_UNKNOWN, MODEL_FILES, MODEL_ORDER, REF_DATA, evalues_path, golden,
                     load_model)
from oracle import pyoracle
from rmdetect_b200.candidates import decode_candidate, encode_chains
from rmdetect_b200.dotps import read_dot_ps
from rmdetect_b200.gcx import GC
from rmdetect_b200.synth import SynthFold, synth_bpp_coo


@pytest.fixture(scope="module")
def omodels():
    return {n: pyoracle.OracleModel(load_model(n)) for n in MODEL_FILES}


@pytest.fixture(scope="module")
def oevalues():
    return {n: pyoracle.OracleEValues(evalues_path(n)) for n in MODEL_FILES}


@pytest.mark.parametrize("name", sorted(MODEL_FILES))
def test_eval_vectors(omodels, name):
    om = omodels[name]
    n_ok = 0
    for case in golden("eval_vectors")[name]:
        ok, seqs, poss, pm, pr = om.eval(case["seq"], case["p"], case["gc"], case["unknown_prob"], case["cutoff"])
        assert ok == case["ok"]
        if ok:
            n_ok += 1
            assert seqs == case["seqs"] and poss == case["poss"]
            assert pm == case["prob_model"] and pr == case["prob_rand"]      # bit-exact fp64
    assert n_ok > 100


def test_small_functions(oevalues):
    sv = golden("small_vectors")
    evs = [oevalues[n] for n in sv["models"]]
    for case in sv["gc"]:
        assert pyoracle.gc_compute(case["seq"]) == case["gc"]
        assert [ev.select_gc(case["gc"]) for ev in evs] == case["selected"]
    for case in sv["evalue"]:
        ev = evs[case["model"]]
        ev.set_selected(case["selected"])
        assert ev.compute(case["score"]) == case["evalue"]


def test_synth_twin_matches_numpy():
    rng = random.Random(5)
    for n in (0, 3, 5, 17, 64, 150, 151, 203):
        seq = "".join(rng.choice("ACGU") for _ in range(n))
        if n == 150:
            seq = seq[:70] + "N" + seq[71:]
        a = synth_bpp_coo(seq, 20261019)
        b = pyoracle.synth_bpp(seq, 20261019)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def scan_case_with_oracle(case, omodels, oevalues):
    fold = SynthFold(case["seed"])
    oms = [omodels[n] for n in case["models"]]
    oes = [oevalues[n] for n in case["models"]]
    cands, tries = [], [0, 0]
    for si, (key, seq, cols) in enumerate(case["seqs"]):
        if case["win_len"] == 0 or case["win_step"] == 0:
            wins = [seq]
        else:
            starts = pyoracle.windows(len(seq), case["win_len"], case["win_step"])
            wins = [seq[s:s + case["win_len"]] for s in starts]
        bpp = [fold.coo(w) for w in wins]
        hits, t, _ = pyoracle.scan_sequence(oms, oes, si, seq, case["win_len"], case["win_step"], bpp,
                                            case["min_bpp"], math.log(case["unknown"]), math.log(case["cutoff"]))
        tries[0] += t[0]
        tries[1] += t[1]
        for h in hits:
            h["key"] = key
            h["cols"] = [[None if p is None else (p if cols is None else cols[p]) for p in ch] for ch in h["pos"]]
        cands.extend(hits)
    return cands, tries


@pytest.mark.parametrize("idx", range(15))
def test_scan_cases(omodels, oevalues, idx):
    case = golden("scan_cases")[idx]
    cands, tries = scan_case_with_oracle(case, omodels, oevalues)
    assert tries == case["tries"]
    assert len(cands) == len(case["cands"])
    for got, want in zip(cands, case["cands"]):
        assert case["models"][got["model"]] == want["model"]
        assert got["key"] == want["key"] and got["coords"] == want["coords"]
        assert got["seqs"] == want["seqs"] and got["pos"] == want["pos"] and got["cols"] == want["cols"]
        assert got["score"] == want["score"] and got["evalue"] == want["evalue"]


def test_scan_case_count():
    assert len(golden("scan_cases")) == 15


def test_tests_all_closed_form_lys_and_arich():
    # tests_all is independent of the bpp matrix
    total = 0
    for key, seq, _ in golden("seqsets")["lys.stk"]:
        nwin = len(pyoracle.windows(len(seq), 150, 75))
        assert all(len(seq) >= 150 for _ in [0])
        total += nwin * 70939
    assert total == 6952022
    from rmdetect_b200.scanner import placements_in_window
    m = load_model("ARICH_1.0")
    assert sum(placements_in_window(m, min(150, len(s))) for _, s, _ in golden("seqsets")["arich_test.stk"]) == 706346


def _res_lines(path):
    with open(path) as fh:
        return [ln.rstrip("\n") for ln in fh if ln.strip() and not ln.startswith("#")]


def _rescore(line, seqs_by_key, omodels, oevalues, models):
    cand = decode_candidate(line, models)
    name = cand.model.full_name()
    seq, cols = seqs_by_key[cand.key]
    s, e = cand.coords
    win = seq[s:e]
    ptr = [next(p for p in chain if p is not None) - s for chain in cand.chains_pos]
    gc = GC.compute(seq)
    ok, seqs, poss, pm, pr = omodels[name].eval(win, ptr, gc, LN_UNKNOWN, LN_CUTOFF)
    assert ok
    score = (pm - pr) / math.log(2.0)
    ev = oevalues[name]
    ev.select_gc(gc)
    evalue = ev.compute(score)
    fields = line.split("\t")
    assert "%.5E" % score == fields[4]
    assert "%.5E" % evalue == fields[5]
    assert ";".join(seqs) == fields[8]
    abs_pos = [[None if p is None else p + s for p in ch] for ch in poss]
    assert encode_chains(abs_pos) == fields[9]
    abs_cols = [[None if p is None else (p if cols is None else cols[p]) for p in ch] for ch in abs_pos]
    assert encode_chains(abs_cols) == fields[10]


def test_rescore_shipped_lysine_results(omodels, oevalues):
    # 120 hit lines of cluster_001..008_*.res (reference repo root), lys.stk, W=150/step 75
    seqs = {k: (s, c) for k, s, c in golden("seqsets")["lys.stk"]}
    models = [load_model(n) for n in MODEL_ORDER]
    n = 0
    for f in sorted(os.listdir(REF_DATA)):
        if f.startswith("cluster_") and f.endswith(".res"):
            for line in _res_lines(os.path.join(REF_DATA, f)):
                _rescore(line, seqs, omodels, oevalues, models)
                n += 1
    assert n == 120


def test_rescore_shipped_arich_results(omodels):
    # 204 hit lines of examples/arich/single_alignment/arich.out: scores and gap placement.
    # (Its E-value column was computed with a table that is not in the repository.)
    seqs = {k: (s, c) for k, s, c in golden("seqsets")["arich_test.stk"]}
    models = [load_model("ARICH_1.0")]
    lines = _res_lines(os.path.join(REF_DATA, "arich", "arich.out"))
    assert len(lines) == 204
    for line in lines:
        cand = decode_candidate(line, models)
        seq, cols = seqs[cand.key]
        s, e = cand.coords
        ptr = [next(p for p in chain if p is not None) - s for chain in cand.chains_pos]
        ok, sq, poss, pm, pr = omodels["ARICH_1.0"].eval(seq[s:e], ptr, GC.compute(seq), LN_UNKNOWN, LN_CUTOFF)
        assert ok
        fields = line.split("\t")
        assert "%.5E" % ((pm - pr) / math.log(2.0)) == fields[4]
        assert ";".join(sq) == fields[8]
        assert encode_chains([[None if p is None else p + s for p in ch] for ch in poss]) == fields[9]


def test_real_bpp_window_lysine(omodels, oevalues):
    # dot.ps at the reference root = RNAfold 2.5.1 matrix of window 32-182 of ABCZ01000009.1/166331-166512
    key = "ABCZ01000009.1/166331-166512"
    seq = next(s for k, s, _ in golden("seqsets")["lys.stk"] if k == key)
    wseq, i, j, p = read_dot_ps(os.path.join(REF_DATA, "lys_dot.ps"))
    assert list(pyoracle.windows(len(seq), 150, 75)) == [0, 32]
    assert wseq == seq[32:182]
    empty = (np.zeros(0, np.uint16), np.zeros(0, np.uint16), np.zeros(0, np.float64))
    oms = [omodels[n] for n in MODEL_ORDER]
    oes = [oevalues[n] for n in MODEL_ORDER]
    hits, tries, tpw = pyoracle.scan_sequence(oms, oes, 0, seq, 150, 75, [empty, (i, j, p)],
                                              0.001, LN_UNKNOWN, LN_CUTOFF)
    assert tpw.tolist() == [[70939, 0], [70939, 1806]]
    got = {(MODEL_ORDER[h["model"]], h["pos"][0][0], h["pos"][1][0]): h for h in hits}
    kt = got[("KT_1.0", 164, 81)]
    assert ("%.5E" % kt["score"], "%.5E" % kt["evalue"]) == ("1.04455E+01", "3.04000E-05")
    tga = got[("TGA_1.0", 40, 56)]
    assert ("%.5E" % tga["score"], "%.5E" % tga["evalue"]) == ("8.21393E+00", "1.04900E-03")
    tga2 = got[("TGA_1.0", 67, 175)]
    assert "%.5E" % tga2["score"] == "9.52052E+00"


def test_real_bpp_window_arich(omodels):
    # examples/arich/single_alignment/dot.ps = RNAfold 1.8.4 matrix of CP000686.1/1863685-1863583 (L=103)
    wseq, i, j, p = read_dot_ps(os.path.join(REF_DATA, "arich", "dot.ps"))
    key = "CP000686.1/1863685-1863583"
    seq = next(s for k, s, _ in golden("seqsets")["arich_test.stk"] if k == key)
    assert wseq == seq and len(seq) == 103
    hits, tries, _ = pyoracle.scan_sequence([omodels["ARICH_1.0"]], None, 0, seq, 150, 75, [(i, j, p)],
                                            0.001, LN_UNKNOWN, LN_CUTOFF)
    assert tries == [9506, 452]
    strong = sorted(((h["p"][0], h["p"][1], "%.6G" % h["score"]) for h in hits if h["score"] > 5), key=lambda t: -float(t[2]))
    assert strong == [(43, 69, "15.0855"), (43, 70, "11.6195"), (17, 70, "8.52062")]


@pytest.mark.parametrize("name", ["KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0"])
def test_calibration_histograms(omodels, name):
    # rmcalibrate.py body run by the reference with random.seed(SEED): same sample stream here
    g = golden("calibrate")[name]
    m = load_model(name)
    L = len(m.nodes)
    random.seed(g["seed"])
    n = min(g["samples"], 4 ** L)
    samples = ["".join("ACGU"[random.randint(0, 3)] for _ in range(L)) for _ in range(n)]
    hist, nok = pyoracle.calibrate(omodels[name], samples, GC.get_classes(5, 15, 85), LN_UNKNOWN, LN_CUTOFF)
    for ci, want in enumerate(g["hists"]):
        got = {k / 10.0: v for k, v in enumerate(hist[ci]) if v != 0.0}
        assert sorted(got) == [k for k, _ in want]
        for k, v in want:
            assert got[k] == pytest.approx(v, rel=1e-12)
