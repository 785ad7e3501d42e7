"""fold.RNAfold — the batched ViennaRNA driver — against stand-in `RNAfold` executables that write what the real
binary writes (`dot.ps` in the working directory): parsing, concurrency in private directories, error behaviour,
and (on the GPU) a scan fed by the driver."""
import os
import stat
import sys
import textwrap

import numpy as np
import pytest

from helpers import LN_CUTOFF, LN_UNKNOWN, MODEL_ORDER, REF_DATA, golden, load_evalues, load_model
from rmdetect_b200.dotps import read_dot_ps
from rmdetect_b200.fold import FoldError, RNAfold
from rmdetect_b200.synth import SynthFold, synth_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FAKE = textwrap.dedent('''\
    #!%(python)s
    # stand-in for `RNAfold -p`: reads the sequence from stdin, writes rna.ps and dot.ps into the working directory
    import math, os, sys, time
    sys.path.insert(0, %(root)r)
    from rmdetect_b200.synth import SynthFold
    assert "-p" in sys.argv[1:], sys.argv
    lines = sys.stdin.read().split("\\n")
    seq = lines[0].strip()
    if seq.startswith("FAIL"):
        sys.stderr.write("no luck\\n"); sys.exit(3)
    if os.path.exists("dot.ps"):
        sys.stderr.write("working directory is not private\\n"); sys.exit(4)
    time.sleep(0.05)
    if seq.startswith("NOPS"):
        sys.exit(0)
    i, j, p = SynthFold(%(seed)d).coo(seq)
    with open("rna.ps", "w") as fh:
        fh.write("%%!PS-Adobe-3.0 EPSF-3.0\\n%% data start here\\nshowpage\\n")
    with open("dot.ps", "w") as fh:
        fh.write("%%!PS-Adobe-3.0 EPSF-3.0\\n/ubox { bogus } bind def\\n/sequence { (\\\\\\n" + seq + "\\\\\\n) } def\\n%%data starts here\\n")
        for a, b, v in zip(i.tolist(), j.tolist(), p.tolist()):
            fh.write("%%d %%d %%.9f ubox\\n" %% (a + 1, b + 1, round(math.sqrt(v) * 1e9) / 1e9))
        fh.write("1 5 0.9500000 lbox\\nshowpage\\nend\\n%%%%EOF\\n")
    sys.stdout.write(seq + "\\n" + "." * len(seq) + " (-1.00)\\n")
''')


@pytest.fixture(scope="module")
def fake_rnafold(tmp_path_factory):
    path = tmp_path_factory.mktemp("bin") / "RNAfold"
    path.write_text(FAKE % dict(python=sys.executable, root=ROOT, seed=33))
    path.chmod(path.stat().st_mode | stat.S_IXUSR)
    return str(path)


def _expected(seq, seed=33):
    i, j, p = SynthFold(seed).coo(seq)
    roots = [float("%.9f" % (round(np.sqrt(v) * 1e9) / 1e9)) for v in p.tolist()]
    return i, j, np.array([r ** 2.0 for r in roots])          # lib/rnatools.py:189-195


def test_driver_returns_what_the_fold_wrote(fake_rnafold):
    seq = synth_sequence(150, 2)
    drv = RNAfold(exe=fake_rnafold, workers=1)
    i, j, p = drv.coo(seq)
    ei, ej, ep = _expected(seq)
    assert len(i) > 300 and np.array_equal(i, ei) and np.array_equal(j, ej) and np.array_equal(p, ep)
    b = drv.bpprobs(seq, "")
    assert b.prob_pair(int(i[7]), int(j[7])) == p[7] == b.prob_pair(int(j[7]), int(i[7]))
    assert drv.folds == 2


def test_many_windows_fold_concurrently_in_private_directories(fake_rnafold):
    seq = synth_sequence(600, 5)
    wins = [seq[s:s + 150] for s in range(0, 451, 75)] + [seq[:150]]          # the last one repeats the first
    drv = RNAfold(exe=fake_rnafold, workers=4)
    got = drv.coo_many(wins)
    assert drv.folds == len(wins) - 1                                         # equal windows are folded once
    for w, (i, j, p) in zip(wins, got):
        ei, ej, ep = _expected(w)
        assert np.array_equal(i, ei) and np.array_equal(j, ej) and np.array_equal(p, ep)


def test_real_dot_ps_through_the_driver(tmp_path):
    """The reference's own RNAfold 2.5.1 output (dot.ps of lys window 32-182) handed out by a stand-in binary."""
    ps = os.path.join(REF_DATA, "lys_dot.ps")
    exe = tmp_path / "RNAfold"
    exe.write_text("#!/bin/sh\ncat > /dev/null\ncp %s dot.ps\n" % ps)
    exe.chmod(exe.stat().st_mode | stat.S_IXUSR)
    wseq, ei, ej, ep = read_dot_ps(ps)
    i, j, p = RNAfold(exe=str(exe)).coo(wseq)
    assert len(i) == 637 and np.array_equal(i, ei) and np.array_equal(j, ej) and np.array_equal(p, ep)
    with pytest.raises(FoldError, match="describes 150 letters"):
        RNAfold(exe=str(exe)).coo(wseq[:100])


def test_failures_are_loud(fake_rnafold, tmp_path, monkeypatch):
    with pytest.raises(FoldError, match="failed \\(3\\)"):
        RNAfold(exe=fake_rnafold).coo("FAILACGU")
    with pytest.raises(FoldError, match="wrote no dot.ps"):
        RNAfold(exe=fake_rnafold).coo("NOPSACGU")
    with pytest.raises(FoldError, match="cannot run"):
        RNAfold(exe=str(tmp_path / "missing")).coo("ACGU")
    monkeypatch.delenv("RMDETECT_RNAFOLD", raising=False)
    monkeypatch.setenv("PATH", str(tmp_path))
    with pytest.raises(FoldError, match="RNAfold not found"):
        RNAfold().coo("ACGU")


@pytest.mark.gpu
def test_scan_fed_by_the_driver(fake_rnafold):
    """Scanner with rna_tools = the driver: same candidates as with the matrices handed over directly."""
    from helpers import Cfg, SeqList, fresh_models
    from rmdetect_b200.scanner import Scanner

    class Direct:
        def coo(self, win):
            return _expected(win)

    key, seq = "s", synth_sequence(420, 9)
    out = []
    for provider in (RNAfold(exe=fake_rnafold, workers=4), Direct()):
        models, evs = fresh_models(MODEL_ORDER)
        sc = Scanner(Cfg(models, evs))
        sc.rna_tools = provider
        cands, tries = sc.scan(SeqList([(key, seq, None)]))
        out.append((tries, [(c.model.full_name(), tuple(c.coords), c.chains_pos, c.score, c.evalue) for c in cands]))
    assert out[0] == out[1] and out[0][0][1] > 0 and len(out[0][1]) > 0
