"""rmbuild's probability tables (rmdetect_b200/cptbuild.py) against the reference's own lib/bayes.py + lib/model.py run on
its shipped training rows (tests/golden/bayes_tables.json.gz, make_golden.py: bayes_tables).  The counting is the GPU's
(`rmd_cpt_counts`); on the CPU a numpy stand-in for that one call exercises the host arithmetic."""
import os

import numpy as np
import pytest

from helpers import REF_DATA, golden
from rmdetect_b200 import cptbuild

CASES = ["two_alignments", "edit_network"]


class NumpyCounts:
    """Test stand-in for Context.cpt_counts (same contract as include/rmdetect_b200.h: rmd_cpt_counts)."""

    def cpt_counts(self, letters, K, tables, run, n_runs):
        k3 = K ** 3
        counts = np.zeros((n_runs, len(tables), k3), dtype=np.uint32)
        first = np.full((len(tables), k3), 0x7fffffff, dtype=np.int32)
        for t, (a, b0, b1) in enumerate(np.asarray(tables).tolist()):
            cell = letters[:, a].astype(np.int64)
            if b0 >= 0:
                cell = cell + K * letters[:, b0]
            if b1 >= 0:
                cell = cell + K * K * letters[:, b1]
            np.add.at(counts[:, t, :], (run, cell), 1)
            np.minimum.at(first[t], cell, np.arange(len(cell), dtype=np.int32))
        return counts, first


def check(case, ctx):
    g = golden("bayes_tables")[case]
    rows = cptbuild.load_data(os.path.join(REF_DATA, "arich_build", case + ".data.gz"))
    assert len(rows) == g["rows"]
    nodes = [(nid, conds) for nid, conds in g["nodes"]]
    probs = cptbuild.estimate_tables(rows, nodes, ctx)
    for nid, _ in nodes:
        want = g["probs"][str(nid)]
        assert list(probs[nid].keys()) == list(want.keys()), nid            # key order = the reference's dict order
        for key, d in want.items():
            assert list(probs[nid][key].items()) == list(d.items()), (nid, key)   # letters in order, floats equal
    assert [l for l in cptbuild.prob_lines(nodes, probs).split("\n") if l.startswith("PROB")] == g["prob_lines"]
    return probs, g


@pytest.mark.parametrize("case", CASES)
def test_tables_equal_the_reference(case):
    check(case, NumpyCounts())


def test_shipped_model_agrees_to_its_three_decimals():
    """The `.model` next to the rows was written by the authors' own rmbuild run (Python 2, unrounded weights): every
    probability it prints is reproduced to the printed precision."""
    probs, g = check("two_alignments", NumpyCounts())
    seen = 0
    for line in g["shipped_model"].split("\n"):
        if not line.startswith("PROB"):
            continue
        t = line.split("\t")
        nid, key = int(t[1]), t[2]
        for tok in t[3:]:
            ch, val = tok.split(":")
            if val.strip() != "-":
                assert abs(probs[nid][key][ch] - float(val)) <= 0.0021, (nid, key, ch)
                seen += 1
    assert seen > 80


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_tables_on_the_gpu(case):
    from rmdetect_b200 import _native
    ctx = _native.Context(0)
    check(case, ctx)
    # the counting call itself against the stand-in, on rows with three weight runs and a 6-letter alphabet
    rng = np.random.default_rng(5)
    letters = rng.integers(0, 6, size=(5000, 9)).astype(np.uint8)
    run = np.sort(rng.integers(0, 3, size=5000)).astype(np.int32)
    tables = np.array([[0, -1, -1], [3, 1, -1], [8, 2, 5], [4, 4, 4]], dtype=np.int32)
    got = ctx.cpt_counts(letters, 6, tables, run, 3)
    want = NumpyCounts().cpt_counts(letters, 6, tables, run, 3)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    ctx.close()
