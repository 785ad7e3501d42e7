"""The package's own .model/.evalues readers against structures dumped from the reference."""
import math

import pytest

from helpers import MODEL_FILES, golden, load_evalues, load_model
from rmdetect_b200.gcx import GC
from rmdetect_b200.modelfile import ModelFormatError, parse_model


@pytest.mark.parametrize("name", sorted(MODEL_FILES))
def test_model_matches_reference_structures(name):
    ref = golden("models")[name]
    m = load_model(name)
    assert (m.name, m.version) == (ref["name"], ref["version"])
    assert m.full_name() == name
    assert [m.chains_length[c] for c in range(m.chains_count)] == ref["chains_length"]
    assert m.symmetric == ref["symmetric"]
    assert m.sep_min == ref["sep_min"] and m.sep_max == ref["sep_max"]
    assert [n.id for n in m.order] == ref["order"]
    assert m.OFFSETS == ref["OFFSETS"]
    # the gate sums configurations in this very order (a Python set's iteration order)
    assert [[list(p) for p in cfg] for cfg in m.pairs_list] == ref["pairs_list"]
    for n, rn in zip(m.nodes, ref["nodes"]):
        assert (n.id, n.chain, n.ndx, n.conds) == (rn["id"], rn["chain"], rn["ndx"], rn["conds"])
        assert n.probs == rn["probs"]          # natural logs, bit for bit


def test_placement_counts_match_survey():
    # placements per W=150 window
    from rmdetect_b200.scanner import placements_in_window
    want = {"CL_1.0": 19600, "TGA_1.0": 10731, "GB_1.0": 20448, "KT_1.0": 20160, "ARICH_1.0": 20880}
    for name, n in want.items():
        assert placements_in_window(load_model(name), 150) == n


def test_gc_classes_and_compute():
    sv = golden("small_vectors")
    assert GC.get_classes(5, 15, 85) == sv["gc_classes"]
    assert len(sv["gc_classes"]) == 165
    for case in sv["gc"]:
        assert GC.compute(case["seq"]) == case["gc"]


def test_select_gc_and_compute_match_reference():
    sv = golden("small_vectors")
    evs = [load_evalues(n) for n in sv["models"]]
    for case in sv["gc"]:
        assert [ev.select_gc(case["gc"]) for ev in evs] == case["selected"]
    for case in sv["evalue"]:
        ev = evs[case["model"]]
        ev.selected = case["selected"]
        assert ev.compute(case["score"]) == case["evalue"]


def test_evalues_tables_structure():
    # 165 classes, floor 4^-L, first row all 1.0, columns non-increasing
    floors = {"KT_1.0": 4.0 ** -16, "GB_1.0": 4.0 ** -14, "CL_1.0": 4.0 ** -20, "ARICH_1.0": 4.0 ** -11}
    for name in MODEL_FILES:
        d = load_evalues(name).dense()
        assert d.shape[0] == 165
        assert (d[:, 0] == 1.0).all()
        assert (d[:, 1:] <= d[:, :-1]).all()
        if name in floors:
            assert math.isclose(d.min(), floors[name], rel_tol=2e-3)   # printed with %10.3E


def test_missing_order_and_bad_lines(tmp_path):
    p = tmp_path / "x.model"
    p.write_text("NAME X 1.0\nNODE 0 0 0 -\nNODE 2 0 1 -\n")
    with pytest.raises(ModelFormatError):
        parse_model(str(p))
    p.write_text("NAME X 1.0\nNODE 0 0 0 -\nNODE 1 1 0 -\nPROB 0 * A:1.0\nPROB 1 * A:1.0\n")
    with pytest.raises(ModelFormatError):
        parse_model(str(p))
