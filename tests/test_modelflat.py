"""Flattened device tables re-walked on the CPU against the reference's eval vectors."""
import numpy as np
import pytest

from helpers import MODEL_FILES, golden, load_model
from rmdetect_b200.gcx import encode_sequence, gc_table
from rmdetect_b200.modelflat import F_LEAF, flatten_model, walk_flat


@pytest.mark.parametrize("name", sorted(MODEL_FILES))
def test_walk_flat_matches_reference_eval(name):
    m = load_model(name)
    for case in golden("eval_vectors")[name]:
        fm = flatten_model(m, case["unknown_prob"])
        codes, others = encode_sequence(case["seq"])
        ok, leaf, pm, pr = walk_flat(fm, codes, case["p"][0], case["p"][1], gc_table(case["gc"], others),
                                     case["cutoff"])
        assert ok == case["ok"]
        if ok:
            assert pm == case["prob_model"] and pr == case["prob_rand"]
            offs = fm.leaf_offsets[leaf]
            seqs = ["".join("." if o is None else case["seq"][case["p"][c] + o] for o in offs[c]) for c in range(2)]
            poss = [[None if o is None else case["p"][c] + o for o in offs[c]] for c in range(2)]
            assert seqs == case["seqs"] and poss == case["poss"]


def test_tree_shapes():
    want_leaves = {"KT_1.0": 8, "GB_1.0": 2, "CL_1.0": 16, "TGA_1.0": 1, "ARICH_1.0": 2}
    for name, n in want_leaves.items():
        fm = flatten_model(load_model(name), -9.0)
        assert int(((fm.tree["flags"] & F_LEAF) != 0).sum()) == n == fm.n_leaves
        assert (fm.tree["skip"] > np.arange(len(fm.tree))).all()
        assert fm.gate_cfg[0] == 0 and fm.gate_cfg[-1] == len(fm.gate_pairs)
    assert flatten_model(load_model("CL_1.0"), -9.0).first_pairs.tolist() == [[0, 8], [0, 7], [0, 9], [0, 6]]
    assert flatten_model(load_model("TGA_1.0"), -9.0).symmetric
