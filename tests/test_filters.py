"""rmfilter's selections (rmdetect_b200/candidates.py: top_model, top_seq, top_seq_model, get_model, nooverlap) against
what the reference's own `Candidates` methods return for the 120 hits of its shipped lysine result files
(tests/golden/filters.json.gz, written by tests/golden/make_filters.py)."""
import os

from helpers import REF_DATA, golden, load_model
from rmdetect_b200 import candidates as C
from rmdetect_b200.results import load_results


def _hits():
    models = [load_model(n) for n in ("KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0")]
    g = golden("filters")
    by_file = {}
    out = []
    for fname, line in g["input"]:
        if fname not in by_file:
            with open(os.path.join(REF_DATA, fname)) as fh:
                by_file[fname] = fh.read().split("\n")
        c = C.decode_candidate(by_file[fname][line], models)
        out.append((c, [fname, line]))
    return g, out


def test_selections_equal_the_reference():
    g, hits = _hits()
    ident = {id(c): i for c, i in hits}
    cands = [c for c, _ in hits]
    ids = lambda lst: [ident[id(c)] for c in lst]
    assert len(cands) == 120
    assert ids(C.top_model(cands, 3)) == g["top_model_3"]
    assert ids(C.top_seq(cands, 1)) == g["top_seq_1"]
    assert ids(C.top_seq_model(cands, 2)) == g["top_seq_model_2"]
    assert ids(C.get_model(cands, "kt_1.0")) == g["get_model_kt"]
    assert ids(C.nooverlap(cands)) == g["nooverlap"]
    assert ids(C.nooverlap(sorted(cands))) == g["nooverlap_sorted"]
    assert len(g["nooverlap"]) == 103 and len(g["top_seq_1"]) == 47 and len(g["top_model_3"]) == 12
