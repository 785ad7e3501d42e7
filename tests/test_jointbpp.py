"""Post-scan joint-bpp stage (rmdetect_b200/jointbpp.py) against the reference's own `set_pairing` /
`get_pattern_2d` outputs for all 324 hits of its shipped result files (tests/golden/patterns_2d.json.gz, written by
tests/golden/make_patterns.py), and the energy formula against a stand-in RNAfold."""
import math
import os
import stat
import sys
import textwrap

import pytest

from helpers import REF_DATA, golden, load_model
from rmdetect_b200.candidates import decode_candidate
from rmdetect_b200.fold import RNAfold
from rmdetect_b200.jointbpp import KT, compute_bpp, joint_bpp, pairing_of, pattern_2d

MODELS = ["KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0", "ARICH_1.0"]


def _hit(rec, models):
    with open(os.path.join(REF_DATA, rec["file"])) as fh:
        line = fh.read().split("\n")[rec["line"]]
    return decode_candidate(line, models)


def test_pairing_and_constraint_pattern_equal_the_reference():
    models = [load_model(n) for n in MODELS]
    recs = golden("patterns_2d")
    assert len(recs) == 324
    for rec in recs:
        cand = _hit(rec, models)
        pairs, singles = pairing_of(cand)
        assert [list(p) for p in pairs] == rec["pairs"] and singles == rec["singles"], rec["file"]
        assert pattern_2d(cand) == rec["pattern"]


def test_user_constraint_merged_into_the_pattern():
    """lib/candidates.py:424-428,440-443: the window's slice of a user constraint joins the module's pairs through
    merge_pairs; 509 of the 972 recorded cases merge, the others cut a helix in two and fail in the reference as here."""
    from rmdetect_b200.jointbpp import merge_pairs, pattern_pairs
    models = [load_model(n) for n in MODELS]
    merged = failed = 0
    for rec in golden("patterns_2d"):
        cand = _hit(rec, models)
        for con, want, pairs in rec["merged"]:
            if want is None:
                with pytest.raises(ValueError):
                    pattern_2d(cand, con)
                failed += 1
            else:
                assert pattern_2d(cand, con) == want
                base, _ = pairing_of(cand)
                assert [list(p) for p in merge_pairs(list(base) + pattern_pairs(con[cand.coords[0]:cand.coords[1]]))] == pairs
                merged += 1
    assert (merged, failed) == (509, 463)


FAKE = textwrap.dedent('''\
    #!%(python)s
    # stand-in for `RNAfold --noPS -p [-C]`: free energy of the ensemble -10.00, raised by 0.25 per imposed pair and
    # 0.10 per position forced single
    import sys
    lines = sys.stdin.read().split("\\n")
    seq, con = lines[0].strip(), (lines[1].strip() if "-C" in sys.argv[1:] and len(lines) > 1 else "")
    assert "--noPS" in sys.argv[1:]
    f = -10.0 + 0.25 * con.count("(") + 0.10 * con.count("x")
    if seq.startswith("UUUU"):
        f = -10.0 - 0.5 * (1 if con else 0)        # constrained ensemble BELOW the free one: ratio > 1
    sys.stdout.write("%%s\\n%%s (%%6.2f)\\n%%s [%%6.2f]\\n frequency of mfe structure in ensemble 0.1; ensemble diversity 3.2\\n"
                     %% (seq, "." * len(seq), -9.5, "." * len(seq), f))
''')


@pytest.fixture(scope="module")
def fake_rnafold(tmp_path_factory):
    path = tmp_path_factory.mktemp("bin") / "RNAfold"
    path.write_text(FAKE % dict(python=sys.executable))
    path.chmod(path.stat().st_mode | stat.S_IXUSR)
    return str(path)


def test_joint_bpp_formula_and_list_stage(fake_rnafold):
    models = [load_model(n) for n in MODELS]
    recs = [r for r in golden("patterns_2d") if r["file"].startswith("cluster_00")][:12]
    cands = [_hit(r, models) for r in recs]
    for c in cands:
        c.bpp = None
    seqs = {c.key: "ACGU" * 200 for c in cands}
    drv = RNAfold(exe=fake_rnafold, workers=4, use_constraint=True)
    want = []
    for c, r in zip(cands, recs):
        f_con = -10.0 + 0.25 * len(r["pairs"]) + 0.10 * len(r["singles"])
        want.append(math.e ** ((-10.0 - round(f_con, 2)) / KT))               # lib/candidates.py:412-414
    assert struct_energy(drv) == (".....", -10.0)
    assert joint_bpp(cands[0], seqs[cands[0].key], drv) == pytest.approx(want[0], rel=1e-12)
    kept = compute_bpp(cands, seqs.__getitem__, drv)
    assert [c.bpp for c in kept] == pytest.approx(want, rel=1e-12) and len(kept) == len(cands)
    assert all(0.0 < b < 1.0 for b in want)
    # the reference as shipped: the constraint is inert and every hit gets e ** 0 = 1.0 (all cluster_*.res show 1.00000E+00)
    for c in cands:
        c.bpp = None
    kept = compute_bpp(cands, seqs.__getitem__, RNAfold(exe=fake_rnafold, workers=2, use_constraint=False))
    assert [c.bpp for c in kept] == [1.0] * len(cands)
    # a ratio above one counts as zero and the hit is dropped (lib/candidates.py:416-417, :217-218)
    for c in cands:
        c.bpp = None
    kept = compute_bpp(cands[:3], lambda key: "UUUU" * 200, drv)
    assert kept == []


def struct_energy(drv):
    return drv.fold("ACGUA", constraint="", ensemble=True)
