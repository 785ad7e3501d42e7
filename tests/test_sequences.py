"""rmdetect_b200/sequences.py against the reference's own `Sequences` (goldens written by tests/golden/make_sequences.py
from lib/sequences.py + lib/file_io.py): the example inputs of BASELINE configs[0] and [1] and files that exercise the
readers' rules, with and without both strands."""
import gzip
import io
import json
import os

import pytest

from rmdetect_b200.sequences import Sequence, Sequences, read_clustal, read_fasta, read_stockholm

HERE = os.path.dirname(os.path.abspath(__file__))
EX = os.path.join(HERE, "golden", "ref_data", "examples")
GOLD = json.load(gzip.open(os.path.join(HERE, "golden", "sequences.json.gz")))


def view(s):
    return dict(rows=[[q.key, q.seq, q.seq_gapped, bool(q.rev), q.col_index, q.seq_len] for q in s.sequence_list],
                alignment=bool(s.alignment), from_stdin=bool(s.from_stdin), max_pos=s.sequence_max_pos, max_col=s.sequence_max_col,
                index=s.sequence_index, seq_count=s.seq_count())


@pytest.mark.parametrize("case", sorted(k for k in GOLD if not k.startswith("str|")))
def test_files_load_as_the_reference_loads_them(case, capsys):
    name, both = case.split("|")
    assert view(Sequences(fname=os.path.join(EX, name), both_strands=both == "1")) == GOLD[case]


def test_strings_and_stdin():
    assert view(Sequences(seq="acgu\ntt-ga\n")) == GOLD["str|0"]
    assert view(Sequences(seq="acgu\nttnga\n", both_strands=True)) == GOLD["str|1"]
    got = view(Sequences(stdin=io.StringIO("acgu\ntt-ga\n")))
    assert got == GOLD["str|0"]


def test_the_inputs_cover_the_rules():
    rows = {c: GOLD[c]["rows"] for c in GOLD}
    assert len(rows["lys.stk|0"]) == 49 and len(rows["arich_test.stk|0"]) == 100 and len(rows["rnaselci.stk|0"]) == 26
    assert rows["lys.stk|0"][0][4] is not None                     # alignments carry a column index
    assert len(rows["lys.stk|1"]) == 98 and rows["lys.stk|1"][49][0].endswith("--") and rows["lys.stk|1"][49][4] is None
    assert rows["unbalanced.stk|0"] == [] and rows["nohead.stk|0"] == []
    assert [r[0] for r in rows["blocks.stockholm|0"]] == ["seqA", "seqC"]   # seqB holds an N, seqD comes after //
    assert rows["blocks.stockholm|0"][0][2] == "ACGU.....ACGUUUUU"
    assert [r[0] for r in rows["mixed.fasta|0"]] == ["first", "second|x", "first"] and GOLD["mixed.fasta|0"]["index"]["first"] == 2
    assert rows["small.aln|0"][0][2] == "ACGU--ACGUGGCC" and rows["small.aln|0"][1][1] == "ACGUUUACGUGGC"


def test_readers_and_errors(tmp_path):
    assert read_fasta(io.StringIO(">a b\nAC\nGU\n\n>c\nU\n")) == [("a", "ACGU"), ("c", "U")]
    assert read_stockholm(io.StringIO("# STOCKHOLM 1.0\nk  AC-G\nk2 AC G U\n//\n"), warn=None) == []      # spaces in a row
    assert read_stockholm(io.StringIO("# STOCKHOLM 1.0\nk  AC-G\n#=GC SS_cons <..>\n//\n"), warn=None) == [("k", "AC.G")]
    assert read_clustal(io.StringIO("not clustal\nx ACGU\n"), warn=None) == []
    with pytest.raises(SystemExit):
        Sequences(fname=str(tmp_path / "missing.seq"))
    s = Sequence("k", " ac-gt\n")
    assert (s.seq, s.seq_gapped, s.col_index, s.seq_len) == ("ACGU", "AC-GU", [0, 1, 3, 4], 4)
    r = Sequence("k", "AACGUN").reverse()
    assert (r.key, r.seq, r.rev, r.col_index) == ("k--", "NACGUU", True, None)
    p = tmp_path / "x.txt"
    p.write_text("ACGU\n")
    assert Sequences(fname=str(p)).sequence_list == []            # an unknown ending loads nothing, as in the reference


def test_scanner_takes_these_objects():
    """The drop-in Scanner reads key, seq and col_index of the list's items (lib/scanner.py:29-68)."""
    s = Sequences(fname=os.path.join(EX, "rnaselci.stk"))
    q = s.sequence_list[0]
    assert isinstance(q.key, str) and set(q.seq) <= set("ACGU") and len(q.col_index) == len(q.seq)
    assert s.get_sequence(q.key) is q and s.seq_count() == 26
