"""`.res` files (rmdetect_b200/results.py): the reference's shipped result files are read and written back byte for
byte — header, hit lines with their run-length position strings, `%.5E` numbers — and the incremental save modes of
the scan (new / append / rewrite) behave as lib/results.py:13-40."""
import glob
import os

from helpers import REF_DATA, load_model
from rmdetect_b200.results import load_results, write_results

MODELS = ["KT_1.0", "GB_1.0", "CL_1.0", "TGA_1.0", "ARICH_1.0"]


def test_shipped_result_files_round_trip(tmp_path):
    models = [load_model(n) for n in MODELS]
    files = sorted(glob.glob(os.path.join(REF_DATA, "cluster_*.res"))) + [os.path.join(REF_DATA, "arich", "arich.out")]
    assert len(files) == 9
    n_hits = 0
    for path in files:
        r = load_results(path, models)
        n_hits += len(r["cands"])
        out = tmp_path / os.path.basename(path)
        write_results(str(out), r["cands"], "n", r["tries"], seq_file=r["seq_file"], both_strands=r["both_strands"],
                      n_sequences=r["n_sequences"], constraint=r["constraint"], comments=r["comments"])
        assert out.read_bytes() == open(path, "rb").read(), path
    assert n_hits == 324
    r = load_results(files[0], models)
    assert (r["seq_file"], r["both_strands"], r["n_sequences"], r["tries"]) == ("examples/lysine/lys.stk", False, 49, [6952022, 294495])
    assert sorted(r["cands"])[0].key <= sorted(r["cands"])[-1].key          # candidates order like the reference's __cmp__


def test_incremental_save_modes(tmp_path):
    models = [load_model(n) for n in MODELS]
    r = load_results(os.path.join(REF_DATA, "cluster_006_KT_1.0.res"), models)
    out = str(tmp_path / "scan.res")
    write_results(out, r["cands"][:2], "n", [10, 1], seq_file="x.seq", n_sequences=1)
    write_results(out, r["cands"][2:5], "a", [0, 0])                          # appended hits carry no header
    got = load_results(out, models)
    assert len(got["cands"]) == 5 and got["tries"] == [10, 1] and open(out).read().count("##tests_all") == 1
    write_results(out, r["cands"], "w", r["tries"], seq="ACGU", both_strands=True, n_sequences=2, constraint="(..)", comments="# done\n")
    got = load_results(out, models)
    assert (got["seq"], got["seq_file"], got["constraint"], got["both_strands"], got["n_sequences"]) == ("ACGU", None, "(..)", True, 2)
    assert len(got["cands"]) == len(r["cands"]) and got["comments"] == "# done\n"
