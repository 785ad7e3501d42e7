"""The `.res` result file: header + one hit line per candidate.

Layout and semantics are the reference's (`lib/results.py:13-40` write, `:42-83` load; the hit line is
`candidates.encode_candidate`, reference `lib/candidates.py:653-698`), so rmout / rmfilter / rmcluster keep reading
the files:

    <comments>                      free text, every line starting with '#'
    ##seq_file = <path>   |   ##seq = <sequence>          (input from a file / from stdin)
    ##constraint = <dot-bracket>                           only when one was given
    ##both_strands = YES|NO
    ##sequences = <n>
    ##tests_all = <placements tried>
    ##tests_pairs = <placements past the bpp gate>
    <hit lines>

Modes as in the reference: "n" new file, "w" rewrite (both write the header), "a" append hit lines only — the scan
saves incrementally (`rmdetect.py:114-144`).
"""
from __future__ import annotations

import sys

from .candidates import decode_candidate


def write_results(out_name, cands, mode, tries, seq_file=None, seq=None, both_strands=False, n_sequences=0,
                  constraint="", comments=""):
    fo = open(out_name, "a" if mode == "a" else "w") if out_name is not None else sys.stdout
    try:
        if comments:
            fo.write(comments)
        if mode in ("n", "w"):
            if seq is not None:
                fo.write("##seq = %s\n" % seq)
            else:
                fo.write("##seq_file = %s\n" % seq_file)
            if constraint != "":
                fo.write("##constraint = %s\n" % constraint)
            fo.write("##both_strands = %s\n" % ("YES" if both_strands else "NO"))
            fo.write("##sequences = %s\n" % n_sequences)
            fo.write("##tests_all = %d\n" % tries[0])
            fo.write("##tests_pairs = %d\n" % tries[1])
        for cand in cands:
            fo.write("%s\n" % cand)
    finally:
        if out_name is not None:
            fo.close()


def load_results(in_name, models):
    """dict(seq_file, seq, constraint, both_strands, n_sequences, tries, cands, comments) of a `.res` file."""
    out = dict(seq_file=None, seq=None, constraint="", both_strands=False, n_sequences=-1, tries=[0, 0], cands=[],
               comments="")
    fi = open(in_name) if in_name is not None else sys.stdin
    try:
        for raw in fi:
            line = raw.strip()
            if line.startswith("##seq = "):
                out["seq"] = line.split("=")[1].strip()
            elif line.startswith("##seq_file = "):
                out["seq_file"] = line.split("=")[1].strip()
            elif line.startswith("##constraint = "):
                out["constraint"] = line.split("=")[1].strip()
            elif line.startswith("##both_strands = "):
                out["both_strands"] = line.split("=")[1].strip() == "YES"
            elif line.startswith("##sequences = "):
                out["n_sequences"] = int(line.split("=")[1].strip())
            elif line.startswith("##tests_all = "):
                out["tries"][0] = int(line.split("=")[1].strip())
            elif line.startswith("##tests_pairs = "):
                out["tries"][1] = int(line.split("=")[1].strip())
            elif line.startswith("#") or line == "":
                out["comments"] += raw
            else:
                out["cands"].append(decode_candidate(line, models))
    finally:
        if in_name is not None:
            fi.close()
    return out
