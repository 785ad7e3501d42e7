#!/usr/bin/env python
"""Golden vectors for rmdetect_b200/sequences.py from the reference's own `Sequences` (lib/sequences.py, lib/file_io.py).

Run in the build container (reference at /root/reference, read-only):   python tests/golden/make_sequences.py

Copies the reference's small example inputs (data, not code) to tests/golden/ref_data/examples/, writes a few more
inputs there that exercise the readers' rules (FASTA with lower case / T / blank lines, an interleaved Stockholm file with
the non-standard gap characters and a row that has an invalid character, a Stockholm file whose consensus structure is
unbalanced, a Clustal file), loads every one of them with the reference, with and without both strands, and stores what
it built in sequences.json.gz.  One shim: the reference's Clustal.add_seq calls dict.has_key (Python 2 only,
lib/file_io.py:480); it is replaced at run time by the same statement spelt with `in`.
"""
import contextlib
import gzip
import io
import json
import os
import shutil
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/lib")

import file_io as ref_io              # noqa: E402
import sequences as ref_sequences     # noqa: E402


def _clustal_add_seq(self, key, sequence):                 # lib/file_io.py:478-496 with `in` for has_key
    if key not in self.index:
        self.index[key] = len(self.seqs)
        self.seqs.append((key, sequence))
        self.max_key_len = max(len(key), self.max_key_len)
    else:
        i = self.index[key]
        (k, s) = self.seqs[i]
        self.seqs[i] = (key, s + sequence)


ref_io.Clustal.add_seq = _clustal_add_seq

DST = os.path.join(HERE, "ref_data", "examples")
os.makedirs(DST, exist_ok=True)
for src in ("lysine/lys.seq", "lysine/lys.stk", "rnasel/rnaselci.seq", "rnasel/rnaselci.stk", "arich/single_alignment/arich_test.stk"):
    shutil.copyfile(os.path.join(REF, "examples", src), os.path.join(DST, os.path.basename(src)))
os.system("chmod -R u+w %s" % DST)

EXTRA = {
    "mixed.fasta": ">first some description\nacgtnACGU\n\nTTGACA\n>second|x\nGGGaaaccc\n>first\nAC-GU.AC_G\n",
    "blocks.stockholm": "junk before the signature\n# STOCKHOLM 1.0\n#=GF ID test\n\nseqA  ACGU-_:,~acgu\nseqB  ACGUNACGUACGUA\n"
                        "seqC  GGGG....CCCC..\n#=GC SS_cons <<<<....>>>>..\n#=GC RF xxxxxxxxxxxxxx\n#=GR seqA SS ..............\n\n"
                        "seqA  UUUU\nseqC  AAAA\n#=GC SS_cons ....\n//\nseqD ACGU\n",
    "unbalanced.stk": "# STOCKHOLM 1.0\nseqA  ACGUACGU\n#=GC SS_cons <<..>>>.\n//\n",
    "nohead.stk": "seqA  ACGUACGU\n//\n",
    "small.aln": "CLUSTAL W (1.83) multiple sequence alignment\n\nseq1      ACGU--ACGU\nseq2      ACGUTTACGU\n          ****  ****\n\n"
                 "seq1      GGCC\nseq2      GG-C\n          ** *\n",
    "plain.seq": "acgtu\nACG-U\n  \n",
}
for name, text in EXTRA.items():
    with open(os.path.join(DST, name), "w") as fh:
        fh.write(text)


def load(**kw):
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        s = ref_sequences.Sequences(**kw)
    return dict(rows=[[q.key, q.seq, q.seq_gapped, bool(q.rev), q.col_index, q.seq_len] for q in s.sequence_list],
                alignment=bool(s.alignment), from_stdin=bool(s.from_stdin), max_pos=s.sequence_max_pos, max_col=s.sequence_max_col,
                index=s.sequence_index, seq_count=s.seq_count())


out = {}
for name in sorted(os.listdir(DST)):
    for both in (False, True):
        out["%s|%d" % (name, both)] = load(fname=os.path.join(DST, name), both_strands=both)
out["str|0"] = load(seq="acgu\ntt-ga\n", both_strands=False)
out["str|1"] = load(seq="acgu\nttnga\n", both_strands=True)
raw = json.dumps(out, separators=(",", ":"), sort_keys=True).encode()
with gzip.GzipFile(os.path.join(HERE, "sequences.json.gz"), "wb", mtime=0) as fh:
    fh.write(raw)
print("wrote sequences.json.gz", len(raw), "bytes,", len(out), "cases")
