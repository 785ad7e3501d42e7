#!/usr/bin/env python
"""Golden vectors for rmfilter's selections, from the reference's own `Candidates` static methods.

Run in the build container (reference mounted at /root/reference):   python tests/golden/make_filters.py

The 120 hits of the shipped lysine result files are loaded into the reference's `Candidate` objects (as in
make_patterns.py) and passed through `Candidates.top_model / top_seq / top_seq_model / get_model / nooverlap`
(lib/candidates.py:99-207).  Shims: `Candidate.__lt__` derived from the reference's own `__cmp__` (Python 3 ignores
`__cmp__`, SURVEY.md defect D4) and `list()` around the `filter` object nooverlap returns (D5).  Hits are identified
by (file, line).
"""
import contextlib
import glob
import gzip
import io
import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/lib")

import config as ref_config                                   # noqa: E402
from lib.candidates import Candidate, Candidates               # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rmdetect_b200.candidates import decode_chains             # noqa: E402

Candidate.__lt__ = lambda self, other: self.__cmp__(other) < 0    # shim D4


def main():
    os.chdir(REF)
    cfg = ref_config.Config()
    cfg.data_path = os.path.join(HERE, "ref_data", "models")
    with contextlib.redirect_stderr(io.StringIO()):
        cfg.configure()
    cands = []
    for path in sorted(glob.glob(os.path.join(HERE, "ref_data", "cluster_*.res"))):
        for no, line in enumerate(open(path)):
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            data = line.split()
            model = next(m for m in cfg.models if m.name == data[0] and m.version == data[1])
            cand = Candidate(model, data[2], [int(t) for t in data[3].split("-")])
            cand.set_data_load(float(data[4]), float(data[5]), float(data[6]), decode_chains(data[9]),
                               [list(t) for t in data[8].split(";")], decode_chains(data[10]))
            cand.ident = [os.path.basename(path), no]
            cands.append(cand)
    ids = lambda lst: [c.ident for c in lst]
    out = dict(input=ids(cands),
               top_model_3=ids(Candidates.top_model(list(cands), 3)),
               top_seq_1=ids(Candidates.top_seq(list(cands), 1)),
               top_seq_model_2=ids(Candidates.top_seq_model(list(cands), 2)),
               get_model_kt=ids(Candidates.get_model(list(cands), "kt_1.0")),
               nooverlap=ids(list(Candidates.nooverlap(list(cands)))),
               nooverlap_sorted=ids(list(Candidates.nooverlap(sorted(cands)))))
    with gzip.open(os.path.join(HERE, "filters.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
