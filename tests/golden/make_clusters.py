#!/usr/bin/env python
"""Golden vectors for rmcluster's algorithm, from the reference's own `lib/cluster.py`.

Run in the build container (reference mounted at /root/reference):   python tests/golden/make_clusters.py

The 120 hits of the shipped lysine result files are loaded into the reference's `Candidate` objects (as in
make_filters.py) and clustered by `ClusterAlgorithm.go_cluster(dlimit)` (lib/cluster.py:185-240) for several column
limits, in file order and in sorted order.  Shim: `lib.cluster.filter` returns a list (`filter(...)[0]` at :188 and the
`filter` objects of :223,:236 need Python 2's list; SURVEY.md defect D5).  Hits are identified by (file, line).
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
import lib.cluster as ref_cluster                              # noqa: E402
from lib.candidates import Candidate                           # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rmdetect_b200.candidates import decode_chains             # noqa: E402

Candidate.__lt__ = lambda self, other: self.__cmp__(other) < 0    # shim D4
ref_cluster.filter = lambda f, l: list(__builtins__.filter(f, l))  # shim D5

FILTER = "ok = (count >= 3 and occur >= 0.050000 and score >= 0.000000 and bpp >= 0.000000 and mi >= 0.000000 and h >= 0.000000)"


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
    out = dict(input=[c.ident for c in cands], seq_count=49, runs=[])
    for order in ("file", "sorted", "reversed"):
        lst = list(cands) if order == "file" else sorted(cands) if order == "sorted" else list(reversed(cands))
        for dlimit in (0, 2, 5, 50):
            algo = ref_cluster.ClusterAlgorithm(list(lst), 49, None, FILTER)
            with contextlib.redirect_stderr(io.StringIO()):
                algo.go_cluster(dlimit)
            out["runs"].append(dict(order=order, dlimit=dlimit, clusters=[
                dict(model=c.model_name, row=c.row, col=c.col, count=c.count, occur=c.occur, score=c.score, bpp=c.bpp,
                     mi=c.mi, h=c.h, text=str(c), cands=[x.ident for x in c.cands]) for c in algo.clusters]))
    with gzip.open(os.path.join(HERE, "clusters.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print([(r["order"], r["dlimit"], len(r["clusters"])) for r in out["runs"]])


if __name__ == "__main__":
    main()
