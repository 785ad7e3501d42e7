#!/usr/bin/env python
"""Golden vectors for the post-scan joint-bpp stage, from the reference's own code.

Run in the build container (reference mounted at /root/reference):   python tests/golden/make_patterns.py

For every hit line of the reference's shipped result files (tests/golden/ref_data/cluster_*.res, arich/arich.out) it
builds the reference's own `Candidate` (lib/candidates.py:339-366; the line is split here because
`CandidateParser.parse` hands one-shot `map` objects to `coords` and the position lists under Python 3) and records what
`set_pairing()` (:550-587) and `get_pattern_2d("")` (:420-445) return.  Nothing of the reference is copied.
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

import random                                                   # noqa: E402

import config as ref_config                                   # noqa: E402
import lib.candidates as ref_cands                             # noqa: E402
from lib.candidates import Candidate                           # noqa: E402


class _ListKeysDict(dict):
    """Shim D5 for merge_pairs (lib/candidates.py:9-49): `keys = pairs_dict.keys(); keys.sort()` needs the list
    Python 2 returned.  The function itself runs unmodified."""

    def keys(self):
        return list(dict.keys(self))


_orig_pairs_2_dict = ref_cands.pairs_2_dict
ref_cands.pairs_2_dict = lambda pairs: _ListKeysDict(_orig_pairs_2_dict(pairs))


def user_constraint(rng, n):
    """A user constraint over a sequence of n letters: a few nested / side-by-side helices, forced singles, dots."""
    pat = ["."] * n
    for _ in range(rng.randint(1, 4)):
        a = rng.randrange(0, max(1, n - 12))
        b = rng.randrange(min(n - 1, a + 8), n)
        k = rng.randint(1, 4)
        if all(pat[a + t] == "." and pat[b - t] == "." for t in range(k)) and a + k <= b - k:
            for t in range(k):
                pat[a + t], pat[b - t] = "(", ")"
    for _ in range(rng.randint(0, 5)):
        q = rng.randrange(n)
        if pat[q] == ".":
            pat[q] = "x"
    # keep the brackets balanced whatever the loops above left
    depth, out = 0, []
    for c in pat:
        if c == ")" and depth == 0:
            c = "."
        depth += (c == "(") - (c == ")")
        out.append(c)
    for q in range(n - 1, -1, -1):
        if depth and out[q] == "(":
            out[q] = "."
            depth -= 1
    return "".join(out)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rmdetect_b200.candidates import decode_chains             # noqa: E402  (the reference's parse_chains returns py3 map objects)


def models_of(data_path):
    cfg = ref_config.Config()
    cfg.data_path = data_path
    with contextlib.redirect_stderr(io.StringIO()):
        cfg.configure()
    return cfg.models


def main():
    os.chdir(REF)
    sets = [(sorted(glob.glob(os.path.join(HERE, "ref_data", "cluster_*.res"))), models_of(os.path.join(HERE, "ref_data", "models"))),
            ([os.path.join(HERE, "ref_data", "arich", "arich.out")], models_of(os.path.join(HERE, "ref_data", "arich")))]
    out = []
    for files, models in sets:
        for path in files:
            for no, line in enumerate(open(path)):
                line = line.strip()
                if not line or line.startswith("#"):
                    continue
                data = line.split()
                model = next(m for m in models if m.name == data[0] and m.version == data[1])
                cand = Candidate(model, data[2], [int(t) for t in data[3].split("-")])
                cand.set_data_load(float(data[4]), float(data[5]), float(data[6]), decode_chains(data[9]),
                                   [list(t) for t in data[8].split(";")], decode_chains(data[10]))
                pattern = cand.get_pattern_2d("")
                rec = dict(file=os.path.relpath(path, os.path.join(HERE, "ref_data")), line=no, pattern=pattern,
                           pairs=[list(p) for p in cand.pairs], singles=list(cand.singles))
                # the same hit with a user constraint merged in (lib/candidates.py:424-428,440-443); the window is cut
                # out of a constraint as long as the sequence.  Unbalanced slices make the reference raise: skipped.
                rng = random.Random(len(out) * 7919 + no)
                merged = []
                for _ in range(3):
                    con = user_constraint(rng, cand.coords[1] + rng.randint(0, 20))
                    try:
                        merged.append([con, cand.get_pattern_2d(con), [list(p) for p in cand.pairs]])
                    except (TypeError, NameError):
                        merged.append([con, None, None])
                rec["merged"] = merged
                out.append(rec)
    with gzip.open(os.path.join(HERE, "patterns_2d.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print("wrote %d patterns" % len(out))


if __name__ == "__main__":
    main()
