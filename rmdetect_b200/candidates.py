"""Hit records and the `.res` line codec.

Field names and the text layout are the reference's (lib/candidates.py:339-366,
489-514, 653-698; lib/results.py:13-40) so that rmout/rmfilter/rmcluster keep
reading the files.  Ordering follows the reference's `__cmp__` (key ascending,
E-value ascending, score descending), provided here as rich comparisons
because Python 3 ignores `__cmp__` (SURVEY.md defect D4).
"""
from __future__ import annotations

import math

LN2 = math.log(2.0)


class Candidate:
    __slots__ = ("model", "key", "coords", "chains_seqs", "chains_pos", "chains_cols",
                 "score", "evalue", "bpp", "pairs", "pairs_align", "singles", "_keep")

    def __init__(self, model, key, coords):
        self.model = model
        self.key = key
        self.coords = coords
        self.chains_seqs = []
        self.chains_pos = []
        self.chains_cols = []
        self.score = None
        self.evalue = None
        self.bpp = None
        self.pairs = None
        self.pairs_align = None
        self.singles = None
        self._keep = True

    def set_data_build(self, prob_model, prob_rand, chains_seqs, chains_pos, evalues):
        # reference lib/candidates.py:360-366
        self.chains_seqs = chains_seqs
        self.chains_pos = chains_pos
        self.chains_cols = []
        self.score = (prob_model - prob_rand) / LN2
        self.evalue = evalues.compute(self.score)

    def set_data_load(self, score, evalue, bpp, chains_pos, chains_seqs, chains_cols):
        self.score, self.evalue, self.bpp = score, evalue, bpp
        self.chains_pos, self.chains_seqs, self.chains_cols = chains_pos, chains_seqs, chains_cols

    def update_chains_pos(self, offset):
        # reference lib/candidates.py:508-514
        self.coords = (self.coords[0] + offset, self.coords[1] + offset)
        for chain in self.chains_pos:
            for i, pos in enumerate(chain):
                if pos is not None:
                    chain[i] = pos + offset

    def update_chains_cols(self, cols):
        # reference lib/candidates.py:489-506
        self.chains_cols = [[None if pos is None else (pos if cols is None else cols[pos]) for pos in chain]
                            for chain in self.chains_pos]

    def get_pos(self, c, ndx):
        return self.chains_pos[c][ndx]

    def get_col(self, c, ndx):
        return self.chains_cols[c][ndx]

    # ordering: reference lib/candidates.py:380-394
    def _sort_key(self):
        return (self.key, self.evalue, -self.score)

    def __lt__(self, other):
        return self._sort_key() < other._sort_key()

    def __str__(self):
        return encode_candidate(self)


def encode_chains(chains):
    """Run-length position strings: "" = previous + 1, "-" = gap (reference lib/candidates.py:684-698)."""
    parts = []
    for chain in chains:
        toks = []
        for i, n in enumerate(chain):
            if n is None:
                toks.append("-")
            elif i > 0 and chain[i - 1] is not None and n - 1 == chain[i - 1]:
                toks.append("")
            else:
                toks.append(str(n))
        parts.append(";".join(toks))
    return "/".join(parts)


def decode_chains(text):
    out = []
    for part in text.split("/"):
        chain = []
        for tok in part.split(";"):
            if tok == "-":
                chain.append(None)
            elif tok == "":
                chain.append(chain[-1] + 1)
            else:
                chain.append(int(tok))
        out.append(chain)
    return out


def _first(values):
    return next(v for v in values if v is not None)


def encode_candidate(cand):
    """One `.res` line without the newline (reference lib/candidates.py:653-667)."""
    bpp = cand.bpp if cand.bpp is not None else float("nan")
    heads = ";".join("%d/%d" % (_first(p), _first(c)) for p, c in zip(cand.chains_pos, cand.chains_cols))
    return "%s\t%s\t%s\t%d-%d\t%.5E\t%.5E\t%.5E\t%s\t%s\t%s\t%s\t" % (
        cand.model.name, cand.model.version, cand.key, cand.coords[0], cand.coords[1],
        cand.score, cand.evalue, bpp, heads,
        ";".join("".join(s) for s in cand.chains_seqs),
        encode_chains(cand.chains_pos), encode_chains(cand.chains_cols))


def decode_candidate(line, models):
    """Parse a `.res` hit line (reference lib/candidates.py:624-651)."""
    data = line.split()
    model = next((m for m in models if m.name == data[0] and m.version == data[1]), None)
    if model is None:
        raise KeyError("model '%s' '%s' not found" % (data[0], data[1]))
    s, e = data[3].split("-")
    cand = Candidate(model, data[2], (int(s), int(e)))
    cand.set_data_load(float(data[4]), float(data[5]), float(data[6]),
                       decode_chains(data[9]), [list(t) for t in data[8].split(";")], decode_chains(data[10]))
    return cand


def filter_candidates(cands, min_score=None, min_bpp=None, min_evalue=None):
    """`Candidates.filter` (reference lib/candidates.py:78-97): strict thresholds, then sort."""
    out = []
    for c in cands:
        if min_score is not None and c.score <= min_score:
            continue
        if min_bpp is not None and c.bpp <= min_bpp:
            continue
        if min_evalue is not None and (c.evalue == 0.0 or -math.log(c.evalue) <= min_evalue):
            continue
        out.append(c)
    out.sort()
    return out


# ---- rmfilter's selections (reference lib/candidates.py:99-207, driven by rmfilter.py:62-83) -------------------
def _top(cands, n, key_of):
    out, seen = [], {}
    for c in sorted(cands):
        k = key_of(c)
        seen[k] = seen.get(k, 0) + 1
        if seen[k] <= n:
            out.append(c)
    return out


def top_model(cands, n):
    """The n best hits of every model (`--top-model`, reference lib/candidates.py:99-116)."""
    return _top(cands, n, lambda c: c.model.full_name())


def top_seq(cands, n):
    """The n best hits of every sequence (`--top-seq`, :118-135)."""
    return _top(cands, n, lambda c: c.key)


def top_seq_model(cands, n):
    """The n best hits of every (sequence, model) (`--top-seq-model`, :137-154)."""
    return _top(cands, n, lambda c: c.key + "_" + c.model.full_name())


def get_model(cands, model_name):
    """Hits of one model, sorted (`--model`, :156-169)."""
    return [c for c in sorted(cands) if c.model.full_name().upper() == model_name.upper()]


def nooverlap(cands):
    """Of two hits of one sequence whose windows overlap and whose chains overlap pairwise, keep the earlier one only
    if it scores strictly higher and has a positive bpp, else the later one (`--no-overlap`, :171-207)."""
    cands = list(cands)
    for i in range(len(cands)):
        if cands[i] is None:
            continue
        for j in range(i + 1, len(cands)):
            if cands[j] is None:
                continue
            a, b = cands[i], cands[j]
            dup = a.coords[0] < b.coords[1] and b.coords[0] < a.coords[1] and a.key == b.key
            if dup:
                for pi, pj in zip(a.chains_pos, b.chains_pos):
                    if pi[0] > pj[-1] or pi[-1] < pj[0]:
                        dup = False
                        break
            if dup:
                if a.score > b.score and a.bpp > 0:
                    cands[j] = None
                else:
                    cands[i] = None
                    break
    return [c for c in cands if c is not None]
