"""rmbuild's conditional probability tables: the probabilities a `.model` file carries, estimated from alignment rows
(SURVEY.md §8(f) row 4).

`rmbuild.py:300-352` cuts one letter per model node out of every row of the training alignments (`.data` file: key,
weight = 100 / rows of its alignment, N0 … Nn), then fills every node's table: `Model.prob_joint` for a node without
parents, `Model.prob_cond` for the others (`lib/model.py:100-147`), on top of `lib/bayes.py`'s `JointProb` (weighted
counts of letter tuples, rare tuples discounted, `:109-150`) and `CondProb` (`P(a, b) / P(b)`, `:255-290`).

The counting runs on the GPU (`rmd_cpt_counts`, csrc/post_kernels.cuh): integer counts per cell and per run of equally
weighted rows, plus the first row of every cell.  From these the host replays the reference's arithmetic exactly — its
row-by-row float sums (`probs[k] = probs.get(k, 0.0) + weight`), the order in which its dicts meet their keys (which
fixes the order of the discount's subtractions and of the PROB lines), `round(v, 3)` and `prob_normalize`.
`Model.exclude` is never called by rmbuild, so the `exclude` branch of `JointProb.__compute` is not restated.
"""
from __future__ import annotations

import numpy as np

LETTER_ORDER = "ACGU."


def load_data(path):
    """Rows of a `.data` file (rmbuild.py:42-64): list of dicts with key, weight and one letter per N<i> column."""
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as fh:
        head = fh.readline().rstrip("\n").split("\t")
        rows = []
        for line in fh:
            t = line.rstrip("\n").split("\t")
            if len(t) != len(head):
                continue
            row = dict(zip(head, t))
            row["weight"] = float(row["weight"])
            rows.append(row)
    return rows


def _repeat_add(acc, w, c):
    """acc after c sequential additions of w (the float sum the reference builds row by row)."""
    if c == 0:
        return acc
    return float(np.cumsum(np.concatenate([[acc], np.full(int(c), w)]))[-1])


class Counts:
    """Weighted letter-tuple counts of alignment rows, per table (node, parents)."""

    def __init__(self, rows, n_nodes, ctx):
        self.n_rows = len(rows)
        letters = sorted({row["N%d" % i] for row in rows for i in range(n_nodes)})
        if len(letters) > 16:
            raise ValueError("more than 16 distinct symbols in the training rows")
        self.letters = letters
        self.K = max(2, len(letters))
        code = {ch: k for k, ch in enumerate(letters)}
        self.codes = np.array([[code[row["N%d" % i]] for i in range(n_nodes)] for row in rows], dtype=np.uint8)
        w = [row["weight"] for row in rows]
        self.run = np.zeros(len(rows), dtype=np.int32)           # stretches of rows with one weight
        self.run_weight = [w[0]]
        for r in range(1, len(rows)):
            if w[r] != w[r - 1]:
                self.run_weight.append(w[r])
            self.run[r] = len(self.run_weight) - 1
        self.ctx = ctx
        self.total = 0.0
        for k, wk in enumerate(self.run_weight):                 # `total += obs.weight` over all rows (lib/bayes.py:124)
            self.total = _repeat_add(self.total, wk, int((self.run == k).sum()))

    def tables(self, specs):
        """specs: list of (node column, parent columns); returns per spec the reference's `probs` dict BEFORE the discount:
        ordered list of (key tuple, weight sum), keys in the order the rows first show them."""
        tab = np.full((len(specs), 3), -1, dtype=np.int32)
        for t, cols in enumerate(specs):
            tab[t, :len(cols)] = cols
        counts, first = self.ctx.cpt_counts(self.codes, self.K, tab, self.run, len(self.run_weight))
        out = []
        for t, cols in enumerate(specs):
            cells = np.flatnonzero(first[t] != 0x7fffffff)
            cells = cells[np.argsort(first[t][cells], kind="stable")]
            items = []
            for cell in cells.tolist():
                key, rest = [], cell
                for _ in cols:
                    key.append(self.letters[rest % self.K])
                    rest //= self.K
                acc = 0.0
                for k, wk in enumerate(self.run_weight):
                    acc = _repeat_add(acc, wk, int(counts[k, t, cell]))
                items.append((tuple(key), acc))
            out.append(items)
        return out


def joint_prob(items, total):
    """`JointProb.__compute` after the counting (lib/bayes.py:126-150): tuples rarer than 0.001 (0.010 when all gaps) are
    dropped and their weight leaves the total, the rest is divided by it."""
    probs = dict(items)
    new_total = total
    for k, v in items:
        p = v / total
        gap = "".join(k).replace(".", "") == ""
        if (not gap and p < 0.001) or (gap and p < 0.010):
            new_total -= v
            del probs[k]
    return {k: v / new_total for k, v in probs.items()}


def prob_normalize(d):
    """lib/model.py:139-147: the rounding residue is spread evenly."""
    residue = 1.0
    for v in d.values():
        residue -= v
    n = float(len(d))
    for k in d:
        d[k] += residue / n


def estimate_tables(rows, nodes, ctx):
    """Every node's `probs` as rmbuild.py:346-352 sets them.  nodes: list of (id, conds or None); rows: `load_data`.
    Returns {id: {parent key or "*": {letter: probability}}} with the reference's key order."""
    n_nodes = len(nodes)
    cnt = Counts(rows, n_nodes, ctx)
    specs, owner = [], []
    for nid, conds in nodes:
        if conds:
            specs.append((nid,) + tuple(conds)); owner.append((nid, "ab"))
            specs.append(tuple(conds)); owner.append((nid, "b"))
        else:
            specs.append((nid,)); owner.append((nid, "a"))
    if any(len(sp) > 3 for sp in specs):
        raise ValueError("a node with more than two parents")
    tabs = dict(zip(owner, cnt.tables(specs)))
    out = {}
    for nid, conds in nodes:
        if not conds:                                            # Model.prob_joint, lib/model.py:100-113
            d = {k[0]: round(v, 3) for k, v in joint_prob(tabs[(nid, "a")], cnt.total).items()}
            prob_normalize(d)
            out[nid] = {"*": d}
            continue
        jp_ab = joint_prob(tabs[(nid, "ab")], cnt.total)         # CondProb, lib/bayes.py:272-288
        jp_b = joint_prob(tabs[(nid, "b")], cnt.total)
        result = {}
        for k, v in jp_ab.items():                               # Model.prob_cond, lib/model.py:115-137
            k0, k2 = k[0], k[1:]
            if k2 not in jp_b:
                continue
            v = round(v / jp_b[k2], 3)
            if (k0 == "." and v >= 0.01) or v >= 0.005:
                result.setdefault("".join(k2), {})[k0] = v
        for d in result.values():
            prob_normalize(d)
        out[nid] = result
    return out


def prob_lines(nodes, probs):
    """The PROB block of a `.model` file (lib/model_parser.py:269-291)."""
    text = []
    for nid, _ in nodes:
        for key, d in probs[nid].items():
            text.append("PROB\t%2d\t%s" % (nid, key) + "".join("\t%s:%.3f" % (ch, d[ch]) if ch in d else "\t%s:-   " % ch
                                                                for ch in LETTER_ORDER) + "\n")
        text.append("\n")
    return "".join(text)
