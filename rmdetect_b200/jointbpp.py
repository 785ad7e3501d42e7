"""Post-scan joint base-pair probability of a hit.

After the scan `rmdetect.py` keeps hits with `score > --min-score`, computes for each of them the probability
that ALL of the module's pairs form together, drops hits at or below `--min-bpp`, sorts and writes
(`rmdetect.py:44-47`).  The probability comes from two ensemble free energies of the hit's window
(`lib/candidates.py:396-419`, "thanks to Rolf Backofen"): folded freely and folded with the module's pairs
imposed, `bpp = e ** ((F_free - F_constrained) / kT)`, kT = 0.61702 kcal/mol (37 °C); a value above 1 is
taken as 0.  The reference as shipped never passes `-C` to RNAfold (`lib/rnatools.py:134`, SURVEY.md defect
D7), so its constraint is inert and every hit gets 1.0 (all `cluster_*.res` files show `1.00000E+00`); here the
constraint is applied unless the fold driver was built with `use_constraint=False`.

Restated from the reference: `set_pairing` (`lib/candidates.py:550-587`), `get_pattern_2d` (`:420-445`) with the
user-constraint merge (`merge_pairs`, `pattern_2_pairs`, `:9-67`), `Candidates.compute_bpp` (`:209-224`).  Folding itself is ViennaRNA's
(`fold.RNAfold.fold`); the two folds per hit of a list run concurrently.
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor

from .modelfile import PTYPE_CAN, PTYPE_MUST

KT = 0.61702
CANONICAL = ("AU", "CG", "GC", "GU", "UA", "UG")


def pairing_of(cand, strict=True):
    """(pairs, singles) in window coordinates: MUST pairs, CAN pairs whose letters can pair, and — as positions
    that must stay single — the letters of CAN pairs that cannot, plus the model's unpaired nodes."""
    pairs, singles = [], []
    nodes = cand.model.nodes
    for pr in cand.model.pairing:
        n1, n2 = nodes[pr.id1], nodes[pr.id2]
        p1, p2 = cand.chains_pos[n1.chain][n1.ndx], cand.chains_pos[n2.chain][n2.ndx]
        if p1 is None or p2 is None:
            continue
        p1, p2 = p1 - cand.coords[0], p2 - cand.coords[0]
        pair = (min(p1, p2), max(p1, p2))
        if not strict:
            pairs.append(pair)
        elif pr.ptype == PTYPE_MUST:
            pairs.append(pair)
        elif pr.ptype == PTYPE_CAN:
            if cand.chains_seqs[n1.chain][n1.ndx] + cand.chains_seqs[n2.chain][n2.ndx] in CANONICAL:
                pairs.append(pair)
            else:
                singles.extend(pair)
    for nid in cand.model.unpaired:
        n = nodes[nid]
        p = cand.chains_pos[n.chain][n.ndx]
        if p is not None:
            singles.append(p - cand.coords[0])
    return pairs, singles


def pattern_pairs(pat):
    """Pairs (open, close) of a dot-bracket string in closing order (`pattern_2_pairs`, lib/candidates.py:51-67).
    The reference fails on an unbalanced string (it extends a list with False, or names an undefined `false`); here
    that is a ValueError."""
    pairs, stack = [], []
    for i, c in enumerate(pat):
        if c == "(":
            stack.append(i)
        if c == ")":
            if not stack:
                raise ValueError("constraint closes a pair it never opened at position %d" % i)
            pairs.append((stack.pop(), i))
    if stack:
        raise ValueError("constraint leaves %d pairs open" % len(stack))
    return pairs


def merge_pairs(pairs):
    """`merge_pairs` (lib/candidates.py:9-49): of two pairs that share a base or cross each other, drop the one named
    less often in `pairs` (a tie drops the one that opens later); the survivors in first-seen order."""
    count = {}
    for k in pairs:
        count[k] = count.get(k, 0) + 1
    keys = sorted(count)
    for (o1, c1) in keys:
        v1 = count.get((o1, c1))
        if v1 is None:
            continue
        for (o2, c2) in keys:
            v2 = count.get((o2, c2))
            if v2 is None or (o1 == o2 and c1 == c2):
                continue
            apart = not (o1 == o2 or o1 == c2 or c1 == o2 or c1 == c2)
            nested = c2 < o1 or o2 > c1 or (o2 > o1 and c2 < c1) or (o2 < o1 and c2 > c1)
            if not apart or not nested:
                if v1 < v2:
                    k = (o1, c1)
                elif v1 > v2:
                    k = (o2, c2)
                elif o1 > o2:
                    k = (o1, c1)
                else:
                    k = (o2, c2)
                if count.get(k) is not None:
                    del count[k]
    return list(count)


def pattern_2d(cand, constraint=""):
    """Constraint string of the hit's window for `RNAfold -C` (`Candidate.get_pattern_2d`, lib/candidates.py:420-445):
    brackets on the module's pairs, `x` on its singles.  A user constraint (rmdetect.py --constraint, as long as the
    sequence) is cut to the window; its pairs join the module's through `merge_pairs`, and its other symbols fill the
    positions the module leaves open (:424-428,440-443)."""
    pairs, singles = pairing_of(cand)
    constraint = constraint[cand.coords[0]:cand.coords[1]]
    if constraint != "":
        pairs = merge_pairs(list(pairs) + pattern_pairs(constraint))
    pat = ["."] * (cand.coords[1] - cand.coords[0])
    for a, b in pairs:
        pat[a], pat[b] = "(", ")"
    for p in singles:
        pat[p] = "x"
    if constraint != "":
        for i, c in enumerate(constraint):
            if pat[i] == "." and c not in "()":
                pat[i] = c
    return "".join(pat)


def joint_bpp(cand, seq, fold, constraint=""):
    """`Candidate.compute_bpp` (lib/candidates.py:396-419) with `fold` a fold.RNAfold.  Like the reference, the free
    fold receives the user's constraint un-cut (:409), the constrained one the window's merged pattern."""
    win = seq[cand.coords[0]:cand.coords[1]]
    _, f_con = fold.fold(win, constraint=pattern_2d(cand, constraint), ensemble=True)
    _, f_all = fold.fold(win, constraint=constraint, ensemble=True)
    bpp = math.e ** ((f_all - f_con) / KT)
    return 0.0 if bpp > 1.0 else bpp


def compute_bpp(cands, sequence_of, fold, workers=None, constraint=""):
    """`Candidates.compute_bpp` (lib/candidates.py:209-224): fills `bpp` where it is None, drops hits whose value is
    0.0, keeps the order.  `sequence_of(key)` returns the hit's sequence."""
    todo = [c for c in cands if c.bpp is None]
    n = max(1, min(int(workers or fold.workers), len(todo) or 1))
    if n == 1:
        vals = [joint_bpp(c, sequence_of(c.key), fold, constraint) for c in todo]
    else:
        with ThreadPoolExecutor(max_workers=n) as pool:
            vals = list(pool.map(lambda c: joint_bpp(c, sequence_of(c.key), fold, constraint), todo))
    for c, v in zip(todo, vals):
        c.bpp = v
    return [c for c in cands if c.bpp != 0.0]
