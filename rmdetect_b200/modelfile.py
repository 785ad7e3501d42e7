"""`.model` text format reader and the derived search structures.

The on-disk format stays exactly the reference's (reference
lib/model_parser.py:49-70): ``NAME/NODE/PROB/PAIRS/NO_PAIRING/ORDER/SEPARATION/
SYMMETRIC`` records, ``#`` comments, probabilities stored as natural logs.  The
objects built here expose the attribute names the reference's callers rely on
(``nodes, order, pairing, chains_length, chains_count, OFFSETS, pairs_list,
sep_min, symmetric, full_name()``; reference lib/model.py:13-53) so that either
this loader or the reference's own ``ModelParser`` can feed the flattener in
``modelflat.py``.

Nothing here runs per placement: it is load-time host logic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

PTYPE_MUST, PTYPE_NO, PTYPE_CAN = 1, 2, 3        # reference lib/pair.py:2-4
_PTYPES = {"MUST": PTYPE_MUST, "NO": PTYPE_NO, "CAN": PTYPE_CAN}
GC_ROW = "GC"                                    # reference lib/model_parser.py:146-148


class ModelFormatError(ValueError):
    """Raised where the reference prints a message and calls quit()."""


@dataclass
class NodeDef:
    """One model position (reference lib/node.py:3-13)."""
    id: int
    chain: int
    ndx: int
    conds: list | None
    probs: dict = field(default_factory=dict)   # key -> {letter: log p} | "GC"

    def has_gap(self) -> bool:
        # reference lib/node.py:27-35: any non-GC row that lists "."
        return any(row != GC_ROW and "." in row for row in self.probs.values())


@dataclass
class PairDef:
    id1: int
    id2: int
    ptype: int


def gap_offsets(nodes, chains_length):
    """All 2^g gap configurations (reference lib/model.py:156-181).

    Returns a list of per-chain offset lists; a gapped position holds None and
    every later position of that chain moves one letter to the left.  The
    doubling order (existing configurations first, then their gapped copies,
    nodes visited in id order) fixes the leaf order of the search tree.
    """
    configs = [[list(range(chains_length[c])) for c in range(len(chains_length))]]
    for node in nodes:
        if not node.has_gap():
            continue
        gapped = []
        for cfg in configs:
            new = [list(ch) for ch in cfg]
            row = new[node.chain]
            for i in range(len(row) - 1, node.ndx, -1):
                row[i] = row[i - 1]
            row[node.ndx] = None
            gapped.append(new)
        configs.extend(gapped)
    return configs


def pairing_configs(nodes, pairing, offsets):
    """Distinct per-configuration pair tuples (reference lib/model.py:184-209).

    The reference collects the tuples in a Python ``set`` and iterates it, so
    the order in which gate configurations are summed is the set's iteration
    order for these int tuples under the running interpreter.  Building the
    same set from the same tuples reproduces that order here (int/tuple hashes
    are not salted).  A paired position that may be a gap is a format error
    (reference lib/model.py:198-202).
    """
    seen = set()
    for cfg in offsets:
        row = []
        for pr in pairing:
            n1, n2 = nodes[pr.id1], nodes[pr.id2]
            tup = (n1.chain, cfg[n1.chain][n1.ndx], n2.chain, cfg[n2.chain][n2.ndx], pr.ptype)
            if tup[1] is None or tup[3] is None:
                raise ModelFormatError(
                    "canonical pairs should not allow gaps: check pairing (%d, %d)" % (pr.id1, pr.id2))
            row.append(tup)
        seen.add(tuple(row))
    return [list(x) for x in list(seen)]


class ModuleModel:
    """A parsed module model plus its derived enumeration data."""

    def __init__(self, name, version, nodes, order, pairing, unpaired,
                 sep_min=None, sep_max=None, symmetric=False):
        self.name = name
        self.version = version
        self.nodes = nodes
        self.pairing = pairing
        self.unpaired = unpaired
        self.sep_min = sep_min
        self.sep_max = sep_max          # parsed, never used (reference lib/model.py:24)
        self.symmetric = symmetric
        if order is None:
            raise ModelFormatError(
                "model %s has no ORDER record; automatic ordering (reference "
                "lib/model.py:212-303) is an rmbuild feature outside the scan path" % name)
        by_id = {n.id: n for n in nodes}
        self.order = [by_id[i] for i in order]

        self.chains_length = {}
        for n in nodes:                                   # reference lib/model.py:148-154
            self.chains_length[n.chain] = self.chains_length.get(n.chain, 0) + 1
        self.chains_count = len(self.chains_length)
        if sorted(self.chains_length) != list(range(self.chains_count)):
            raise ModelFormatError("chain ids must be 0..n-1")

        self.OFFSETS = gap_offsets(nodes, self.chains_length)
        self.count = len(self.OFFSETS)
        self.pairs_list = pairing_configs(nodes, pairing, self.OFFSETS)

    def full_name(self) -> str:
        return ("%s_%s" % (self.name, self.version)).upper()   # reference lib/model.py:58-61


def parse_model(path) -> ModuleModel:
    """Read one `.model` file (reference lib/model_parser.py:23-74)."""
    name = version = ""
    nodes: list[NodeDef] = []
    pairing: list[PairDef] = []
    unpaired: list[int] = []
    order = None
    sep_min = sep_max = None
    symmetric = False

    def fail(what, lineno):
        raise ModelFormatError("%s in line %d of %s" % (what, lineno, path))

    with open(path) as fh:
        for lineno, raw in enumerate(fh, 1):
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            tok = line.split()
            if line.startswith("NAME"):
                if len(tok) != 3:
                    fail("syntax error parsing NAME", lineno)
                name, version = tok[1], tok[2]
            elif line.startswith("NODE"):
                if len(tok) != 5:
                    fail("syntax error parsing NODE", lineno)
                nid, chain, ndx = (int(t) for t in tok[1:4])
                if nid != len(nodes):                       # ids are sequential (ref :124-132)
                    fail("missing node id %d" % len(nodes), lineno)
                conds = None if tok[4] == "-" else [int(t) for t in tok[4].split(",")]
                nodes.append(NodeDef(nid, chain, ndx, conds))
            elif line.startswith("PROB"):
                if len(tok) < 2:
                    fail("syntax error parsing PROB", lineno)
                row = {}
                for item in tok[3:]:
                    if item == GC_ROW:
                        row = GC_ROW
                        break
                    letter, val = item.split(":")
                    if val == "-" or float(val) == 0.0:     # absent letters fall to unknown_prob
                        continue
                    row[letter] = math.log(float(val))      # ref :150-156
                nodes[int(tok[1])].probs[tok[2]] = row
            elif line.startswith("PAIRS"):
                if len(tok) != 4:
                    fail("syntax error parsing PAIRS", lineno)
                pairing.append(PairDef(int(tok[1]), int(tok[2]), _PTYPES[tok[3]]))
            elif line.startswith("NO_PAIRING"):
                unpaired.extend(int(t) for t in tok[1:])
            elif line.startswith("ORDER"):
                if len(tok) != len(nodes) + 1:
                    fail("ORDER declaration has %d nodes, expected %d" % (len(tok) - 1, len(nodes)), lineno)
                order = [int(t) for t in tok[1:]]
            elif line.startswith("SEPARATION"):
                if len(tok) != 3:
                    fail("syntax error parsing SEPARATION", lineno)
                sep_min = None if tok[1] == "-" else int(tok[1])
                sep_max = None if tok[2] == "-" else int(tok[2])
            elif line.startswith("SYMMETRIC"):
                symmetric = True
            elif line.startswith("STRANDS"):
                # the reference's handler names an undefined class
                fail("STRANDS records are not supported", lineno)
            # REF_SEQS / DATA_SOURCE only matter to rmbuild and are ignored here
    return ModuleModel(name, version, nodes, order, pairing, unpaired, sep_min, sep_max, symmetric)
