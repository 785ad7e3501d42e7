"""Flatten a module model into the dense tables the CUDA path consumes.

Input is any object with the reference ``Model`` attributes (this package's
``ModuleModel`` or the reference's own ``lib.model.Model`` after the D1 shim):
``nodes, order, OFFSETS, pairs_list, chains_length, chains_count, symmetric,
sep_min``.  Output is a ``FlatModel`` of POD numpy arrays:

* gate program  — MUST pairs per pairing configuration, in ``pairs_list`` order
  (reference lib/scanner.py:195-214 iterates exactly this);
* first pairs   — distinct first MUST pair of every configuration, used by the
  kernel to derive candidate placements from the non-zeros of the bpp matrix;
* search tree   — the gap-configuration decision tree of reference
  lib/node.py:38-117 laid out in DFS preorder with skip links, parents resolved
  to (chain, offset) on the path, save/restore slots for the running sums;
* CPTs          — dense ``[6^npar rows][6 letters]`` fp64 log-probabilities
  per model node, ``unknown_prob`` pre-filled (reference lib/node.py:215-220),
  GC-sentinel rows marked with +inf.

Letter codes: A=0 C=1 G=2 U=3 gap=4 other=5.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np

from .modelfile import GC_ROW, PTYPE_MUST

LETTERS = "ACGU."
CODE_GAP = 4
CODE_OTHER = 5
NSYM = 6
GC_MARK = math.inf

F_LEAF, F_CHECK_DIST, F_SAVE, F_RESTORE = 1, 2, 4, 8
MAX_BRANCH_DEPTH = 8
MAX_PARENTS = 2

# mirrors `struct rmd_tree_node` in include/rmdetect_b200.h (16 bytes)
TREE_DTYPE = np.dtype([
    ("chain", "i1"), ("oft", "i1"),
    ("par_chain", "i1", (2,)), ("par_oft", "i1", (2,)),
    ("flags", "u1"), ("bdepth", "u1"),
    ("leaf", "u2"), ("skip", "u2"), ("cpt_row", "u2"),
    ("npar", "u1"), ("_pad", "u1"),
])
assert TREE_DTYPE.itemsize == 16

# mirrors `struct rmd_gate_pair` (4 bytes)
PAIR_DTYPE = np.dtype([("c1", "i1"), ("o1", "i1"), ("c2", "i1"), ("o2", "i1")])


class UnsupportedModel(NotImplementedError):
    """The model is valid for the reference but outside what the kernels cover."""


@dataclass
class FlatModel:
    name: str
    version: str
    chain_len: tuple
    symmetric: bool
    sep_min: int
    n_nodes: int
    tree: np.ndarray          # TREE_DTYPE[n_tree]
    leaf_dist: np.ndarray     # int32[n_leaves][2]: (dist for p1-p0, dist for p0-p1)
    cpt: np.ndarray           # float64[n_rows][6]
    gate_pairs: np.ndarray    # PAIR_DTYPE[n_pairs]
    gate_cfg: np.ndarray      # int32[n_cfg+1] prefix into gate_pairs
    first_pairs: np.ndarray   # int8[n_first][2]: (offset on chain 0 row, offset on chain 1 col)
    brute_candidates: bool    # some configuration opens with a same-chain pair
    root_branch: bool         # the implicit root has more than one successor
    leaf_offsets: list        # OFFSETS entry per leaf id (host side: solution reconstruction)
    level_node: list          # model node id per ORDER level

    @property
    def n_leaves(self):
        return len(self.leaf_offsets)


def _node_cpt(node, unknown_prob):
    """Dense table for one model node: rows indexed by parent letter codes."""
    npar = 0 if node.conds is None else len(node.conds)
    rows = NSYM ** npar
    tab = np.full((rows, NSYM), unknown_prob, dtype=np.float64)
    for key, row in node.probs.items():
        if npar == 0:
            if key != "*":
                continue                       # unreachable row: the lookup key is always "*"
            ridx = 0
        else:
            if len(key) != npar or any(ch not in LETTERS for ch in key):
                continue                       # a key the scorer can never form from ACGU./gap
            ridx = 0
            for ch in key:
                ridx = ridx * NSYM + LETTERS.index(ch)
        if row == GC_ROW:
            tab[ridx, :] = GC_MARK
        else:
            for letter, logp in row.items():
                if letter in LETTERS:
                    tab[ridx, LETTERS.index(letter)] = logp
    return tab


def flatten_model(model, unknown_prob) -> FlatModel:
    nchains = model.chains_count
    if nchains != 2:
        raise UnsupportedModel("the CUDA path covers two-chain models; %s has %d chains"
                               % (model.name, nchains))
    if model.sep_min is None:
        # the reference fails with a TypeError while building min_dists (lib/node.py:79)
        raise UnsupportedModel("model %s has no SEPARATION minimum" % model.name)
    chain_len = tuple(int(model.chains_length[c]) for c in range(nchains))
    if max(chain_len) > 100:
        raise UnsupportedModel("chain longer than 100 positions")
    offsets = model.OFFSETS
    order = list(model.order)
    n_levels = len(order)
    for cfg in offsets:
        for ch in cfg:
            if ch and ch[0] is None:
                raise UnsupportedModel("a gap at the first position of a chain changes the "
                                       "duplicate rule (reference lib/scanner.py:230-233)")

    # ---- CPTs, one block per model node id --------------------------------
    cpt_blocks, cpt_row0 = [], {}
    nrows = 0
    for node in model.nodes:
        npar = 0 if node.conds is None else len(node.conds)
        if npar > MAX_PARENTS:
            raise UnsupportedModel("node %d has %d parents (max %d)" % (node.id, npar, MAX_PARENTS))
        tab = _node_cpt(node, unknown_prob)
        cpt_row0[node.id] = nrows
        nrows += tab.shape[0]
        cpt_blocks.append(tab)
    if nrows >= 65536:
        raise UnsupportedModel("CPT too large")
    cpt = np.ascontiguousarray(np.concatenate(cpt_blocks, axis=0))

    # ---- search tree in DFS preorder (reference lib/node.py:97-117) --------
    recs = []          # dicts, preorder

    def children_groups(cfg_ids, level):
        node = order[level]
        groups = {}
        for ci in cfg_ids:
            groups.setdefault(offsets[ci][node.chain][node.ndx], []).append(ci)
        return list(groups.items())     # dict keeps first-occurrence order, like the reference

    def build(cfg_ids, level, path, bdepth, parent_is_branch, dist_done):
        """Emit the tree node for ORDER level `level` restricted to `cfg_ids`."""
        node = order[level]
        oft = offsets[cfg_ids[0]][node.chain][node.ndx]
        rec = dict(chain=node.chain, oft=-1 if oft is None else int(oft), flags=0,
                   bdepth=bdepth, leaf=0, skip=0, cpt_row=cpt_row0[node.id],
                   par_chain=[-1, -1], par_oft=[-1, -1], npar=0)
        me = len(recs)
        recs.append(rec)
        if parent_is_branch:
            rec["flags"] |= F_RESTORE
        if len(cfg_ids) == 1:
            rec["leaf"] = cfg_ids[0]
            if not dist_done:                      # first singleton node on this path (ref :70-79)
                rec["flags"] |= F_CHECK_DIST
                dist_done = True
        here = path + [(node.id, node.chain, oft)]
        if node.conds is not None:                 # parents = ancestors on this path (ref :83-91,143-151)
            rec["npar"] = len(node.conds)
            for k, cid in enumerate(node.conds):
                hit = [p for p in here if p[0] == cid]
                if not hit:
                    raise UnsupportedModel("unresolved parent %d of node %d: ORDER must list parents first"
                                           % (cid, node.id))
                _, pch, poft = hit[-1]
                rec["par_chain"][k] = pch
                rec["par_oft"][k] = -1 if poft is None else int(poft)
        if level + 1 == n_levels:
            rec["flags"] |= F_LEAF
            rec["leaf"] = cfg_ids[0]
        else:
            groups = children_groups(cfg_ids, level + 1)
            branch = len(groups) > 1
            if branch:
                rec["flags"] |= F_SAVE
                if bdepth >= MAX_BRANCH_DEPTH:
                    raise UnsupportedModel("gap tree branches deeper than %d" % MAX_BRANCH_DEPTH)
            for _, ids in groups:
                build(ids, level + 1, here, bdepth + 1 if branch else bdepth, branch, dist_done)
        rec["skip"] = len(recs)
        return me

    top = children_groups(list(range(len(offsets))), 0)
    root_branch = len(top) > 1
    for _, ids in top:
        build(ids, 0, [], 1 if root_branch else 0, root_branch, False)
    if len(recs) >= 65535:
        raise UnsupportedModel("search tree too large")
    tree = np.zeros(len(recs), dtype=TREE_DTYPE)
    for i, r in enumerate(recs):
        for k in ("chain", "oft", "flags", "bdepth", "leaf", "skip", "cpt_row", "npar"):
            tree[i][k] = r[k]
        tree[i]["par_chain"] = r["par_chain"]
        tree[i]["par_oft"] = r["par_oft"]

    # ---- per-leaf separation distances (reference lib/node.py:73-79) -------
    leaf_dist = np.zeros((len(offsets), 2), dtype=np.int32)
    for li, cfg in enumerate(offsets):
        for c in range(2):
            leaf_dist[li, c] = max(o for o in cfg[c] if o is not None) + model.sep_min

    # ---- gate program (reference lib/scanner.py:199-209) -------------------
    pairs, cfg_ptr, firsts, brute = [], [0], [], False
    for cfg in model.pairs_list:
        must = [(c1, o1, c2, o2) for (c1, o1, c2, o2, pt) in cfg if pt == PTYPE_MUST]
        if not must:
            continue                      # contributes nothing to sum or count
        pairs.extend(must)
        cfg_ptr.append(len(pairs))
        c1, o1, c2, o2 = must[0]
        if c1 == c2:
            brute = True
        else:
            fp = (o1, o2) if c1 == 0 else (o2, o1)
            if fp not in firsts:
                firsts.append(fp)
    gate_pairs = np.zeros(len(pairs), dtype=PAIR_DTYPE)
    for i, p in enumerate(pairs):
        gate_pairs[i] = p
    first_pairs = np.array(firsts, dtype=np.int8).reshape(-1, 2)

    return FlatModel(
        name=model.name, version=model.version, chain_len=chain_len,
        symmetric=bool(model.symmetric), sep_min=int(model.sep_min), n_nodes=len(model.nodes),
        tree=tree, leaf_dist=leaf_dist, cpt=cpt, gate_pairs=gate_pairs,
        gate_cfg=np.array(cfg_ptr, dtype=np.int32), first_pairs=first_pairs,
        brute_candidates=brute, root_branch=root_branch,
        leaf_offsets=[[list(ch) for ch in cfg] for cfg in offsets],
        level_node=[n.id for n in order])


# ---------------------------------------------------------------------------
# A CPU walk of the flattened tables.  It exists to test the flattener against
# the oracle without a GPU (tests/test_modelflat.py); the product never calls it.
# ---------------------------------------------------------------------------
def walk_flat(fm: FlatModel, codes, p0, p1, gc_log, cutoff):
    """Evaluate placement (p0, p1) on letter codes; returns (ok, leaf, pm, pr).

    Same decisions, same fp64 additions in the same order as the kernel's
    `eval_placement` (rmdetect_b200/csrc/scan_kernels.cuh).
    """
    ptr = (p0, p1)
    log_quarter = math.log(0.25)
    best_p, best_r, best_leaf = -500.0, -500.0, -1
    if not (0.0 >= 0.0 + cutoff):
        return False, -1, best_p, best_r
    sp = [0.0] * (MAX_BRANCH_DEPTH + 1)
    sr = [0.0] * (MAX_BRANCH_DEPTH + 1)
    pm = pr = 0.0
    t, n = 0, len(fm.tree)
    while t < n:
        nd = fm.tree[t]
        flags = int(nd["flags"])
        if flags & F_RESTORE:
            pm, pr = sp[int(nd["bdepth"]) - 1], sr[int(nd["bdepth"]) - 1]
        if flags & F_CHECK_DIST:
            d01, d10 = (int(v) for v in fm.leaf_dist[int(nd["leaf"])])
            d = p1 - p0
            if (0 <= d < d01) or (0 <= -d < d10):
                t = int(nd["skip"])
                continue
        nt = CODE_GAP if nd["oft"] < 0 else int(codes[ptr[int(nd["chain"])] + int(nd["oft"])])
        row = 0
        for k in range(int(nd["npar"])):
            po = int(nd["par_oft"][k])
            pc = CODE_GAP if po < 0 else int(codes[ptr[int(nd["par_chain"][k])] + po])
            row = row * NSYM + min(pc, CODE_OTHER)
        v = fm.cpt[int(nd["cpt_row"]) + row, min(nt, CODE_OTHER)]
        if v == GC_MARK:
            v = gc_log[nt]
        r = log_quarter if nt == CODE_GAP else gc_log[nt]
        pm = pm + v
        pr = pr + r
        if not (pm >= pr + cutoff):
            t = int(nd["skip"])
            continue
        if flags & F_LEAF:
            if pm > best_p:
                best_p, best_r, best_leaf = pm, pr, int(nd["leaf"])
        elif flags & F_SAVE:
            sp[int(nd["bdepth"])], sr[int(nd["bdepth"])] = pm, pr
        t += 1
    return best_leaf >= 0, best_leaf, best_p, best_r
