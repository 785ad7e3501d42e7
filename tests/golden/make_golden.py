#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory from the reference itself.

Run in the build container, where the reference is mounted read-only at
/root/reference:   python tests/golden/make_golden.py

It imports the reference's own modules (never edits them), applies the shims
SURVEY.md §0.1 lists for a Python-3 run (D1 materialise `pairs_list`; D2/D3
inject a bpp provider instead of forking RNAfold; D4/D5 only touch sorting and
printing, which are done here without the reference's broken helpers), drives
`Scanner.scan`, `NodeSearch.eval/get_solution`, `EValues.select_gc/compute`,
`GC.compute` and the body of `rmcalibrate.py`, and writes what they return.
It also copies the reference's data files the tests need (models, example
alignments as parsed sequences, shipped result files, the two real `dot.ps`
matrices).  No reference source code is copied.

The GPU box has no /root/reference: tests only read the files written here.
"""
import contextlib
import gzip
import io
import json
import math
import os
import random
import shutil
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/lib")

from rmdetect_b200.synth import SynthFold, synth_sequence   # noqa: E402  (the shared bpp stand-in)

import config as ref_config          # noqa: E402
import scanner as ref_scanner        # noqa: E402
import sequences as ref_sequences    # noqa: E402
from lib.bpprobs import BPProbs as RefBPProbs   # noqa: E402
from lib.gcx import GC as RefGC                  # noqa: E402
from lib.candidates import CandidateParser       # noqa: E402

SEED = 20261019
MODEL_ORDER = ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]     # os.listdir order seen in the build container


def load_config(data_path, names=None):
    cfg = ref_config.Config()
    cfg.data_path = data_path
    with contextlib.redirect_stderr(io.StringIO()):
        cfg.configure()
    for m in cfg.models:
        m.pairs_list = list(m.pairs_list)                    # shim D1
    if names is not None:                                    # fixed, listdir-independent order
        by = {m.full_name(): (m, e) for m, e in zip(cfg.models, cfg.evalues)}
        cfg.models = [by[n][0] for n in names]
        cfg.evalues = [by[n][1] for n in names]
    return cfg


def dump(name, obj):
    path = os.path.join(HERE, name)
    raw = json.dumps(obj, separators=(",", ":")).encode()
    if name.endswith(".gz"):
        with gzip.GzipFile(path, "wb", mtime=0) as fh:
            fh.write(raw)
    else:
        with open(path, "wb") as fh:
            fh.write(raw)
    print("wrote", name, len(raw))


# --------------------------------------------------------------------------- data files
def copy_data():
    dst = os.path.join(HERE, "ref_data")
    os.makedirs(dst + "/models", exist_ok=True)
    os.makedirs(dst + "/arich", exist_ok=True)
    for f in sorted(os.listdir(REF + "/models")):
        shutil.copyfile(REF + "/models/" + f, dst + "/models/" + f)
    a = REF + "/examples/arich/single_alignment/"
    shutil.copyfile(a + "arich_1.0.model", dst + "/arich/arich_1.0.model")
    shutil.copyfile(REF + "/examples/arich/calibrate/arich_1.0.evalues", dst + "/arich/arich_1.0.evalues")
    shutil.copyfile(a + "arich.out", dst + "/arich/arich.out")
    shutil.copyfile(a + "dot.ps", dst + "/arich/dot.ps")
    shutil.copyfile(REF + "/dot.ps", dst + "/lys_dot.ps")
    for f in sorted(os.listdir(REF)):
        if f.startswith("cluster_") and f.endswith(".res"):
            shutil.copyfile(REF + "/" + f, dst + "/" + f)
    for f in os.listdir(dst):
        pass
    os.system("chmod -R u+w %s" % dst)


def seqset(path, both=False):
    with contextlib.redirect_stderr(io.StringIO()):
        s = ref_sequences.Sequences(fname=path, both_strands=both)
    return [[q.key, q.seq, q.col_index] for q in s.sequence_list]


def dump_seqsets():
    ex = REF + "/examples/"
    sets = {
        "lys.seq": seqset(ex + "lysine/lys.seq"),
        "lys.stk": seqset(ex + "lysine/lys.stk"),
        "rnaselci.seq": seqset(ex + "rnasel/rnaselci.seq"),
        "rnaselci.stk": seqset(ex + "rnasel/rnaselci.stk"),
        "arich_test.stk": seqset(ex + "arich/single_alignment/arich_test.stk"),
        "lys.seq.both": seqset(ex + "lysine/lys.seq", both=True),
    }
    dump("seqsets.json.gz", sets)
    return sets


# --------------------------------------------------------------------------- model structures
def dump_models(cfgs):
    out = {}
    for cfg in cfgs:
        for m in cfg.models:
            out[m.full_name()] = dict(
                name=m.name, version=m.version,
                chains_length=[m.chains_length[c] for c in range(m.chains_count)],
                symmetric=bool(m.symmetric), sep_min=m.sep_min, sep_max=m.sep_max,
                order=[n.id for n in m.order],
                OFFSETS=m.OFFSETS,
                pairs_list=[[list(p) for p in cfgp] for cfgp in m.pairs_list],
                nodes=[dict(id=n.id, chain=n.chain, ndx=n.ndx, conds=n.conds, probs=n.probs) for n in m.nodes],
            )
    dump("models.json.gz", out)


# --------------------------------------------------------------------------- eval vectors
def sample_motif(model, rng):
    """Ancestral sample of the model's letters along ORDER (gaps included)."""
    letters = {}
    for node in model.order:
        key = "*" if node.conds is None else "".join(letters[c] for c in node.conds)
        row = node.probs.get(key, {})
        if row == "GC" or not row:
            letters[node.id] = rng.choice("ACGU")
        else:
            ks = list(row.keys())
            ws = [math.exp(row[k]) for k in ks]
            letters[node.id] = rng.choices(ks, ws)[0]
    chains = []
    for c in range(model.chains_count):
        ids = sorted((n.ndx, n.id) for n in model.nodes if n.chain == c)
        chains.append("".join(letters[i] for _, i in ids).replace(".", ""))
    return chains


def eval_vectors(cfgs, per_model=400):
    rng = random.Random(SEED)
    out = {}
    for model in [m for cfg in cfgs for m in cfg.models]:
        cases = []
        for k in range(per_model):
            kind = k % 4
            n = rng.randint(40, 150)
            alphabet = "ACGU" if kind != 3 else "ACGUN"
            seq = [rng.choice(alphabet) for _ in range(n)]
            L0, L1 = model.chains_length[0], model.chains_length[1]
            p0 = rng.randint(0, n - L0 - 1)
            p1 = rng.randint(0, n - L1 - 1)
            if kind in (1, 2, 3):                       # plant a sampled motif, sometimes mutated
                c0, c1 = sample_motif(model, rng)
                if rng.random() < 0.5:
                    p0, p1 = sorted((p0, p1))
                    p1 = min(max(p1, p0 + L0 + rng.randint(0, 6)), n - L1 - 1)
                else:
                    p1, p0 = sorted((p0, p1))
                    p0 = min(max(p0, p1 + L1 + rng.randint(0, 6)), n - L0 - 1)
                for t, ch in enumerate(c0):
                    if p0 + t < n:
                        seq[p0 + t] = ch
                for t, ch in enumerate(c1):
                    if p1 + t < n:
                        seq[p1 + t] = ch
                if kind == 2:
                    for _ in range(rng.randint(1, 3)):
                        seq[rng.randrange(n)] = rng.choice("ACGU")
                if kind == 3 and rng.random() < 0.7:
                    seq[rng.choice([p0, p1]) + rng.randint(0, 3)] = "N"
            seq = "".join(seq)
            gc = RefGC.compute(seq) if kind != 0 or k % 8 else RefGC.neutral()
            for ch in set(seq):                         # neutral() lacks N: keep the vector well defined
                gc.setdefault(ch, math.log(0.25))
            cutoff = math.log(0.1) if k % 5 else math.log(1e-3)
            unk = math.log(1e-4)
            ok = model.root.eval(seq, [p0, p1], gc, unk, cutoff, False)
            rec = dict(seq=seq, p=[p0, p1], gc=gc, cutoff=cutoff, unknown_prob=unk, ok=bool(ok))
            if ok:
                seqs, poss, pm, pr = model.root.get_solution()
                rec.update(seqs=["".join(s) for s in seqs], poss=poss, prob_model=pm, prob_rand=pr)
            cases.append(rec)
        out[model.full_name()] = cases
        print(model.full_name(), "eval ok", sum(c["ok"] for c in cases), "of", len(cases))
    dump("eval_vectors.json.gz", out)


# --------------------------------------------------------------------------- scans
def cand_record(c):
    return dict(model=c.model.full_name(), key=c.key, coords=list(c.coords),
                seqs=["".join(s) for s in c.chains_seqs], pos=c.chains_pos, cols=c.chains_cols,
                score=c.score, evalue=c.evalue)


def run_scan(cfg, seqs, win_len, win_step, seed, min_bpp=0.001, cutoff=0.1, unknown=1e-4, trace=False):
    cfg.win_len, cfg.win_step = win_len, win_step
    cfg.min_bpp_mean = min_bpp
    cfg.cutoff = math.log(cutoff)
    cfg.unknown_prob = math.log(unknown)
    sc = ref_scanner.Scanner(cfg)
    sc.rna_tools = SynthFold(seed, factory=RefBPProbs)       # shims D2/D3
    events = []
    if trace:
        sc.cback_before_scan = lambda n: events.append(["before_scan", n])
        sc.cback_before_sequence = lambda n, i, k: events.append(["before_sequence", n, i, k])
        sc.cback_before_window = lambda n, p: events.append(["before_window", n, p])
        sc.cback_after_window = lambda c, t: events.append(["after_window", len(c), list(t)])
        sc.cback_after_sequence = lambda c, t: events.append(["after_sequence", len(c), list(t)])
        sc.cback_after_scan = lambda c, t: events.append(["after_scan", len(c), list(t)])

    class Box:
        pass
    box = Box()
    box.sequence_list = []
    for key, seq, cols in seqs:
        q = Box()
        q.key, q.seq, q.col_index = key, seq, cols
        box.sequence_list.append(q)
    cands, tries = sc.scan(box)
    return dict(tries=list(tries), cands=[cand_record(c) for c in cands], events=events)


def scan_cases(sets):
    cases = []

    def add(name, models, seqs, win_len=150, win_step=75, seed=SEED, data="models", **kw):
        path = REF + "/models" if data == "models" else REF + "/examples/arich/calibrate"
        cfg = load_config(path, models)
        res = run_scan(cfg, seqs, win_len, win_step, seed, **kw)
        cases.append(dict(name=name, models=models, data=data, seqs=seqs, win_len=win_len, win_step=win_step,
                          seed=seed, min_bpp=kw.get("min_bpp", 0.001), cutoff=kw.get("cutoff", 0.1),
                          unknown=kw.get("unknown", 1e-4), **res))
        print(name, res["tries"], len(res["cands"]))

    # BASELINE config 1: lys.seq, KT only
    add("lys_seq_kt", ["KT_1.0"], sets["lys.seq"], trace=True)
    # config 2: all four models, rnasel (triple overlap) and alignments
    add("rnasel_seq_all", MODEL_ORDER, sets["rnaselci.seq"], trace=True)
    add("rnasel_stk_all", MODEL_ORDER, sets["rnaselci.stk"][:8], trace=True)
    add("lys_stk_all_first6", MODEL_ORDER, sets["lys.stk"][:6])
    add("lys_seq_both_all", list(reversed(MODEL_ORDER)), sets["lys.seq.both"])
    add("arich_stk_first12", ["ARICH_1.0"], sets["arich_test.stk"][:12], data="arich")
    # window geometry edge cases on a synthetic sequence
    s600 = synth_sequence(600, SEED + 1)
    add("synth600_step10", MODEL_ORDER, [["s600", s600[:260], None]], win_len=150, win_step=10)
    add("synth600_w64", MODEL_ORDER, [["s600", s600[:300], None]], win_len=64, win_step=17)
    add("synth600_w200", MODEL_ORDER, [["s600", s600, None]], win_len=200, win_step=100)
    add("synth_whole", MODEL_ORDER, [["s600", s600[:320], None]], win_len=0, win_step=0)
    add("synth_exact", MODEL_ORDER, [["a", s600[:150], None], ["b", s600[:151], None], ["c", s600[:225], None],
                                     ["d", s600[:226], None]])
    add("synth_short", MODEL_ORDER, [["e0", "", None], ["e1", "A", None], ["e5", "GGACU", None],
                                     ["e9", s600[:9], None], ["e12", s600[:12], None], ["e21", s600[:21], None],
                                     ["e40", s600[:40], None]])
    # non-ACGU letters and a dense gate (low min_bpp admits nearly every non-zero first pair)
    sN = list(s600[:240])
    for pos in (5, 77, 78, 79, 140, 200, 201):
        sN[pos] = "N"
    sN[33] = "R"
    add("synth_N", MODEL_ORDER, [["sN", "".join(sN), None]])
    add("synth_lowgate", MODEL_ORDER, [["s600", s600[:300], None]], min_bpp=1e-6)
    add("synth_opts", MODEL_ORDER, [["s600", s600[100:400], None]], cutoff=0.01, unknown=1e-3, min_bpp=0.01)
    dump("scan_cases.json.gz", cases)


# --------------------------------------------------------------------------- small functions
def small_vectors(cfg):
    rng = random.Random(SEED + 7)
    gcs = []
    for _ in range(60):
        n = rng.randint(20, 400)
        w = [rng.random() + 0.05 for _ in range(4)]
        alphabet = "ACGU" if rng.random() < 0.8 else "ACGUN"
        if alphabet == "ACGUN":
            w.append(0.1)
        seq = "".join(rng.choices(alphabet, w, k=n))
        if rng.random() < 0.15:
            seq = seq.replace(rng.choice("ACGU"), "")      # a base missing altogether (quirk A.3-7)
        if not seq:
            continue
        gc = RefGC.compute(seq)
        sel = []
        for ev in cfg.evalues:
            ev.select_gc(gc)
            sel.append(ev.selected)
        gcs.append(dict(seq=seq, gc=gc, selected=sel))
    ev_cases = []
    for mi, ev in enumerate(cfg.evalues):
        for _ in range(300):
            ev.selected = rng.randrange(len(ev.gc_classes))
            x = rng.choice([rng.uniform(-4, 30), round(rng.uniform(0, 26), 1) + 0.05, rng.randint(0, 26) + 0.25,
                            rng.randint(0, 26) + 0.75, 0.0, -0.0, -0.04, -0.06])
            ev_cases.append(dict(model=mi, selected=ev.selected, score=x, evalue=ev.compute(x)))
    dump("small_vectors.json.gz", dict(models=[m.full_name() for m in cfg.models], gc=gcs, evalue=ev_cases,
                                       gc_classes=RefGC.get_classes(5, 15, 85)))


# --------------------------------------------------------------------------- calibration
def calibrate_vectors():
    """Run the body of the reference's rmcalibrate.py in-process with a seeded RNG."""
    import types
    stub = types.ModuleType("rpy2")
    stub.robjects = types.ModuleType("rpy2.robjects")
    stub.robjects.r = None
    sys.modules["rpy2"] = stub                                   # shim D8
    sys.modules["rpy2.robjects"] = stub.robjects
    src = open(REF + "/rmcalibrate.py").read()
    out = {}
    for name, samples in (("KT_1.0", 20000), ("GB_1.0", 20000), ("CL_1.0", 20000), ("TGA_1.0", 70000)):
        tmp = "/tmp/golden_%s.evalues" % name
        argv = sys.argv
        sys.argv = ["rmcalibrate.py", "--data-path", REF + "/models", "--samples", str(samples), name, tmp]
        random.seed(SEED)
        g = {"__name__": "__main__"}
        with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
            exec(compile(src, "rmcalibrate.py", "exec"), g)
        sys.argv = argv
        hists = [sorted([k, v] for k, v in h.items()) for h in g["gc_hists"]]
        out[name] = dict(samples=samples, seed=SEED, hists=hists, text=open(tmp).read())
        print("calibrate", name, samples, "rows", len(g["scores"]))
    dump("calibrate.json.gz", out)


# --------------------------------------------------------------------------- whole result files (round 2)
def res_files(sets):
    """BASELINE configs[0] and [1] at full size: the reference's own flow from Scanner.scan to the bytes of the
    `.res` file (rmdetect.py:31-70,237-249 -> lib/results.py:13-40), on the complete example inputs.

    Shims: D1-D3 as everywhere; D4 (`Candidate.__lt__` from its `__cmp__`, needed by `.sort()`); D5 (`filter()[0]` in
    `CandidateParser.code`: the builtin is shadowed inside lib.candidates by a list-returning one); D2/D7 for the joint
    bpp: RNAfold is absent and HEAD's folds ignore the constraint, so every hit gets the 1.0 all shipped result files
    show.  The comment block (time stamps) is left empty."""
    import lib.candidates as ref_cands
    import results as ref_results
    ref_cands.Candidate.__lt__ = lambda a, b: a.__cmp__(b) < 0                         # shim D4
    ref_cands.filter = lambda fn, seq: [x for x in seq if fn(x)]                        # shim D5

    def joint_bpp(cands, seqs, constraint="", cback=None):                             # shims D2 / D7
        for c in cands:
            c.bpp = 1.0
        return cands

    class Seqs:
        from_stdin = False

        def __init__(self, fname, triples, both=False):
            self.fname, self.both_strands = fname, both
            self.sequence_list = []
            for key, seq, cols in triples:
                q = type("S", (), {})()
                q.key, q.seq, q.col_index = key, seq, cols
                self.sequence_list.append(q)

        def seq_count(self):                                                           # lib/sequences.py:65-69
            return len(self.sequence_list) * 2 if self.both_strands else len(self.sequence_list)

    out = []
    plan = [("lys_seq_kt", ["KT_1.0"], "models", "lys.seq", "examples/lysine/lys.seq", False),
            ("lys_seq_both_all", MODEL_ORDER, "models", "lys.seq.both", "examples/lysine/lys.seq", True),
            ("rnasel_seq_all", MODEL_ORDER, "models", "rnaselci.seq", "examples/rnasel/rnaselci.seq", False),
            ("rnasel_stk_all", MODEL_ORDER, "models", "rnaselci.stk", "examples/rnasel/rnaselci.stk", False),
            ("lys_stk_all", MODEL_ORDER, "models", "lys.stk", "examples/lysine/lys.stk", False),
            ("arich_stk", ["ARICH_1.0"], "arich", "arich_test.stk", "examples/arich/single_alignment/arich_test.stk", False)]
    for name, models, data, setname, fname, both in plan:
        path = REF + "/models" if data == "models" else REF + "/examples/arich/calibrate"
        cfg = load_config(path, models)
        cfg.win_len, cfg.win_step = 150, 75
        cfg.min_bpp_mean, cfg.cutoff, cfg.unknown_prob = 0.001, math.log(0.1), math.log(1e-4)
        sc = ref_scanner.Scanner(cfg)
        sc.rna_tools = SynthFold(SEED, factory=RefBPProbs)
        seqs = Seqs(fname, sets[setname], both)
        cands, tries = sc.scan(seqs)
        keys = [q.key for q in seqs.sequence_list]
        names = [m.full_name() for m in cfg.models]
        every = [[names.index(c.model.full_name()), keys.index(c.key), c.coords[0], c.chains_pos[0][0], c.chains_pos[1][0],
                  c.score, c.evalue] for c in cands]
        # rmdetect.py:44-47,70
        kept = ref_cands.Candidates.filter(cands, min_score=5.0)
        kept = joint_bpp(kept, seqs)
        kept = ref_cands.Candidates.filter(kept, min_bpp=0.001)
        kept.sort()
        tmp = "/tmp/golden_%s.res" % name
        ref_results.Results.write(tmp, kept, seqs, "w", tries, "", "")
        # the sequence's both-strand list is already doubled when both_strands is set (lib/sequences.py:65-69 quirk)
        out.append(dict(name=name, models=models, data=data, seqset=setname, seq_file=fname, both_strands=both,
                        seed=SEED, tries=list(tries), n_sequences=seqs.seq_count(), all=every, res=open(tmp).read()))
        print("res", name, tries, len(cands), "->", len(kept), "lines")
    dump("res_files.json.gz", out)


def calibrate_checkpoints():
    """Convergence bookkeeping of rmcalibrate.py (:64-99 table, :134-157 diff): the reference's own functions on its
    own histograms after 10 000 and after all samples of the seeded runs of `calibrate_vectors`."""
    import types
    stub = types.ModuleType("rpy2")
    stub.robjects = types.ModuleType("rpy2.robjects")
    stub.robjects.r = None
    sys.modules["rpy2"] = stub
    sys.modules["rpy2.robjects"] = stub.robjects
    src = open(REF + "/rmcalibrate.py").read()
    out = {}
    for name, samples in (("KT_1.0", 20000), ("GB_1.0", 20000), ("CL_1.0", 20000), ("TGA_1.0", 70000)):
        runs = []
        for n in (10000, samples):
            tmp = "/tmp/golden_ckpt_%s.evalues" % name
            argv = sys.argv
            sys.argv = ["rmcalibrate.py", "--data-path", REF + "/models", "--samples", str(n), name, tmp]
            random.seed(SEED)
            g = {"__name__": "__main__"}
            with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
                exec(compile(src, "rmcalibrate.py", "exec"), g)
            sys.argv = argv
            runs.append(g)
        a, b = runs
        L = len(next(m for m in b["config"].models if m.full_name().upper() == name).nodes)
        gce_a, _ = b["get_gc_evalues"](a["gc_hists"], 1.0 / float(4 ** L))
        gce_b, scores_b = b["get_gc_evalues"](b["gc_hists"], 1.0 / float(4 ** L))
        diff = b["diff_gc_evalues"](gce_a, gce_b, scores_b)
        mid = len(gce_b) // 2
        out[name] = dict(samples=[10000, samples], seed=SEED, scores=scores_b, diff=diff,
                         hists_a=[sorted([k, v] for k, v in h.items()) for h in a["gc_hists"]],
                         middle_a=sorted([k, v] for k, v in gce_a[mid].items()),
                         middle_b=sorted([k, v] for k, v in gce_b[mid].items()))
        print("checkpoint", name, "diff", diff, "rows", len(scores_b))
    dump("calibrate_ckpt.json.gz", out)


# --------------------------------------------------------------------------- rmbuild's probability tables (round 2)
def bayes_tables():
    """lib/bayes.py + lib/model.py:100-147 as rmbuild.py:343-352 drives them, on the reference's own training rows
    (examples/arich/*/ARICH_1.0.data with the node dependencies of the .def file next to it).  The rows are copied as
    fixtures (gzip); recorded: every node's `probs` and the PROB block `ModelParser.write` prints for them."""
    import lib.bayes as ref_bayes
    from lib.model_parser import ModelParser
    dst = os.path.join(HERE, "ref_data", "arich_build")
    os.makedirs(dst, exist_ok=True)
    out = {}
    for name in ("two_alignments", "edit_network"):
        src = REF + "/examples/arich/" + name
        raw = open(src + "/ARICH_1.0.data", "rb").read()
        with gzip.GzipFile(os.path.join(dst, name + ".data.gz"), "wb", mtime=0) as fh:
            fh.write(raw)
        lines = raw.decode().split("\n")
        head = lines[0].split("\t")
        data = []
        for line in lines[1:]:
            t = line.split("\t")
            if len(t) == len(head):
                row = dict(zip(head, t))
                row["weight"] = float(row["weight"])
                data.append(row)
        os.chdir(src)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            def_model = ModelParser().parse(src + "/ARICH_1.0.def")
            pmodel = ref_bayes.Model(data)
            for i, node in enumerate(def_model.nodes):
                if node.conds is None:
                    node.probs = def_model.prob_joint(pmodel, "N%d" % i)
                else:
                    node.probs = def_model.prob_cond(pmodel, "N%d" % i, ["N%d" % c for c in node.conds])
            def_model.data_sources = None
            def_model.ref_seqs = None
            tmp = "/tmp/golden_build_%s.model" % name
            ModelParser().write(tmp, def_model)
        text = open(tmp).read()
        block = "".join(l + "\n" for l in text.split("\n") if l.startswith("PROB") or l == "")
        out[name] = dict(rows=len(data), nodes=[[n.id, n.conds] for n in def_model.nodes],
                         probs={str(n.id): n.probs for n in def_model.nodes},
                         prob_lines=[l for l in text.split("\n") if l.startswith("PROB")],
                         shipped_model=open(src + "/arich_1.0.model").read())
        print("bayes", name, len(data), "rows", len(out[name]["prob_lines"]), "PROB lines")
    os.system("chmod -R u+w %s" % dst)
    dump("bayes_tables.json.gz", out)


def main():
    only = set(sys.argv[1:])                      # e.g. `make_golden.py res ckpt` regenerates just those parts
    want = lambda part: not only or part in only
    copy_data()
    sets = dump_seqsets()
    cfg = load_config(REF + "/models", MODEL_ORDER)
    cfg_a = load_config(REF + "/examples/arich/calibrate", ["ARICH_1.0"])
    if want("models"):
        dump_models([cfg, cfg_a])
    if want("eval"):
        eval_vectors([cfg, cfg_a])
    if want("small"):
        small_vectors(cfg)
    if want("scan"):
        scan_cases(sets)
    if want("calibrate"):
        calibrate_vectors()
    if want("res"):
        res_files(sets)
    if want("ckpt"):
        calibrate_checkpoints()
    if want("bayes"):
        bayes_tables()


if __name__ == "__main__":
    main()
