"""Genome-scale driver (rmdetect_b200/genome.py): piece geometry and strands on the CPU; on the GPU the pieces of
both strands must add up to the scan of each whole strand, and the reverse strand must equal the oracle's
scan of the reverse-complement string."""
import numpy as np
import pytest

from helpers import LN_CUTOFF, LN_UNKNOWN, MODEL_ORDER, evalues_path, load_evalues, load_model
from rmdetect_b200.dist import plan_chunks
from rmdetect_b200.genome import hit_checksum, rank_pieces, strand_gc, strand_windows, window_start
from rmdetect_b200.scanner import window_starts
from rmdetect_b200.synth import synth_codes, synth_sequence


@pytest.mark.parametrize("n", [1, 149, 150, 151, 225, 226, 1000, 12345])
def test_window_arithmetic_equals_the_window_list(n):
    starts = window_starts(n, 150, 75)
    assert strand_windows(n, 150, 75) == len(starts)
    assert [window_start(k, len(starts), n, 150, 75) for k in range(len(starts))] == starts.tolist()


@pytest.mark.parametrize("n,world,piece", [(12345, 1, 1000), (12345, 3, 20), (5000, 4, 7), (300, 8, 50), (100, 2, 5)])
def test_pieces_partition_the_windows(n, world, piece):
    n_win = strand_windows(n, 150, 75)
    seen = []
    for r in range(world):
        plan, _ = plan_chunks(n, 150, 75, world)
        pieces = rank_pieces(n, 150, 75, r, world, piece)
        if pieces:                                    # the rank's pieces tile the chunk dist.plan_chunks gives it
            assert pieces[0][0] == plan[r][0] and pieces[-1][1] == plan[r][1]
        for a, b, bh, lo, hi in pieces:
            seen.extend(range(a, b))
            assert 0 < b - a <= piece and b <= bh <= min(n_win, b + 2)
            assert lo == window_start(a, n_win, n, 150, 75)
            # scanned as a sequence of its own, the piece has exactly the windows [a, bh) of the parent
            own = window_starts(hi - lo, 150, 75) + lo
            assert own.tolist() == [window_start(k, n_win, n, 150, 75) for k in range(a, bh)]
    assert seen == list(range(n_win))


def test_reverse_strand_twin():
    n, seed = 1000, 7
    fwd = synth_codes(n, seed)
    rev = synth_codes(n, seed, 0, reverse_total=n)
    assert np.array_equal(rev, 3 - fwd[::-1])
    assert np.array_equal(synth_codes(100, seed, 250, reverse_total=n), rev[250:350])
    comp = {"A": "U", "C": "G", "G": "C", "U": "A"}
    assert synth_sequence(n, seed, 0, reverse_total=n) == "".join(comp[c] for c in reversed(synth_sequence(n, seed)))
    g = strand_gc(np.bincount(fwd, minlength=4), True)
    want = np.bincount(rev, minlength=4)
    assert all(abs(np.exp(g[b]) * n - want[k]) < 1e-9 for k, b in enumerate("ACGU"))


def _engine():
    from rmdetect_b200.scanner import ScanEngine
    eng = ScanEngine([load_model(m) for m in MODEL_ORDER], [load_evalues(m) for m in MODEL_ORDER], LN_UNKNOWN)
    eng.ctx.set_screen_policy(0)                      # prefix-screen tables even for these short strands
    return eng


@pytest.mark.gpu
def test_pieces_of_both_strands_add_up_to_the_whole_strands():
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.genome import scan_synthetic_genome
    total, seed = 40_000, 20261019
    eng = _engine()
    want = dict(tests_all=0, tests_pairs=0, hits=0, checksum=0)
    counts = np.bincount(synth_codes(total, seed), minlength=4)
    for strand in (0, 1):                             # each whole strand as one resident job
        gc = strand_gc(counts, strand == 1)
        cls = []
        for ev in eng.evalues:
            ev.select_gc(gc)
            cls.append(ev.selected)
        eng.ctx.synth_strand(total, strand == 1)
        n_win, _, _ = eng.ctx.synth_job(total, 0, seed, seed, 150, 75, 0.001, LN_CUTOFF, gc_table(gc), cls)
        if strand == 1:
            assert np.array_equal(eng.ctx.download_codes(total), synth_codes(total, seed, 0, reverse_total=total))
        r = eng.ctx.scan_resident(0, n_win)
        h = r["hits"]
        want["tests_all"] += r["tries"][0]; want["tests_pairs"] += r["tries"][1]; want["hits"] += len(h)
        want["checksum"] = (want["checksum"] + hit_checksum(strand, h["start"], h["p0"], h["p1"], h["model"], h["leaf"],
                                                            h["score"])) & ((1 << 64) - 1)
    assert want["hits"] > 1000
    for world, piece in ((1, 1000), (1, 37), (3, 50)):
        got = dict(tests_all=0, tests_pairs=0, hits=0, checksum=0)
        for rank in range(world):                     # ranks one after the other on the one GPU: no collective needed
            r = scan_synthetic_genome(eng, total, True, seed, 150, 75, 0.001, LN_CUTOFF, rank=rank, world=world,
                                      piece_windows=piece, forward_counts=None if world == 1 else counts)
            for k in ("tests_all", "tests_pairs", "hits"):
                got[k] += r[k]
            got["checksum"] = (got["checksum"] + r["checksum"]) & ((1 << 64) - 1)
        assert got == want, (world, piece)
    # the command line's score filter in front of the download: same hits as filtering afterwards
    kept = []
    r = scan_synthetic_genome(eng, total, True, seed, 150, 75, 0.001, LN_CUTOFF, piece_windows=90, min_score=5.0,
                              on_hits=lambda strand, lo, h: kept.append(h.copy()))
    kept = np.concatenate(kept)
    every = []
    scan_synthetic_genome(eng, total, True, seed, 150, 75, 0.001, LN_CUTOFF, piece_windows=90,
                          on_hits=lambda strand, lo, h: every.append(h.copy()))
    every = np.concatenate(every)
    assert r["tests_pairs"] == want["tests_pairs"] and r["hits"] == len(kept) == int((every["score"] > 5.0).sum()) > 0
    assert np.array_equal(kept, every[every["score"] > 5.0])
    eng.close()


@pytest.mark.gpu
def test_reverse_strand_scan_matches_oracle():
    from oracle import pyoracle
    from rmdetect_b200.genome import scan_synthetic_genome
    total, seed = 3000, 5
    eng = _engine()
    got = []
    r = scan_synthetic_genome(eng, total, True, seed, 150, 75, 0.001, LN_CUTOFF, piece_windows=11,
                              on_hits=lambda strand, lo, h: got.extend(
                                  (strand, int(x["model"]), int(x["start"]) + lo, int(x["p0"]), int(x["p1"]), int(x["leaf"]),
                                   float(x["score"]), float(x["evalue"])) for x in h))
    eng.close()
    oms = [pyoracle.OracleModel(load_model(n)) for n in MODEL_ORDER]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in MODEL_ORDER]
    want, tests = [], [0, 0]
    for strand in (0, 1):
        seq = synth_sequence(total, seed, 0, reverse_total=total if strand else 0)
        starts = pyoracle.windows(len(seq), 150, 75)
        hits, tries, _ = pyoracle.scan_sequence(oms, oes, 0, seq, 150, 75, [pyoracle.synth_bpp(seq[s:s + 150], seed) for s in starts],
                                                0.001, LN_UNKNOWN, LN_CUTOFF)
        tests[0] += tries[0]; tests[1] += tries[1]
        want.extend((strand, w["model"], w["coords"][0], w["p"][0], w["p"][1], w["cfg"], w["score"], w["evalue"]) for w in hits)
    assert [r["tests_all"], r["tests_pairs"]] == tests
    assert got == want


# ---- world size 2 on CPU (gloo): the driver's two exchanges — letter counts before, counters and checksum after ----
class _OracleCtx:
    """Plays the device context for the genome driver: pieces come from the numpy twin of the synthetic stream and
    are scanned by the CPU oracle (a test of the host-side plumbing; the CUDA path is covered above)."""

    def __init__(self, names):
        from oracle import pyoracle
        self.py = pyoracle
        self.oms = [pyoracle.OracleModel(load_model(n)) for n in names]
        self.oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
        self.total, self.reverse, self.min_score = 0, False, None

    def synth_strand(self, total, reverse):
        self.total, self.reverse = total, reverse

    def synth_counts(self, n, first, seed):
        return np.bincount(synth_codes(n, seed, first, reverse_total=self.total if self.reverse else 0), minlength=4).astype(np.int64)

    def set_min_score(self, v):
        self.min_score = v

    def synth_job(self, n, first, seed, bpp_seed, win_len, win_step, min_bpp, cutoff, tab, cls):
        self.seq = synth_sequence(n, seed, first, reverse_total=self.total if self.reverse else 0)
        self.args = (bpp_seed, win_len, win_step, min_bpp, cutoff, {b: float(tab[k]) for k, b in enumerate("ACGU")}, cls)
        return len(self.py.windows(n, win_len, win_step)), 0, 0.0

    def scan_resident(self, w0, w1, copy=False):
        from rmdetect_b200._native import HIT_DTYPE
        bpp_seed, W, step, min_bpp, cutoff, gc, cls = self.args
        for oe, c in zip(self.oes, cls):
            oe.selected = c
        starts = self.py.windows(len(self.seq), W, step)
        bpp = [self.py.synth_bpp(self.seq[s:s + W], bpp_seed) for s in starts]
        hits, tries, tpw = self.py.scan_sequence(self.oms, self.oes, 0, self.seq, W, step, bpp, min_bpp, LN_UNKNOWN, cutoff, gc=gc)
        if self.min_score is not None:
            hits = [h for h in hits if h["score"] > self.min_score]
        out = np.zeros(len(hits), dtype=HIT_DTYPE)
        for k, h in enumerate(hits):
            out[k] = (int(np.searchsorted(starts, h["coords"][0])), h["coords"][0], 0, h["coords"][1] - h["coords"][0], h["p"][0], h["p"][1],
                      h["model"], h["cfg"], 1, h["prob_model"], h["prob_rand"], h["score"], h["evalue"])
        gate = np.zeros((len(starts), len(self.oms)), dtype=np.uint32)
        gate[:, 0] = tpw[:len(starts), 1]
        return dict(hits=out, gate_pass=gate, raw_count=np.zeros_like(gate), ms_scan=0.0, ms_total=0.0)


class _OracleEngine:
    def __init__(self, names):
        self.models = [load_model(n) for n in names]
        self.evalues = [load_evalues(n) for n in names]
        self.ctx = _OracleCtx(names)


def _genome_worker(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from rmdetect_b200.genome import scan_synthetic_genome
    names = ["TGA_1.0", "KT_1.0"]
    r = scan_synthetic_genome(_OracleEngine(names), 1300, True, 17, 150, 75, 0.001, LN_CUTOFF, rank=rank, world=world, dist=dist,
                              piece_windows=3, min_score=2.0)
    if rank == 0:
        one = scan_synthetic_genome(_OracleEngine(names), 1300, True, 17, 150, 75, 0.001, LN_CUTOFF, piece_windows=1000, min_score=2.0)
        keys = ("tests_all", "tests_pairs", "hits", "windows", "checksum", "counts")
        q.put(({k: r[k] for k in keys}, {k: one[k] for k in keys}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_a_genome_over_gloo():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_genome_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    two, one = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert two == one and one["hits"] > 0 and one["windows"] == 2 * 17 and sum(one["counts"]) == 1300


@pytest.mark.gpu
def test_fasta_genome_in_pieces_equals_the_scanner(tmp_path):
    """Real input through the genome driver: a FASTA file (lower case, T, N runs, IUPAC codes), both strands, cut into
    pieces of a few windows and shared by two ranks, gives the hits and tries of the drop-in `Scanner.scan` on the same
    sequences (lib/sequences.py:103-111 order: forward sequences, then the reverse complements under key + "--"); the
    device-resident table holds the same rows."""
    import random
    from helpers import Cfg, SeqList, fresh_models
    from rmdetect_b200.genome import read_fasta, reverse_complement, scan_genome
    from rmdetect_b200.hittable import HitTable
    from rmdetect_b200.scanner import Scanner
    from rmdetect_b200.synth import SynthFold
    rng = random.Random(5)
    chr1 = [rng.choice("ACGT") for _ in range(2600)]
    chr1[700:760] = "N" * 60
    for k in (50, 51, 1200, 2300):
        chr1[k] = rng.choice("RYKM")
    chr2 = "".join(rng.choice("acgu") for _ in range(420))
    path = tmp_path / "g.fasta"
    body = "".join(chr1)
    path.write_text(">chr1 test sequence\n" + "\n".join(body[k:k + 70] for k in range(0, len(body), 70)) + "\n\n>chr2\n" + chr2 + "\n")
    seqs = read_fasta(str(path))
    assert [k for k, _ in seqs] == ["chr1", "chr2"] and seqs[0][1] == body.replace("T", "U") and seqs[1][1] == chr2.upper()
    names = ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]
    models, evs = fresh_models(names)
    sc = Scanner(Cfg(models, evs))
    sc.rna_tools = SynthFold(9)
    both = [(k, s, None) for k, s in seqs] + [(k + "--", reverse_complement(s), None) for k, s in seqs]
    cands, tries = sc.scan(SeqList(both))
    want = sorted((c.key, c.coords[0], c.chains_pos[0][0], c.chains_pos[1][0], c.model.full_name(), c.score, c.evalue) for c in cands)
    eng = sc._engine
    eng.ctx.set_return_dropped(False)
    T = HitTable(eng)
    got, totals = [], []
    for rank in range(2):
        def on_hits(key, h, got=got):
            for r in h:
                got.append((key, int(r["start"]), int(r["start"] + r["p0"]), int(r["start"] + r["p1"]),
                            models[int(r["model"])].full_name(), float(r["score"]), float(r["evalue"])))
        totals.append(scan_genome(eng, seqs, SynthFold(9), both_strands=True, rank=rank, world=2, piece_windows=5, table=T, on_hits=on_hits))
    assert sorted(got) == want and len(want) > 20
    assert [totals[0][k] + totals[1][k] for k in ("tests_all", "tests_pairs")] == tries
    assert totals[0]["hits"] + totals[1]["hits"] == len(want)
    rows = T.rows()
    assert sorted((T.keys[r["key"]], int(r["start"]), int(r["start"] + r["p0"]), int(r["start"] + r["p1"]), models[int(r["model"])].full_name(),
                   float(r["score"]), float(r["evalue"])) for r in rows) == want
    # letters of the rows are the sequence's (N and IUPAC codes as "other")
    seq_of = {k: q for k, q, _ in both}
    for c in T.candidates():
        for letters, pos in zip(c.chains_seqs, c.chains_pos):
            for ch, p in zip(letters, pos):
                assert (ch == ".") == (p is None)
                if p is not None:
                    true = seq_of[c.key][p]
                    assert ch == (true if true in "ACGU" else "N")
    sc._engine.close()


def test_fasta_reader_follows_the_reference(tmp_path):
    """lib/file_io.py:5-24 + lib/sequences.py:103-111,138-150 on a small file (output of the reference's own Sequences
    class for this text, recorded here)."""
    from rmdetect_b200.genome import read_fasta, reverse_complement
    path = tmp_path / "t.fasta"
    path.write_text(">chr1 some text\nacgtnnACGU\nRYacg\n\n>chr2\nGGGTTT\n")
    seqs = read_fasta(str(path))
    assert seqs == [("chr1", "ACGUNNACGURYACG"), ("chr2", "GGGUUU")]
    assert [(k + "--", reverse_complement(q)) for k, q in seqs] == [("chr1--", "CGUYRACGUNNACGU"), ("chr2--", "AAACCC")]


# ---- world size 2 on CPU (gloo) for the real-input driver: pieces, strands, the closing all-reduce ----------------------
class _OracleJobEngine:
    """Plays `ScanEngine` for `scan_genome`: `build_job` keeps the piece, `ctx.scan` scans it with the CPU oracle (the
    host-side plumbing of the N > 1 path; the CUDA path is covered by test_fasta_genome_in_pieces_equals_the_scanner)."""

    def __init__(self, names):
        from oracle import pyoracle
        self.py = pyoracle
        self.models = [load_model(n) for n in names]
        self.evalues = [load_evalues(n) for n in names]
        self.oms = [pyoracle.OracleModel(m) for m in self.models]
        self.oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
        self.ctx = self
        self.min_score = None

    def set_min_score(self, v):
        self.min_score = v

    def build_job(self, seqs, win_len, win_step, fold, constraint, min_bpp, cutoff, gc=None):
        (key, seq, _), = seqs
        starts = window_starts(len(seq), win_len, win_step)
        return dict(seq=seq, gc=gc, fold=fold, n_win=len(starts), starts=[starts], win=(win_len, win_step), min_bpp=min_bpp, cutoff=cutoff)

    def scan(self, job, copy=False):
        from rmdetect_b200._native import HIT_DTYPE
        W, step = job["win"]
        seq, starts = job["seq"], job["starts"][0]
        for oe, ev in zip(self.oes, self.evalues):
            ev.select_gc(job["gc"])
            oe.selected = ev.selected
        bpp = [job["fold"].coo(seq[s:s + W]) for s in starts.tolist()]
        hits, tries, tpw = self.py.scan_sequence(self.oms, self.oes, 0, seq, W, step, bpp, job["min_bpp"], LN_UNKNOWN, job["cutoff"], gc=job["gc"])
        if self.min_score is not None:
            hits = [h for h in hits if h["score"] > self.min_score]
        out = np.zeros(len(hits), dtype=HIT_DTYPE)
        for k, h in enumerate(hits):
            out[k] = (int(np.searchsorted(starts, h["coords"][0])), h["coords"][0], 0, h["coords"][1] - h["coords"][0], h["p"][0], h["p"][1],
                      h["model"], h["cfg"], 1, h["prob_model"], h["prob_rand"], h["score"], h["evalue"])
        gate = np.zeros((len(starts), len(self.oms)), dtype=np.uint32)
        gate[:, 0] = tpw[:len(starts), 1]
        return dict(hits=out, gate_pass=gate, raw_count=np.zeros_like(gate), ms_scan=0.0, ms_total=0.0)


def _fasta_worker(rank, world, port, q):
    import os
    import random
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from rmdetect_b200.genome import scan_genome
    from rmdetect_b200.synth import SynthFold
    rng = random.Random(3)
    seqs = [("chrA", "".join(rng.choice("ACGU") for _ in range(1100))), ("chrB", "".join(rng.choice("ACGUN") for _ in range(400)))]
    names = ["TGA_1.0", "KT_1.0"]
    got = []
    r = scan_genome(_OracleJobEngine(names), seqs, SynthFold(4), both_strands=True, rank=rank, world=world, dist=dist, piece_windows=3,
                    min_score=2.0, on_hits=lambda key, h: got.extend((key, int(x["start"]), int(x["p0"]), int(x["p1"]), int(x["model"])) for x in h))
    gathered = [None] * world
    dist.all_gather_object(gathered, got)
    if rank == 0:
        one_hits = []
        one = scan_genome(_OracleJobEngine(names), seqs, SynthFold(4), both_strands=True, piece_windows=1000, min_score=2.0,
                          on_hits=lambda key, h: one_hits.extend((key, int(x["start"]), int(x["p0"]), int(x["p1"]), int(x["model"])) for x in h))
        keys = ("tests_all", "tests_pairs", "hits", "windows", "checksum", "strands")
        q.put(({k: r[k] for k in keys}, {k: one[k] for k in keys}, sorted(sum(gathered, [])), sorted(one_hits)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_fasta_sequences_over_gloo():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fasta_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    two, one, hits2, hits1 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert two == one and hits2 == hits1 and one["hits"] > 0 and one["strands"] == 4
