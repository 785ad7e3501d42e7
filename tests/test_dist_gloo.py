"""World-size-2 test of the sharding + gather logic (rmdetect_b200/dist.py) on CPU with gloo.

The per-rank scan is played by the CPU oracle (this is a test of the host-side plumbing: chunk
planning with halo, coordinate restoration, trimming of halo windows, variable-length gather); the
GPU scan itself is covered by tests/test_gpu_parity.py.
"""
import math
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_fns(names, seed, win_len, win_step, min_bpp, unknown, cutoff):
    """(build_job_fn, scan_fn) with the shape ScanEngine gives, backed by the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import evalues_path, load_model
    from oracle import pyoracle
    from rmdetect_b200._native import HIT_DTYPE
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]

    def build(seqs, gc):
        return dict(seqs=seqs, gc=gc)

    def scan(job):
        (key, seq, _), = job["seqs"]
        starts = pyoracle.windows(len(seq), win_len, win_step)
        bpp = [pyoracle.synth_bpp(seq[s:s + win_len], seed) for s in starts]
        hits, tries, tpw = pyoracle.scan_sequence(oms, oes, 0, seq, win_len, win_step, bpp, min_bpp, unknown, cutoff,
                                                  gc=job["gc"])
        out = np.zeros(len(hits), dtype=HIT_DTYPE)
        for k, h in enumerate(hits):
            w = int(np.searchsorted(starts, h["coords"][0]))          # window index from its start
            out[k] = (w, h["coords"][0], 0, h["coords"][1] - h["coords"][0], h["p"][0], h["p"][1], h["model"], h["cfg"], 1,
                      h["prob_model"], h["prob_rand"], h["score"], h["evalue"])
        gate = np.zeros((len(starts), len(names)), dtype=np.uint32)
        gate[:, 0] = tpw[:len(starts), 1]                             # per-window totals are all the trimming needs
        return dict(hits=out, gate_pass=gate)

    return build, scan, [m.model for m in oms]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rmdetect_b200.dist import gather_hits, plan_chunks, scan_chunk
    from rmdetect_b200.gcx import GC
    from rmdetect_b200.synth import synth_sequence
    names = ["TGA_1.0", "KT_1.0"]
    W, step, seed = 150, 75, 11
    unknown, cutoff = math.log(1e-4), math.log(0.1)
    seq = synth_sequence(1730, 123)
    gc = GC.compute(seq)
    build, scan, models = _oracle_fns(names, seed, W, step, 0.001, unknown, cutoff)
    plan, starts = plan_chunks(len(seq), W, step, world)
    hits, tries = scan_chunk(scan, build, "chr", seq, None, gc, W, step, plan[rank], models)
    all_hits, total = gather_hits(hits, tries, dist)
    if rank == 0:
        whole, wtries = scan_chunk(scan, build, "chr", seq, None, gc, W, step, plan_chunks(len(seq), W, step, 1)[0][0], models)
        q.put((all_hits.tobytes(), total, whole.tobytes(), wtries, len(starts)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_chunk_scan_equals_single_scan():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, total, want, wtries, n_win = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from rmdetect_b200._native import HIT_DTYPE
    got = np.frombuffer(got, dtype=HIT_DTYPE)
    want = np.frombuffer(want, dtype=HIT_DTYPE)
    assert n_win == 23 and len(want) > 10
    assert total == wtries
    assert got.tobytes() == want.tobytes()          # same hits, same order, same coordinates, seam duplicates gone


def test_plan_chunks_geometry():
    from rmdetect_b200.dist import halo_windows, plan_chunks
    assert halo_windows(150, 75) == 2 and halo_windows(150, 10) == 15
    plan, starts = plan_chunks(1000, 150, 75, 4)
    assert [p[:2] for p in plan] == [(0, 3), (3, 6), (6, 9), (9, 13)]
    assert plan[0][2] == 5 and plan[3][2] == 13
    assert plan[0][3:] == (0, 4 * 75 + 150) and plan[3][3:] == (9 * 75, 1000)
    plan, _ = plan_chunks(100, 150, 75, 4)           # one window: ranks 0..2 get nothing
    assert [p[1] - p[0] for p in plan] == [0, 0, 0, 1]
