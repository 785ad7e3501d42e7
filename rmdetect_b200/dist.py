"""Multi-GPU scans: one process per GPU, windows sharded in contiguous chunks with a halo.

The scan path shards by window: a window needs only its own letters, its own bpp
matrix and the GC table of its whole sequence.  A long sequence is cut into per-rank chunks whose
boundaries fall on window starts; each chunk is scanned as a sequence of its own with the parent's GC
table, so window geometry inside the chunk equals the parent's (reference lib/scanner.py:80-111).
The duplicate rule (lib/scanner.py:216-244) looks at later overlapping windows, so every chunk but
the last carries `halo` extra windows that are scanned only to make that decision and then dropped.
There is no collective on the data path; hit lists (small) are gathered on rank 0 with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from ._native import HIT_DTYPE
from .scanner import placements_in_window, window_starts


def halo_windows(win_len, win_step):
    """Later windows that can overlap a given one: starts closer than win_len, plus the re-anchored tail."""
    return -(-win_len // win_step)


def plan_chunks(seq_len, win_len, win_step, world):
    """Per rank: (first window, one past its last own window, one past its last scanned window,
    letter range [lo, hi) of the chunk).  Ranks beyond the window count get empty chunks."""
    starts = window_starts(seq_len, win_len, win_step)
    n_win = len(starts)
    halo = halo_windows(win_len, win_step)
    plan = []
    for r in range(world):
        a, b = n_win * r // world, n_win * (r + 1) // world
        bh = min(n_win, b + halo) if b > a else b
        if b > a:
            lo = int(starts[a])
            hi = seq_len if bh == n_win else int(starts[bh - 1]) + win_len
        else:
            lo = hi = 0
        plan.append((a, b, bh, lo, hi))
    return plan, starts


def scan_chunk(scan_fn, build_job_fn, key, seq, col_index, gc, win_len, win_step, plan_entry, models):
    """Scan one rank's chunk and return (hits in parent coordinates, [tests_all, tests_pairs]).

    `build_job_fn(seqs, gc)` and `scan_fn(job)` are ScanEngine.build_job / ScanEngine.scan_job with the
    remaining arguments bound; tests substitute CPU stand-ins.
    """
    a, b, bh, lo, hi = plan_entry
    if b <= a:
        return np.empty(0, dtype=HIT_DTYPE), [0, 0]
    chunk = seq[lo:hi]
    job = build_job_fn([(key, chunk, None)], gc)
    res = scan_fn(job)
    hits = np.array(res["hits"], dtype=HIT_DTYPE, copy=True)
    own = hits["window"] < (b - a)                     # halo windows only served the duplicate rule
    hits = hits[own]
    hits["window"] += a
    hits["start"] += lo
    n_own = b - a
    tests_pairs = int(np.asarray(res["gate_pass"])[:n_own].sum())
    starts = window_starts(len(chunk), win_len, win_step)[:n_own]
    tests_all = sum(placements_in_window(m, min(win_len, len(chunk) - int(s))) for s in starts for m in models)
    return hits, [tests_all, tests_pairs]


def gather_hits(hits, tries, dist=None, dst=0):
    """Concatenate per-rank hit lists in rank order on `dst`; sums `tries` everywhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return hits, list(tries)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = torch.tensor([len(hits), tries[0], tries[1]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)                                  # counts: the sizes of the variable-length gather
    counts = [int(m[0]) for m in metas]
    total = [int(sum(int(m[1]) for m in metas)), int(sum(int(m[2]) for m in metas))]
    cap = max(max(counts), 1)
    buf = torch.zeros(cap * HIT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    if len(hits):
        raw = torch.from_numpy(np.frombuffer(hits.tobytes(), dtype=np.uint8).copy())
        buf[:raw.numel()] = raw.to(dev)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst)                           # padded gather of the (small) hit lists
    if rank != dst:
        return np.empty(0, dtype=HIT_DTYPE), total
    parts = [np.frombuffer(g.cpu().numpy().tobytes(), dtype=HIT_DTYPE, count=c) for g, c in zip(gathered, counts)]
    return np.concatenate(parts) if parts else np.empty(0, dtype=HIT_DTYPE), total


def bind_host_thread_to_gpu(device):
    """Pin the calling thread (and the page-locked buffers it allocates from now on) to the CPUs next to GPU
    `device` (NVML's ideal affinity).  With one process per GPU on a two-socket host this keeps every rank's
    host-to-device copies off the inter-socket link; returns the previous affinity set (restore it with
    os.sched_setaffinity(0, previous)), or None when NVML is not available."""
    import os
    try:
        import pynvml
        import torch
        props = torch.cuda.get_device_properties(device)
        pynvml.nvmlInit()
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        previous = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return previous
    except Exception:
        return None
