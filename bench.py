#!/usr/bin/env python
"""Benchmark of the scan hot path: nt·models scanned per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mb M]

Workload = BASELINE.json configs[2]: a synthetic random chromosome of M Mb (default 10) per GPU,
windows 150/75, all four shipped models, stand-in bpp matrices (rmdetect_b200/synth.py; ViennaRNA
is external to the path and its time is not part of the metric; generation is reported apart).
One step = one full scan of the chromosome: enumeration + gate + scoring + E-values + ordering +
duplicate removal + download of the hit list.

  value  inputs (2-bit sequence, bpp lists) resident in HBM before the timed region
  e2e    the same scan through the C ABI call with HOST buffers (rmd_scan): H2D of sequence and bpp
         lists and D2H of the hits inside the timed region
N > 1: one process per GPU (torchrun), every rank scans its own chromosome chunk (weak scaling),
hit counts are all-gathered over NCCL, time = max over ranks.
`--impl reference` times the CPU oracle port (oracle/, the reference's algorithm restated in C,
OpenMP over windows on all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261019
WIN_LEN, WIN_STEP = 150, 75
MODEL_FILES = ["cl_1", "tga_1", "gb_1", "kt_1"]         # the order os.listdir gave the reference in the build container
DATA = os.path.join(ROOT, "tests", "golden", "ref_data", "models")
LN_UNKNOWN, LN_CUTOFF, MIN_BPP = math.log(1e-4), math.log(0.1), 0.001
CAL_SAMPLES = 1 << 20                                    # calibration samples per model, GPU and round
CAL_ROUNDS = 8                                           # timed rounds of one batch per model (a round takes a millisecond)
ISSUE_SLOTS_PER_SM_CYCLE = 4                             # one warp instruction per scheduler and cycle
WORKLOAD = "config3: synthetic %d Mb random chromosome per GPU, win 150/75, models CL,TGA,GB,KT, synthetic bpp (gate on)"


def ncu_summary(kernel="scan_fast"):
    """DRAM bytes and warp instructions of one launch of the dominant kernel over the 10 Mb workload, read from the
    newest committed `ncu --set full` summary under profiles/ (metric,unit,value rows; tools/ncu_metrics.py)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_%s_ncu_full.csv" % kernel)))
    if not files:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
    vals = {}
    for line in open(files[-1]):
        t = line.strip().split(",")
        if len(t) == 3 and t[1] in scale:
            try:
                vals[t[0]] = float(t[2]) * scale[t[1]]
            except ValueError:
                pass
    if "dram__bytes_read.sum" not in vals:
        return None
    pct = {}
    for line in open(files[-1]):
        t = line.strip().split(",")
        if len(t) == 3 and t[1] == "%":
            try:
                pct[t[0]] = float(t[2])
            except ValueError:
                pass
    return {"file": os.path.relpath(files[-1], ROOT), "pct": pct,
            "dram_bytes": vals["dram__bytes_read.sum"] + vals.get("dram__bytes_write.sum", 0.0),
            "warp_instructions": vals.get("smsp__inst_executed.sum"), "ms": vals.get("gpu__time_duration.sum")}


def load_models():
    from rmdetect_b200.evalues import parse_evalues
    from rmdetect_b200.modelfile import parse_model
    models = [parse_model(os.path.join(DATA, n + ".model")) for n in MODEL_FILES]
    evs = [parse_evalues(os.path.join(DATA, n + ".evalues")) for n in MODEL_FILES]
    return models, evs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([t.strip() for t in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cal_roofline(kernel_ms, samples, n_class):
    """calibrate_engine_kernel (csrc/calib_kernels.cuh): one tree walk per sample; the 165 weights per sample are not computed one
    by one (samples without a solution are counted by letter composition, their weights follow from 165 x 1771 products), so
    the kernel is bound by instruction issue in the tree walk.  The algorithmic work is counted in (sample, class) pairs, as
    the reference does it; the pipe figures come from the committed ncu capture of one launch of 2^20 samples
    (profiles/r*_calibrate_ncu_full.csv), the rate is measured live."""
    ncu = ncu_summary("calibrate")
    out = {"bound": "issue (tree walk)", "achieved": samples * n_class / (kernel_ms * 1e-3), "unit": "(sample, class) pairs/s"}
    if ncu:
        p = ncu["pct"]
        out.update({"source": ncu["file"],
                    "fp64_pipe_pct_of_peak": p.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": p.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "ncu_launch_ms": ncu.get("ms"), "dram_bytes": ncu.get("dram_bytes")})
    return out


def chromosome_counts(n, first=0):
    """(A, C, G, U) counts of the synthetic chromosome: its GC table is taken over the whole sequence (lib/scanner.py:74-75)."""
    import numpy as np
    from rmdetect_b200.synth import synth_codes
    counts = np.zeros(4, dtype=np.int64)
    for a in range(0, n, 1 << 22):
        counts += np.bincount(synth_codes(min(1 << 22, n - a), SEED, start=first + a), minlength=4)
    return counts


def cpu_baseline(models, n_threads, target_s, n_total, counts):
    """Oracle port on the host cores over a bounded sample: a prefix of the chromosome sized, from a pilot run, to take
    about `target_s` seconds.  The stand-in bpp lists are generated outside the timed scan (as on the GPU arm, where
    they are inputs), and the GC table is the whole chromosome's, so the sample's counts and hit checksum can be
    compared with the GPU's for the same windows.  Returns a dict."""
    from oracle import pyoracle
    from rmdetect_b200.synth import synth_sequence
    max_win = 1 + -(-(n_total - WIN_LEN) // WIN_STEP)

    def run(nwin):
        length = n_total if nwin >= max_win else (nwin - 1) * WIN_STEP + WIN_LEN + WIN_STEP     # never re-anchors a tail window inside the sample
        seq = synth_sequence(length, SEED)
        return pyoracle.bench_scan2(models, seq, WIN_LEN, WIN_STEP, 0, nwin, SEED, n_threads, MIN_BPP, LN_UNKNOWN, LN_CUTOFF,
                                    counts=counts)

    pilot = min(max_win, 64 * n_threads)
    run(pilot)                                                                   # warm the threads
    r = run(pilot)
    nwin = int(max(pilot, min(max_win, pilot * target_s / max(r["s_scan"], 1e-3))))
    r = run(nwin)
    r.update(windows=nwin, value=nwin * WIN_STEP * len(models) / r["s_scan"], whole=nwin >= max_win)
    return r


def cal_cpu_baseline(models, cores, target_s):
    """The oracle's restatement of rmcalibrate.py:265-295 on all host cores (one model instance per thread; ctypes
    releases the GIL), device-twin samples, for about `target_s` seconds: the calibration metric's CPU baseline."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    from rmdetect_b200.gcx import GC
    from rmdetect_b200.synth import synth_calibration_samples
    classes = GC.get_classes(5, 15, 85)
    m = models[-1]                                                               # KT_1.0: L = 16
    L = len(m.nodes)

    def work(k, n):
        om = pyoracle.OracleModel(m)
        rows = synth_calibration_samples(SEED, k * n, n, L)
        strings = ["".join("ACGU"[c] for c in row) for row in rows]
        t0 = time.perf_counter()
        pyoracle.calibrate(om, strings, classes, LN_UNKNOWN, LN_CUTOFF, 512)
        return time.perf_counter() - t0

    pilot = work(0, 2000)
    per = int(max(2000, min(200000, 2000 * target_s / max(pilot, 1e-3))))
    with ThreadPoolExecutor(cores) as pool:
        secs = list(pool.map(lambda k: work(k, per), range(cores)))
    return {"value": per * cores / max(secs), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "oracle/rmd_oracle.c orc_calibrate, %s, %d samples on each of %d threads, %.1f s"
                      % (m.full_name(), per, cores, max(secs))}


def python_reference(cores):
    """The reference's own Python `Scanner.scan` and `rmcalibrate.py` loop on this host (oracle/ref_bench.py on the
    modules staged in oracle/_ref).  A subprocess: it forks workers, this process holds CUDA."""
    script = os.path.join(ROOT, "oracle", "ref_bench.py")
    try:
        out = subprocess.run([sys.executable, script, "--windows", "20", "--procs", str(min(cores, 64)), "--cal-samples", "4000"],
                             capture_output=True, text=True, timeout=600)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:                                                      # a baseline that cannot run is reported, not fatal
        return {"unavailable": "%s: %s" % (type(exc).__name__, exc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    models, _ = load_models()
    cores = os.cpu_count() or 1
    n = args.mb * 1_000_000
    counts = chromosome_counts(n)
    vals = []
    for k in range(args.warmup + args.steps):
        r = cpu_baseline(models, cores, 4.0 if k >= args.warmup else 1.0, n, counts)
        if k >= args.warmup:
            vals.append(r)
    value = sum(r["value"] for r in vals) / len(vals)
    ms = 1e3 * sum(r["s_scan"] for r in vals) / len(vals)
    r = vals[-1]
    line = {
        "impl": "reference", "metric": "nt·models scanned/sec", "value": value, "unit": "nt·model/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.mb},
        "cpu_baseline": {"value": value, "unit": "nt·model/s", "cores": cores, "kind": "port",
                         "sample": "oracle/rmd_oracle.c (C restatement of the reference's Python scan), OpenMP over the first "
                                   "%d windows (%d nt) of the chromosome per step, whole-chromosome GC table; stand-in bpp "
                                   "lists generated outside the timed scan (%.2f s per step)"
                                   % (r["windows"], r["windows"] * WIN_STEP, r["s_generate"]),
                         "tries": r["tries"], "hits_before_dedupe": r["hits"], "hit_checksum": "%016x" % r["checksum"],
                         "python_reference": python_reference(cores)},
        "e2e": {"value": value, "unit": "nt·model/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    from rmdetect_b200 import _native
    from rmdetect_b200.gcx import gc_table
    from rmdetect_b200.scanner import ScanEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scan path has no CPU fallback")
    torch.cuda.set_device(local)
    from rmdetect_b200.dist import bind_host_thread_to_gpu
    affinity = bind_host_thread_to_gpu(local) if world > 1 and not os.environ.get("RMD_NO_BIND") else None   # host buffers next to the rank's GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    models, evs = load_models()
    eng = ScanEngine(models, evs, LN_UNKNOWN, device=local)
    ctx = eng.ctx
    n = args.mb * 1_000_000
    first = rank * n                                            # every rank scans its own chunk of the stream
    neutral = gc_table({b: math.log(0.25) for b in "ACGU"})
    n_win, n_bpp, ms_gen = ctx.synth_job(n, first, SEED, SEED, WIN_LEN, WIN_STEP, MIN_BPP, LN_CUTOFF, neutral)
    counts = ctx.gc_counts(1)[0]                                # K2: letter counts on the device, logs on the host
    gc = {b: math.log(float(counts[k]) / n) for k, b in enumerate("ACGU")}
    cls = []
    for ev in evs:
        ev.select_gc(gc)
        cls.append(ev.selected)
    ctx.set_gc(gc_table(gc).reshape(1, -1), np.array([cls], dtype=np.int32))
    units = n * len(models)

    # ---- device-resident scan ------------------------------------------------------------------
    ctx.set_hit_format(True)                                    # hits come back as 32-byte records (rmd_hit32), as in the e2e arm
    for _ in range(args.warmup):
        ctx.scan_resident(0, n_win, copy=False)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = scan_ms = post_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        r = ctx.scan_resident(0, n_win, copy=False)
        dev_ms += r["ms_total"]; scan_ms += r["ms_scan"]; post_ms += r["ms_post"]; launches += r["n_launches"]
    barrier()
    wall = time.perf_counter() - t0
    n_hits, n_raw, tries = len(r["hits"]), r["n_raw"], r["tries"]
    ctx.set_hit_format(False)
    # one untimed scan that also returns the hits the duplicate rule removed: what the CPU baseline's sample is checked against
    ctx.set_return_dropped(True)
    full = ctx.scan_resident(0, n_win, copy=True)
    ctx.set_return_dropped(False)
    assert full["n_raw"] == n_raw == len(full["hits"]) and int(full["hits"]["keep"].sum()) == n_hits

    # ---- end to end through the C ABI with host buffers --------------------------------------------
    off, bi, bj, bp = ctx.download_job(n_win, n_bpp, pinned=True)
    codes = _native.pinned_array(n, np.uint8)
    codes[:] = ctx.download_codes(n)
    job = dict(seq_off=np.array([0, n], dtype=np.int64), codes=codes, gc_log=gc_table(gc).reshape(1, -1),
               gc_class=np.array([cls], dtype=np.int32), win_len=WIN_LEN, win_step=WIN_STEP, n_win=n_win,
               bpp_off=off, bpp_i=bi, bpp_j=bj, bpp_p=bp, min_bpp_mean=MIN_BPP, cutoff=LN_CUTOFF)
    # transport form of the lists: 6 bytes per entry (i | j << 8, sqrt(p) * 1e9 as RNAfold prints it, plus the ulp
    # of libm's pow) instead of 12; rebuilt to the same doubles on the device (rmd_compact_bpp)
    ij, q = _native.compact_bpp(bi, bj, bp, pinned=True)
    cjob = dict(job, bpp_i=None, bpp_j=None, bpp_p=None, bpp_ij=ij, bpp_q=q)

    def timed_e2e(j):
        for _ in range(max(1, args.warmup // 2)):
            r = ctx.scan(j, copy=False)
        barrier()
        t1 = time.perf_counter()
        dev = 0.0
        for _ in range(args.steps):
            r = ctx.scan(j, copy=False)
            dev += r["ms_total"]
        barrier()
        return time.perf_counter() - t1, dev, r

    f64_wall, f64_dev_ms, e = timed_e2e(job)                    # fp64 lists on the wire (12 bytes per entry), 72-byte hit records back
    assert e["tries"] == tries and len(e["hits"]) == n_hits
    ctx.set_hit_format(True)                                    # 32-byte hit records on the way back (rmd_hit32)
    e2e_wall, e2e_dev_ms, e = timed_e2e(cjob)                   # compact lists on the wire: the headline e2e
    ctx.set_hit_format(False)
    kept = full["hits"][full["hits"]["keep"] != 0]
    for f in ("window", "p0", "p1", "model", "leaf", "score", "evalue"):
        assert np.array_equal(e["hits"][f], kept[f]), "e2e hit list differs from the device-resident scan in " + f
    h2d_f64 = codes.nbytes + off.nbytes + n_bpp * (2 + 2 + 8) + 16 * 8 + 4 * 4
    h2d = codes.nbytes + off.nbytes + n_bpp * (2 + 4) + 16 * 8 + 4 * 4
    clocks = sampler.summary()                                  # sampled over both timed regions
    assert e["tries"] == tries and len(e["hits"]) == n_hits
    d2h = len(e["hits"]) * e["hits"].dtype.itemsize + 2 * n_win * len(models) * 4 + 64

    # ---- what the host can feed: every rank copies a pinned buffer to its GPU at the same time ----------------
    probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty_like(probe_h, device="cuda")
    probe_d.copy_(probe_h, non_blocking=True)
    barrier()
    t_c = time.perf_counter()
    for _ in range(8):
        probe_d.copy_(probe_h, non_blocking=True)
    barrier()
    h2d_probe_s = time.perf_counter() - t_c
    del probe_h, probe_d

    # ---- second metric: calibration samples per second (BASELINE.json configs[3], throughput mode) ----
    # every rank scores its own CAL_SAMPLES device-drawn samples per model against the 165 GC classes
    # (a) one batch per model and GPU (kernel + histogram download); (b) the reference's loop (rmcalibrate.py:235-298):
    # blocks of 10 000 samples per GPU, the 165 x 512 fp64 histograms all-reduced (NCCL) after every block and the
    # reference's convergence rule run on the reduced table; min(samples, 4^L) samples per model (TGA: 65 536)
    from rmdetect_b200.calibrate import Calibrator, calibrate_model, class_logs
    from rmdetect_b200.gcx import GC
    cl = class_logs(GC.get_classes(5, 15, 85))
    hists = [np.full((len(cl), 512), 0.0) for _ in models]     # written once: the accumulating histograms are resident pages, as in a run that has been going
    for mi, m in enumerate(models):                             # warm-up: the same batch size (buffers are sized by it), other samples
        ctx.calibrate(mi, None, cl, LN_CUTOFF, np.zeros_like(hists[mi]), seed=SEED + 1, first=0, n=CAL_SAMPLES, L=len(m.nodes))
    barrier()
    t2 = time.perf_counter()
    cal_kernel_ms, cal_calls_ms = 0.0, [0.0] * len(models)
    for rnd in range(CAL_ROUNDS):                               # every round scores fresh samples of the stream
        for mi, m in enumerate(models):
            tc = time.perf_counter()
            _, ms_k = ctx.calibrate(mi, None, cl, LN_CUTOFF, hists[mi], seed=SEED, first=(rnd * world + rank) * CAL_SAMPLES,
                                    n=CAL_SAMPLES, L=len(m.nodes))
            cal_calls_ms[mi] += 1e3 * (time.perf_counter() - tc) / CAL_ROUNDS
            cal_kernel_ms += ms_k / CAL_ROUNDS
    barrier()
    cal_wall = time.perf_counter() - t2
    cal_done, cal_blocks, cal_loop_kernel_ms, cal_sum = 0, 0, 0.0, 0.0
    barrier()
    t3 = time.perf_counter()
    for mi, m in enumerate(models):
        st = {}
        _, _, done = calibrate_model(m, None, samples=CAL_SAMPLES * world, seed=SEED, device_samples=True, log=None, dist=dist,
                                     scorer=Calibrator(m, LN_UNKNOWN, LN_CUTOFF, ctx=ctx, model_index=mi),
                                     save_every_block=False, stats=st)
        cal_done += done; cal_blocks += st["blocks"]; cal_loop_kernel_ms += st["ms_kernel"]
        cal_sum += float(st["hist"].sum())                      # the reduced table is the same on every rank
    barrier()
    cal_loop_wall = time.perf_counter() - t3

    # ---- reduce over ranks: time = max, work = sum ---------------------------------------------------
    t = torch.tensor([wall, e2e_wall, dev_ms, e2e_dev_ms, cal_wall, cal_kernel_ms, f64_wall, cal_loop_wall, cal_loop_kernel_ms,
                      h2d_probe_s], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_hits, n_raw, tries[0], tries[1]], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(gathered, cnt)                      # the only collective: small per-rank hit counts
        cnt = torch.stack(gathered).sum(0)
    wall, e2e_wall, dev_ms, e2e_dev_ms, cal_wall, cal_kernel_ms, f64_wall, cal_loop_wall, cal_loop_kernel_ms, h2d_probe_s = t.tolist()
    h2d_ceiling = 8 * (256 << 20) * world / h2d_probe_s / 1e9    # GB/s, all ranks together
    total_units = units * world
    value = total_units * args.steps / wall
    e2e_value = total_units * args.steps / e2e_wall

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # algorithmic bytes of one scan-kernel launch: 2-bit letters per window (38 B + offsets), the bpp lists
        # (12 B per entry + 8 B offset per window), hit records out (32 B) and the two per-group counters
        algo = n_win * (40 + 8 + 16) + n_bpp * 12 + n_raw * 32 + n_win * len(models) * 8
        ms_kernel = scan_ms / args.steps
        achieved = algo / (ms_kernel * 1e-3) / 1e9
        issue = None
        ncu = ncu_summary() if args.mb == 10 else None          # the committed --set full capture is of the 10 Mb launch
        if ncu and ncu.get("warp_instructions") and clocks.get("sm_max_mhz"):
            sm_count = torch.cuda.get_device_properties(local).multi_processor_count
            issue_peak = sm_count * ISSUE_SLOTS_PER_SM_CYCLE * clocks["sm_max_mhz"] * 1e6
            issue_rate = ncu["warp_instructions"] / (ms_kernel * 1e-3)
            issue = {"achieved": issue_rate, "peak": issue_peak, "unit": "warp instructions/s", "frac": issue_rate / issue_peak,
                     "source": "%s (smsp__inst_executed.sum)" % ncu["file"],
                     # the pipe nearest its limit in that capture: integer / logic instructions take the ALU pipe two cycles each
                     "alu_pipe_pct_of_peak": ncu["pct"].get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     "fp64_pipe_pct_of_peak": ncu["pct"].get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                     "lsu_pipe_pct_of_peak": ncu["pct"].get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")}
        if affinity is not None:
            os.sched_setaffinity(0, affinity)                    # the CPU baseline may use every core again
        cores = os.cpu_count() or 1
        cpu = cpu_baseline(models, cores, 15.0, n, counts[:4])
        # the sample is a prefix of rank 0's chromosome with the same GC table: its counts and hits must be the GPU's
        cw = cpu["windows"]
        fh = full["hits"][full["hits"]["window"] < cw]
        from rmdetect_b200.genome import hit_checksum
        gpu_sample = {"tries": [cw * int(tries[0] // n_win), int(full["gate_pass"][:cw].sum())], "hits": int(len(fh)),
                      "checksum": hit_checksum(0, fh["start"], fh["p0"], fh["p1"], fh["model"], fh["leaf"], fh["score"]) if len(fh) else 0}
        if (cpu["tries"], cpu["hits"], cpu["checksum"]) != (gpu_sample["tries"], gpu_sample["hits"], gpu_sample["checksum"]):
            raise SystemExit("bench.py: the GPU scan disagrees with the CPU oracle on the first %d windows: oracle %s, GPU %s"
                             % (cw, {k: cpu[k] for k in ("tries", "hits", "checksum")}, gpu_sample))
        cal_cpu = cal_cpu_baseline(models, cores, 6.0)
        pyref = python_reference(cores)
        line = {
            "metric": "nt·models scanned/sec", "value": value, "unit": "nt·model/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.mb,
                       "windows_per_gpu": n_win, "bpp_entries_per_gpu": n_bpp, "placements_per_step": int(cnt[2]),
                       "gate_passes_per_step": int(cnt[3]), "hits_before_dedupe": int(cnt[1]), "hits": int(cnt[0]),
                       "l2": "inputs larger than L2 (%.0f MB of bpp lists per step)" % (n_bpp * 12 / 1e6),
                       "timing": "wall clock around K synchronous steps bracketed by barrier+synchronize, max over ranks; "
                                 "device_ms_per_step is the CUDA-event time of the same steps",
                       "device_ms_per_step": dev_ms / args.steps, "scan_kernel_ms": ms_kernel,
                       "post_kernels_ms": post_ms / args.steps, "bpp_generation_ms": ms_gen},
            "e2e": {"value": e2e_value, "unit": "nt·model/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_wall / args.steps, "device_ms_per_step": e2e_dev_ms / args.steps,
                    "api": "rmd_scan (C ABI) on pinned host buffers, bpp lists in the 6-byte transport form (rmd_compact_bpp), "
                           "hits back as 32-byte records (rmd_set_hit_format)",
                    "h2d_gbs": h2d * world * args.steps / e2e_wall / 1e9,
                    "h2d_ceiling_gbs": h2d_ceiling,
                    "h2d_ceiling_note": "all ranks copying 8 x 256 MB from pinned memory to their GPUs at the same time "
                                        "(plain cudaMemcpyAsync): what the host can feed; h2d_gbs is what this step asks of it",
                    "fp64_transport": {"value": total_units * args.steps / f64_wall, "h2d_bytes_per_step": int(h2d_f64),
                                       "ms_per_step": 1e3 * f64_wall / args.steps}},
            "calibration": {"metric": "calibration samples scored/sec", "value": CAL_SAMPLES * CAL_ROUNDS * len(models) * world / cal_wall,
                            "unit": "samples/s", "samples_per_model_per_gpu": CAL_SAMPLES, "rounds": CAL_ROUNDS, "classes": int(len(cl)),
                            "ms_total": 1e3 * cal_wall / CAL_ROUNDS, "kernel_ms": cal_kernel_ms,
                            "ms_per_call": [round(x, 4) for x in cal_calls_ms],
                            "note": "rmd_calibrate_synth: per round one batch of device-drawn samples per model and GPU, scored against "
                                    "165 GC classes (weighted histograms), histograms downloaded per call; ms_total and "
                                    "kernel_ms are per round (four calls)",
                            "reference_loop": {
                                "value": cal_done / cal_loop_wall, "unit": "samples/s", "samples": int(cal_done),
                                "blocks": int(cal_blocks), "ms_total": 1e3 * cal_loop_wall, "kernel_ms": cal_loop_kernel_ms,
                                "reduced_table_sum": cal_sum,
                                "note": "rmcalibrate.py:235-298 as calibrate_model runs it: blocks of 10 000 samples per GPU, "
                                        "histograms all-reduced over the GPUs after every block, the reference's convergence "
                                        "rule on the reduced table, min(samples, 4^L) samples per model"},
                            "cpu_baseline": cal_cpu,
                            "python_reference": pyref.get("calibrate_1core") if isinstance(pyref, dict) else None,
                            "roofline": cal_roofline(cal_kernel_ms, CAL_SAMPLES * len(models), len(cl))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu["dram_bytes"] if ncu else None,
                         "traffic_source": "%s (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full of this command)"
                                           % (ncu["file"] if ncu else "none"),
                         "kernel": "scan_fast_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo),
                         "note": "table-lookup / fp64-compare work: issue-bound, not HBM-bound (see DESIGN.md); `issue` is the "
                                 "binding resource: warp instructions of the launch (ncu) / its live duration against "
                                 "148 SMs x 4 schedulers x the SM clock",
                         "issue": issue},
            "cpu_baseline": {"value": cpu["value"], "unit": "nt·model/s", "cores": cores, "kind": "port",
                             "sample": "first %d windows (%d nt) of the same chromosome with its GC table, oracle/rmd_oracle.c with "
                                       "OpenMP on all cores, %.1f s of scanning (stand-in bpp lists generated before, %.1f s); "
                                       "tests_all, tests_pairs, hit count and hit checksum of these windows equal the GPU's"
                                       % (cw, cw * WIN_STEP, cpu["s_scan"], cpu["s_generate"]),
                             "covers_whole_chromosome": bool(cpu["whole"]), "tries": cpu["tries"],
                             "hits_before_dedupe": cpu["hits"], "hit_checksum": "%016x" % cpu["checksum"],
                             "python_reference": {k: v for k, v in pyref.items() if k != "calibrate_1core"} if isinstance(pyref, dict) else pyref},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def run_genome(args):
    """BASELINE.json configs[4]: a synthetic genome of --gnt G letters, both strands, all four models, sharded by
    piece + halo over the ranks (rmdetect_b200/genome.py).  One step = the whole genome; sequence and stand-in bpp
    matrices are generated on the device piece by piece, their time is reported apart."""
    import torch
    from rmdetect_b200.genome import scan_synthetic_genome
    from rmdetect_b200.scanner import ScanEngine
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scan path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    models, evs = load_models()
    eng = ScanEngine(models, evs, LN_UNKNOWN, device=local)
    total = int(args.gnt * 1e9)
    piece = int(args.piece_mb * 1e6) // WIN_STEP
    scan_synthetic_genome(eng, min(total, piece * WIN_STEP * world), True, SEED, WIN_LEN, WIN_STEP, MIN_BPP, LN_CUTOFF,
                          rank=rank, world=world, dist=dist, piece_windows=piece, min_score=args.min_score)   # warm-up: one piece per strand
    sampler = ClockSampler(local)
    sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    from rmdetect_b200.hittable import HitTable
    table = HitTable(eng)                                         # every rank's hits stay on its GPU, as rows of a table
    r = scan_synthetic_genome(eng, total, True, SEED, WIN_LEN, WIN_STEP, MIN_BPP, LN_CUTOFF, rank=rank, world=world,
                              dist=dist, piece_windows=piece, min_score=args.min_score, table=table)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # rmfilter's work on the device-resident list (lib/candidates.py:99-207): order, overlaps, best hits per model
    rows0 = len(table)
    t1 = time.perf_counter()
    table.sort()
    t_sort = time.perf_counter() - t1
    if table.first_unsorted() != -1:
        raise SystemExit("bench.py: the hit table is not in Candidate order after rmd_table_sort")
    t1 = time.perf_counter()
    rows1 = table.nooverlap()
    t_noov = time.perf_counter() - t1
    if table.nooverlap() != rows1:                                # no two survivors are duplicates: a second pass removes nothing
        raise SystemExit("bench.py: rmd_table_nooverlap is not idempotent")
    t1 = time.perf_counter()
    rows2 = table.top_model(1000)
    t_top = time.perf_counter() - t1
    best = table.rows(0, min(rows2, 4))
    tt = torch.tensor([rows0, rows1, rows2], dtype=torch.int64, device="cuda")
    ts = torch.tensor([t_sort, t_noov, t_top], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    table_line = {"rows": int(tt[0]), "sort_s": float(ts[0]), "rows_after_nooverlap": int(tt[1]), "nooverlap_s": float(ts[1]),
                  "rows_after_top_model_1000": int(tt[2]), "top_model_s": float(ts[2]),
                  "best_scores_rank0": [float(x) for x in best["score"]],
                  "note": "hits with score > min_score kept as 64-byte rows on each rank's GPU (rmd_table_*): sorted in the "
                          "reference's Candidate order (checked row against row on the device), nooverlap replayed (a second pass removes nothing), "
                          "1000 best per model kept; seconds are the max over ranks"}
    t = torch.tensor([wall, r["rank_ms_generate"], r["rank_ms_scan_kernel"], r["rank_ms_device"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, ms_gen, ms_kernel, ms_dev = t.tolist()
    clocks = sampler.summary()
    if rank == 0:
        units = 2 * total * len(models)
        line = {
            "metric": "nt·models scanned/sec", "unit": "nt·model/s", "n_gpus": world, "steps": 1, "warmup": 1,
            "value": units / (ms_dev * 1e-3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: synthetic genome of %.3f G letters, both strands, win 150/75, models CL,TGA,GB,KT, "
                                   "synthetic bpp (gate on), pieces of %d windows + halo per GPU" % (args.gnt, piece),
                       "windows": r["windows"], "bpp_entries": r["bpp_entries"], "placements": r["tests_all"],
                       "gate_passes": r["tests_pairs"], "hits_before_dedupe": r["hits_before_dedupe"], "hits": r["hits"],
                       "min_score": args.min_score,
                       "hit_checksum": "%016x" % r["checksum"], "letter_counts_forward": r["counts"],
                       "timing": "value = letters x models / device time of the scan calls (CUDA events, max over ranks); "
                                 "wall_s also holds the on-device generation of sequence and bpp matrices (ViennaRNA's "
                                 "place), which is not part of the metric",
                       "wall_s": wall, "scan_device_s": ms_dev * 1e-3, "scan_kernel_s": ms_kernel * 1e-3,
                       "generation_s": ms_gen * 1e-3,
                       "value_incl_generation": units / wall,
                       "hit_table": table_line},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mb", type=int, default=10, help="chromosome length per GPU in Mb")
    ap.add_argument("--workload", default="chromosome", choices=["chromosome", "genome"],
                    help="chromosome: BASELINE configs[2] (the default, the metric's configuration); genome: configs[4]")
    ap.add_argument("--gnt", type=float, default=3.0, help="genome workload: letters per strand, in units of 1e9")
    ap.add_argument("--min-score", type=float, default=5.0,
                    help="genome workload: the command line's score filter (rmdetect.py --min-score), applied on the device")
    ap.add_argument("--piece-mb", type=float, default=20.0, help="genome workload: letters per resident piece, in Mb")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "genome":
        run_genome(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
