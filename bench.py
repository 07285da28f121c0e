#!/usr/bin/env python
"""bench.py -- the hot path of tidk on B200, measured the way BASELINE.json names it.

Workload (N = 1): BASELINE config[1] -- `tidk find --clade Lepidoptera` (one repeat,
AACCT) with --window 10000 on a synthetic 500 Mb / 30-chromosome genome.
A "step" is one pass of the window counter over the whole genome.

  value     Gbases/s, genome already resident in HBM, K steps, wall clock between
            synchronisations, max over ranks (the per-step device time of the counting
            kernel, from CUDA events on the library's stream, feeds `roofline`)
  e2e       Gbases/s through the C ABI with HOST (pinned) buffers: H2D of the genome,
            kernel, D2H of the counts inside the timed region, every step
  roofline  algorithmic bytes (1 B per scanned base) / kernel time vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's algorithm; the Rust reference
            cannot be built in this image) on the same workload, 1 thread as in the reference

N > 1 (torchrun, one rank per GPU): every rank holds its own 30-chromosome genome
(seed + rank) -- records shard across ranks with no data-path collective ("weak").

`--impl reference` times the CPU oracle instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WINDOW = 10000
REPEATS = [b"AACCT"]  # clades/curated.csv: every Lepidoptera row is AACCT
METRIC = "find_window_count_throughput"
UNIT = "Gbases/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_workload(seed: int):
    from telomeric_identifier_b200 import synth
    scale = float(os.environ.get("TIDK_BENCH_SCALE", "1.0"))
    ids, seq, off = synth.config2(scale=scale, seed=seed)
    return ids, seq, off, scale


def cpu_leg(seq, off, steps, warmup, budget_s=150.0):
    """Time the CPU oracle (1 thread: search.rs:57 / finder.rs:85 are sequential loops)."""
    from oracle import oracle as O
    # bounded sample: whole chromosomes from the front until the whole run fits the budget
    probe_n = min(int(off[1]), seq.size)
    t0 = time.perf_counter()
    O.window_counts(seq[:min(probe_n, 4_000_000)], np.array([0, min(probe_n, 4_000_000)], dtype=np.uint64), WINDOW, REPEATS, False)
    rate = min(probe_n, 4_000_000) / max(time.perf_counter() - t0, 1e-6)
    max_bases = rate * budget_s / max(steps + warmup, 1)
    nrec = 1
    while nrec < len(off) - 1 and int(off[nrec + 1]) <= max_bases:
        nrec += 1
    sub_off = off[:nrec + 1].copy()
    sub = seq[:int(sub_off[-1])]
    for _ in range(warmup):
        O.window_counts(sub, sub_off, WINDOW, REPEATS, False)
    t0 = time.perf_counter()
    for _ in range(steps):
        f, r = O.window_counts(sub, sub_off, WINDOW, REPEATS, False)
    dt = time.perf_counter() - t0
    return dict(value=sub.size * steps / dt / 1e9, unit=UNIT, cores=1, kind="port",
                sample=f"first {nrec} of {len(off) - 1} chromosomes ({sub.size} bases) x {steps} passes; "
                       f"oracle/tidk_oracle.c restates tidk 0.2.7 (the Rust reference cannot be built here)"), dt / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its "NCCL version ..." banner to
    # stdout with printf when NCCL_DEBUG=VERSION, whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at
    # stderr for the whole run and the line goes out through a private duplicate of the original stdout.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    config = {"workload": "C2: tidk find --clade Lepidoptera (AACCT) --window 10000, synthetic 500 Mb / 30 chromosomes",
              "window": WINDOW, "queries": [q.decode() for q in REPEATS], "records_per_gpu": 30,
              "l2": "inputs (500 MB per GPU) larger than the 126 MB L2; no flush needed",
              "sharding": "records across ranks, no collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        ids, seq, off, scale = make_workload(2)
        if scale != 1.0:
            config["scale"] = scale
        cb, per_step = cpu_leg(seq, off, steps, warmup)
        # for information: the same port sharded by record over every host core -- the
        # reference itself never does this (search.rs:57 / finder.rs:85 are sequential loops)
        from oracle import oracle as O
        ncpu = os.cpu_count() or 1
        t0 = time.perf_counter()
        O.window_counts(seq, off, WINDOW, REPEATS, False, threads=ncpu)
        cb["all_cores_for_information"] = {"value": seq.size / (time.perf_counter() - t0) / 1e9, "unit": UNIT, "cores": ncpu,
                                           "sample": "whole genome, records sharded over threads, 1 pass"}
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import telomeric_identifier_b200 as T
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    ids, seq, off, scale = make_workload(2 + rank)
    if scale != 1.0:
        config["scale"] = scale
    bases = int(seq.size)
    ctx = T.Context(local)
    nwin = int(T.api.n_windows(off, WINDOW).sum()) * len(REPEATS)
    # pinned host copies: what a host reader parsing straight into tidk_cuda_host_alloc memory holds
    pseq = ctx.pinned_empty(bases); pseq[:] = seq
    pout = ctx.pinned_empty(2 * nwin * 8).view(np.uint64)  # forward and reverse counts back to back: one D2H copy
    pf, pr = pout[:nwin], pout[nwin:]
    batch = ctx.upload(pseq, off)

    # correctness spot check against the oracle before timing anything (rank 0, first 2 Mb)
    f, r = ctx.window_counts_resident(batch, WINDOW, REPEATS, out=(pf, pr))
    if rank == 0:
        from oracle import oracle as O
        nchk = min(int(off[1]), 2_000_000) // WINDOW * WINDOW
        of, orr = O.window_counts(seq[:nchk], np.array([0, nchk], dtype=np.uint64), WINDOW, REPEATS, False)
        assert (f[:of.size] == of).all() and (r[:orr.size] == orr).all(), "GPU counts differ from the oracle"

    sampler = ClockSampler(local)
    sampler.start()
    t_s = time.perf_counter()
    for _ in range(warmup):
        ctx.window_counts_resident(batch, WINDOW, REPEATS, out=(pf, pr))
    kernel_ms, launches = 0.0, 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.window_counts_resident(batch, WINDOW, REPEATS, out=(pf, pr))
        kernel_ms += ctx.last_kernel_ms
        launches += ctx.last_kernel_launches
    barrier()
    dt = time.perf_counter() - t0
    # nvidia-smi samples every 100 ms: keep the same step running (untimed) until the sampler has
    # seen at least ~1.5 s of this load, so the clocks line describes the kernel that was timed
    while time.perf_counter() - t_s < 1.5:
        ctx.window_counts_resident(batch, WINDOW, REPEATS, out=(pf, pr))
    clocks = sampler.stop()
    clocks["note"] = "sampled over warm-up + timed steps + identical untimed steps (>= 1.5 s of the same kernel)"

    # end to end through the C ABI with host buffers
    e2e_steps = args.e2e_steps or min(steps, 5)
    ctx.window_counts(pseq, off, WINDOW, REPEATS)
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.window_counts(pseq, off, WINDOW, REPEATS)
    barrier()
    dt_e2e = time.perf_counter() - t1
    h2d_bytes, d2h_bytes, host_packed = ctx.last_transfer()

    # the bound for e2e: a bare pinned host -> device copy of the same buffer (PCIe Gen5 x16)
    src = torch.from_numpy(pseq)
    dst = torch.empty(bases, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    ev[1].record()
    torch.cuda.synchronize()
    h2d_gbs = bases * 3 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
    del dst

    # the other half of the hot path (explore) on the same resident genome:
    # C3 (--distance 0.01, 60 slices) and --distance 0.5 (every base scanned once, 8 k's per pass)
    explore = None
    if world == 1:
        explore = {}
        lens = (off[1:] - off[:-1]).astype(np.float64)
        for name, d in (("c3_distance_0.01", 0.01), ("distance_0.5", 0.5)):
            scanned = int(sum(2 * int(np.ceil(l * d)) for l in lens.tolist()))
            best = None
            for _ in range(5):
                t2 = time.perf_counter()
                runs = ctx.explore_runs_resident(batch, d, 5, 12, 100)
                wall = time.perf_counter() - t2
                sm, pm = ctx.explore_timings()
                if best is None or sm < best[0]:
                    best = (sm, pm, wall)
            # the same call through the host-pointer entry point (pinned buffer): with d = 0.01 only the
            # slices are uploaded, with d = 0.5 the whole genome is
            e2e_ms = None
            for _ in range(3):
                t3 = time.perf_counter()
                ctx.explore_runs(pseq, off, d, 5, 12, 100)
                w3 = (time.perf_counter() - t3) * 1e3
                e2e_ms = w3 if e2e_ms is None else min(e2e_ms, w3)
            explore[name] = {"scanned_bases": scanned, "scan_kernel_ms": best[0], "post_ms": best[1],
                             "e2e_host_pointer_call_ms": e2e_ms,
                             "call_wall_ms": best[2] * 1e3, "runs": int(runs.size),
                             "scan_gbases_per_s": scanned / (best[0] * 1e-3) / 1e9,
                             "kernel": "explore_planes_kernel<5,12>", "k_range": [5, 12], "threshold": 100}
            # the CPU port beside it: rayon's one-worker-per-record sharding (explore.rs:97-100) as
            # min(host cores, records) threads, one pass over the same genome, same arguments
            from oracle import oracle as O
            cpu_threads = max(1, min(os.cpu_count() or 1, len(off) - 1))
            t4 = time.perf_counter()
            cpu_runs, _ = O.explore_runs(seq, off, d, 5, 12, 100, False, cpu_threads)
            cpu_dt = time.perf_counter() - t4
            explore[name]["cpu_baseline"] = {"value": scanned / cpu_dt / 1e9, "unit": UNIT, "cores": cpu_threads,
                                             "kind": "port", "sample": "whole genome, one pass, records sharded over threads",
                                             "runs_equal_to_gpu": bool(len(cpu_runs) == int(runs.size))}

    if dist is not None:
        t = torch.tensor([dt, dt_e2e, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, kernel_ms = (float(x) for x in t.tolist())
        tot = torch.tensor([float(bases), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_bases, launches = float(tot[0]), int(tot[1])
    else:
        total_bases = float(bases)

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = kernel_ms / steps
        achieved = bases / (k_ms * 1e-3) / 1e9  # GB/s per GPU at 1 B per scanned base
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "window_counts_kernel<StaticMatcher<5,AACCT>>", "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": bases, "peak_source": peak_src}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get("window_counts_kernel_dram_bytes_per_launch")
            except Exception:
                pass
        cb = None
        if world == 1:  # the CPU baseline is reported at N = 1 only
            cb, _ = cpu_leg(seq, off, 1, 0, budget_s=25.0)
        line = {"metric": METRIC, "value": total_bases * steps / dt / 1e9, "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "kernel_ms_per_step": k_ms, "clocks": clocks,
                "e2e": {"value": total_bases * e2e_steps / dt_e2e / 1e9, "unit": UNIT,
                        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                        "host_bytes_per_step": bases,
                        "upload": ("two-ended: ASCII chunks by DMA from the front, host-packed lo/hi bit-planes (2 bits per base + validity "
                                   "exceptions) from the back, packed by %d host threads inside the timed region"
                                   % min(32, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
                                   if host_packed else "ASCII, straight DMA from the caller's pinned buffer"),
                        "ms_per_step": dt_e2e / e2e_steps * 1e3,
                        "h2d_bound_gbs_rank0": h2d_gbs,
                        "frac_of_h2d_bound_rank0": (bases * e2e_steps / dt_e2e / 1e9) / h2d_gbs},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cb}
        if explore is not None:
            for v in explore.values():
                v["frac_of_hbm_peak"] = v["scan_gbases_per_s"] / peak
            line["explore"] = explore
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
