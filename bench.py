#!/usr/bin/env python
"""bench.py -- the hot path of tidk on B200, measured on the configuration BASELINE.json's target
sentence is written on: config 4, `tidk search --string TTAGGG --window 10000` on the synthetic
human-scale 3.1 Gb / 24-chromosome genome with N gaps.

  N = 1   the whole genome resident on one GPU.
  N > 1   (torchrun, one rank per GPU) the SAME genome cut into N contiguous window ranges
          (sharding.plan_shards: whole chromosomes where possible, long ones split at multiples of
          W, so no halo and no data-path collective): "scaling": "strong".  Every rank generates
          only the chromosomes its range touches.

A "step" is P passes of the window counter over the resident genome (P >= 32, chosen before the warm-up so that a step
lasts about 15 ms on the slowest rank -- 32 at N = 1, a few hundred on an eighth of the genome -- and identical on all
ranks; TIDK_BENCH_PASSES fixes it), issued through the
asynchronous entry point (tidk_cuda_window_counts_resident_async: the counts of pass i leave the device
while pass i + 1 counts) and collected with tidk_cuda_wait at the end of the step, so the counts of every
pass are in host memory when its step ends.  Passes are what makes the timed region long enough (>= 0.3 s
at N = 1) for a wall-clock number; `value` counts every pass.

  value     Gbases/s, genome resident in HBM, wall clock between barriers, max over ranks
  roofline  algorithmic bytes (1 B per scanned base) / device time of the counting kernel (CUDA events on
            the library's stream, summed by tidk_cuda_wait) vs the measured HBM peak
  e2e       Gbases/s through the C ABI with HOST (pinned) buffers: H2D of the (shard of the) genome,
            kernel, D2H of the counts inside the timed region, every step
  e2e_fasta FASTA file (page cache warm) -> TSV file through the C++ `tidk` driver in a fresh process:
            parse, CUDA context, upload, count, format, write (N = 1: one device; N > 1: rank 0 runs
            `--gpus N` in one process while the other ranks wait)
  roofline_explore  the other half of the path (explore) on config-5-shaped reads
            (100 000 HiFi-like reads, 1.5 Gb, --distance 0.5, k = 5..12), N = 1 only
  cpu_baseline  the CPU oracle (a port of the reference's algorithm; the Rust reference cannot be
            built in this image) on the same genome, all host cores, window ranges over threads

`--impl reference` times the CPU oracle instead (rank 0 only), all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WINDOW = 10000
QUERY = [b"TTAGGG"]
PASSES = int(os.environ.get("TIDK_BENCH_PASSES", "0"))  # 0: calibrated (>= 32, a step of about 15 ms)
STEP_MS_TARGET = 15.0
METRIC = "search_window_count_throughput"
UNIT = "Gbases/s"
SCALE = float(os.environ.get("TIDK_BENCH_SCALE", "1.0"))


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


# ------------------------------------------------------------------------------- workload
def c4_offsets():
    from telomeric_identifier_b200 import synth
    lens = synth.config4_lengths(SCALE)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    return off


def c4_shard(rank: int, world: int, threads: int):
    """This rank's contiguous window range of the C4 genome as a local batch (one sub-record per
    piece of a chromosome).  Returns (plan pieces, local sequence, local offsets, global offsets)."""
    from telomeric_identifier_b200 import sharding, synth
    off = c4_offsets()
    pieces = sharding.plan_shards(off, WINDOW, world)[rank]
    recs = sorted({p[0] for p in pieces})
    with ThreadPoolExecutor(max(1, min(threads, len(recs) or 1))) as ex:  # Philox streams keyed by chromosome: any subset, any order
        chrom = dict(zip(recs, ex.map(lambda i: synth.config4_contig(i, SCALE)[1], recs)))
    lens = []
    for rec, first, n in pieces:
        lo = first * WINDOW
        hi = min((first + n) * WINDOW, chrom[rec].size)
        lens.append(hi - lo)
    loff = np.zeros(len(pieces) + 1, dtype=np.uint64)
    if lens:
        loff[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    return pieces, chrom, loff, off


def fill_shard(dst: np.ndarray, pieces, chrom, loff):
    for i, (rec, first, n) in enumerate(pieces):
        lo = first * WINDOW
        a, b = int(loff[i]), int(loff[i + 1])
        dst[a:b] = chrom[rec][lo:lo + (b - a)]


def split_for_threads(off: np.ndarray, threads: int):
    """Window ranges as sub-records so that the record-sharded oracle keeps every core busy (windows are
    independent, search.rs:98: cutting a record at a multiple of W changes no count)."""
    from telomeric_identifier_b200 import sharding
    plans = sharding.plan_shards(off, WINDOW, max(1, threads * 4))
    sub = [0]
    for pieces in plans:
        for rec, first, n in pieces:
            o, e = int(off[rec]), int(off[rec + 1])
            sub.append(min(o + (first + n) * WINDOW, e))
    return np.asarray(sub, dtype=np.uint64)


def cpu_counts(seq, off, threads):
    from oracle import oracle as O
    sub = split_for_threads(off, threads) if threads > 1 else off
    return O.window_counts(seq, sub, WINDOW, QUERY, True, threads=threads)


def write_fasta(path, ids, seq, off):
    with open(path, "wb") as fh:
        for i, rid in enumerate(ids):
            s = seq[int(off[i]):int(off[i + 1])]
            fh.write(f">{rid}\n".encode())
            full = (s.size // 60) * 60
            body = np.empty((full // 60, 61), dtype=np.uint8)
            body[:, :60] = s[:full].reshape(-1, 60)
            body[:, 60] = 10
            fh.write(body.tobytes())
            if full < s.size:
                fh.write(s[full:].tobytes() + b"\n")


def reference_arm(args, emit, config, steps, warmup):
    """The CPU oracle on a bounded sample of the same genome (three whole chromosomes), every host core."""
    from telomeric_identifier_b200 import synth
    ncpu = os.cpu_count() or 1
    recs = [19, 20, 21]
    with ThreadPoolExecutor(3) as ex:
        parts = list(ex.map(lambda i: synth.config4_contig(i, SCALE)[1], recs))
    off = np.zeros(len(parts) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in parts])
    seq = np.concatenate(parts)
    for _ in range(warmup):
        cpu_counts(seq, off, ncpu)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_counts(seq, off, ncpu)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    cpu_counts(seq, off, 1)
    dt1 = time.perf_counter() - t1
    value = seq.size * steps / dt / 1e9
    cb = {"value": value, "unit": UNIT, "cores": ncpu, "kind": "port",
          "sample": f"chromosomes 20, 21, 22 of the C4 genome ({seq.size} bases), one pass per step, window ranges sharded "
                    f"over {ncpu} threads; oracle/tidk_oracle.c restates tidk 0.2.7 (the Rust reference cannot be built here)",
          "single_thread_as_the_reference_runs": {"value": seq.size / dt1 / 1e9, "unit": UNIT, "cores": 1,
                                                  "note": "search.rs:57 is one sequential loop; one pass over the same sample"}}
    emit({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
          "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
          "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
          "cpu_baseline": cb,
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def explore_block(ctx, peak):
    """explore on config-5-shaped reads: 100 000 reads ~N(15 kb, 2 kb), --distance 0.5, k = 5..12, threshold 100."""
    from telomeric_identifier_b200 import synth
    from oracle import oracle as O
    n_reads = max(1000, int(100_000 * SCALE))
    ids, seq, off = synth.config5(n_reads)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    scanned = int((2 * ((lens + 1) // 2)).sum())  # 2 * ceil(len / 2)
    pseq = ctx.pinned_empty(seq.size)
    pseq[:] = seq
    batch = ctx.upload(pseq, off)
    p_ready, p_ms, p_bytes = batch.planes_info()
    out = {"workload": f"C5-shaped: {n_reads} reads, {seq.size} bases, explore --minimum 5 --maximum 12 --threshold 100 --distance 0.5",
           "scanned_bases": scanned, "k_range": [5, 12], "threshold": 100,
           "kernels": "scan = explore_slices_kernel + prefix sums + vt_items_kernel + explore_vertical_kernel<5,12> on the resident "
                      "bit-planes (tiles with anything but upper-case A C G T: explore_planes_kernel on the ASCII copy); post = radix "
                      "sort of the event keys + runs_from_keys_kernel + compaction (CUB)",
           "k1_planes_build_ms": p_ms, "k1_frac_of_hbm_peak": 1.375 * seq.size / (p_ms * 1e-3) / 1e9 / peak if p_ms > 0 else None}
    runs = None
    for name, kmin, kmax in (("k5_12", 5, 12), ("k6_single", 6, 6)):
        best = None
        for _ in range(7):
            t0 = time.perf_counter()
            r = ctx.explore_runs_resident(batch, 0.5, kmin, kmax, 100)
            wall = time.perf_counter() - t0
            sm, pm = ctx.explore_timings()
            if best is None or sm < best[0]:
                best = (sm, pm, wall)
        if name == "k5_12":
            runs = r
        ach = scanned / (best[0] * 1e-3) / 1e9
        all_ms = best[0] + best[1]
        out[name] = {"scan_kernel_ms": best[0], "post_ms": best[1], "resident_call_wall_ms": best[2] * 1e3, "runs": int(r.size),
                     "scan_gbases_per_s": ach,
                     "roofline": {"bound": "instruction fetch / alu pipe, not hbm (0.26 B per base cross HBM): see DESIGN.md 3.3",
                                  "frac_of_alu_pipe_ncu": 0.51 if name == "k5_12" else 0.58, "issue_slots_busy_ncu": 0.39 if name == "k5_12" else 0.47,
                                  "ncu": "profiles/r02_ncu_full_prof_r02.json (top stall no_instruction 2.5 per issue for k = 5..12)",
                                  "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_launch": scanned},
                     "all_kernels": {"ms": all_ms, "achieved": scanned / (all_ms * 1e-3) / 1e9, "frac": scanned / (all_ms * 1e-3) / 1e9 / peak,
                                     "with_k1": {"ms": all_ms + p_ms, "frac": scanned / ((all_ms + p_ms) * 1e-3) / 1e9 / peak}}}
    # the same call through the host-pointer entry point (upload inside)
    e2e_ms = None
    for _ in range(3):
        t0 = time.perf_counter()
        r2 = ctx.explore_runs(pseq, off, 0.5, 5, 12, 100)
        w = (time.perf_counter() - t0) * 1e3
        e2e_ms = w if e2e_ms is None else min(e2e_ms, w)
    h2d, _, packed = ctx.last_transfer()
    out["e2e_host_pointer"] = {"ms_per_call": e2e_ms, "value": seq.size / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                               "h2d_bytes": h2d, "host_packed": packed, "runs_equal_resident": bool(np.array_equal(r2, runs))}
    # CPU port beside it: rayon's one-worker-per-record sharding (explore.rs:97-100) as one thread per core, on a
    # bounded sample (the first tenth of the reads), full run lists compared field by field
    nsub = max(100, n_reads // 10)
    sub_off = off[:nsub + 1]
    ncpu = os.cpu_count() or 1
    t0 = time.perf_counter()
    cpu_runs, _ = O.explore_runs(seq[:int(sub_off[-1])], sub_off, 0.5, 5, 12, 100, False, ncpu)
    cdt = time.perf_counter() - t0
    mine = runs[runs["record"] < nsub]
    mine = mine[np.lexsort((mine["start"], mine["slice"], mine["record"], mine["k"]))]
    same = mine.size == cpu_runs.size and all(np.array_equal(mine[f], cpu_runs[f]) for f in ("record", "slice", "k", "start", "end"))
    out["cpu_baseline"] = {"value": int(2 * ((lens[:nsub] + 1) // 2).sum()) / cdt / 1e9, "unit": UNIT, "cores": ncpu, "kind": "port",
                           "sample": f"first {nsub} reads, one pass, records sharded over threads",
                           "run_lists_equal_to_gpu": bool(same)}
    batch.free()
    return out


def fasta_e2e(pieces_all, world, local):
    """FASTA file -> TSV file through the C++ driver in a fresh process (context creation included)."""
    from telomeric_identifier_b200 import synth
    tidk = os.path.join(ROOT, "telomeric_identifier_b200", "bin", "tidk")
    if not os.path.exists(tidk):
        return {"unavailable": "bin/tidk not built"}
    d = tempfile.mkdtemp(prefix="tidk_bench_")
    fa = os.path.join(d, "c4.fa")
    ids, seq, off = pieces_all
    t0 = time.perf_counter()
    write_fasta(fa, ids, seq, off)
    wrote_s = time.perf_counter() - t0
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([tidk, "search", fa, "-s", "TTAGGG", "-w", str(WINDOW), "-o", "c4", "-d", d, "--gpu-stats",
                            "--device", str(local), "--gpus", str(world)], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        stats = [l for l in r.stderr.splitlines() if l.startswith("{")]
        if r.returncode != 0:
            return {"unavailable": f"tidk search failed rc={r.returncode}: {r.stderr[-300:]}"}
        if best is None or dt < best[0]:
            best = (dt, json.loads(stats[-1]) if stats else {})
    tsv = os.path.join(d, "c4_telomeric_repeat_windows.tsv")
    out = {"value": int(seq.size) / best[0] / 1e9, "unit": UNIT, "wall_s": best[0], "fasta_bytes": os.path.getsize(fa),
           "tsv_bytes": os.path.getsize(tsv), "devices_offered": world, "devices_used": best[1].get("devices_used", world),
           "what": "fresh `tidk search --gpus N` process: FASTA parse (mmap, multi-threaded) + CUDA context creation + upload + count + format + "
                   "write; page cache warm; best of 2.  --gpus N means up to N devices: the driver takes one device per 4 Gb of work (a "
                   "context costs ~0.25 s to create and as much to tear down; TIDK_GPUS_EXACT=1 forces N)", "breakdown": best[1], "fasta_write_s_untimed": wrote_s}
    for f in (fa, tsv):
        try:
            os.unlink(f)
        except OSError:
            pass
    try:
        os.rmdir(d)
    except OSError:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--skip-extras", action="store_true", help="no explore / FASTA / CPU legs (profiling runs)")
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
    ncpu = os.cpu_count() or 1

    config = {"workload": "C4: tidk search --string TTAGGG --window 10000, synthetic 3.1 Gb / 24 chromosomes with N gaps",
              "window": WINDOW, "queries": [q.decode() for q in QUERY],
              "passes_per_step": "calibrated per run: >= 32, a step of about 15 ms on the slowest rank (the line's passes_per_step)",
              "l2": "inputs (3.1 GB at N = 1, 386 MB per GPU at N = 8) larger than the 126 MB L2; no flush needed",
              "sharding": "contiguous window ranges of the one genome across ranks (long chromosomes cut at multiples of W), no collective"}
    if SCALE != 1.0:
        config["scale"] = SCALE

    if args.impl == "reference":
        if rank != 0:
            return
        reference_arm(args, emit, config, steps, warmup)
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

    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    pieces, chrom, off, goff = c4_shard(rank, world, max(1, ncpu // max(1, local_world)))
    bases = int(off[-1])
    ctx = T.Context(local)
    nwin = int(T.api.n_windows(off, WINDOW).sum()) * len(QUERY)
    # pinned host copies: what a host reader parsing straight into tidk_cuda_host_alloc memory holds
    pseq = ctx.pinned_empty(bases)
    fill_shard(pseq, pieces, chrom, off)
    chrom = None
    outs = []
    for _ in range(2):  # two sets of pinned output arrays, used alternately by the asynchronous calls
        po = ctx.pinned_empty(2 * nwin * 8).view(np.uint64)  # forward and reverse counts back to back: one D2H copy
        outs.append((po[:nwin], po[nwin:]))
    batch = ctx.upload(pseq, off)

    # correctness before timing anything: blocking call vs the oracle on the first 2 Mb of rank 0's range,
    # and asynchronous calls vs the blocking one on every rank
    f, r = ctx.window_counts_resident(batch, WINDOW, QUERY)
    f, r = f.copy(), r.copy()
    if rank == 0:
        from oracle import oracle as O
        nchk = min(int(off[1]), 2_000_000) // WINDOW * WINDOW
        of, orr = O.window_counts(pseq[:nchk], np.array([0, nchk], dtype=np.uint64), WINDOW, QUERY, False)
        assert (f[:of.size] == of).all() and (r[:orr.size] == orr).all(), "GPU counts differ from the oracle"
    for o in outs:
        o[0][:] = 0xFFFFFFFF
        ctx.window_counts_resident_async(batch, WINDOW, QUERY, o)
    ctx.wait()
    for o in outs:
        assert np.array_equal(o[0], f) and np.array_equal(o[1], r), "asynchronous counts differ from the blocking call"

    def run_passes(n):
        for p in range(n):
            ctx.window_counts_resident_async(batch, WINDOW, QUERY, outs[p & 1])
        return ctx.wait()

    # passes per step: enough for a step of about STEP_MS_TARGET on the slowest rank (the timed region of the default
    # run then lasts >= 0.3 s at every N), the same number on every rank
    passes = PASSES
    if passes <= 0:
        run_passes(8)
        barrier()
        t_c = time.perf_counter()
        run_passes(32)
        per_pass_ms = (time.perf_counter() - t_c) / 32 * 1e3
        if dist is not None:
            tc = torch.tensor([per_pass_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            per_pass_ms = float(tc.item())
        passes = int(min(2048, max(32, -(-STEP_MS_TARGET // per_pass_ms))))

    def step():
        return run_passes(passes)

    sampler = ClockSampler(local)
    sampler.start()
    t_s = time.perf_counter()
    for _ in range(warmup):
        step()
    kernel_ms, launches = 0.0, 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ms, nl = step()
        kernel_ms += ms
        launches += nl
        if os.environ.get("TIDK_BENCH_VERBOSE"):
            print(f"step kernel ms per pass {ms / passes:.4f}", file=sys.stderr)
    barrier()
    dt = time.perf_counter() - t0
    # nvidia-smi samples every 100 ms: keep the same step running (untimed) until the sampler has
    # seen at least ~1.5 s of this load, so the clocks line describes the kernel that was timed
    while time.perf_counter() - t_s < 1.5:
        step()
    clocks = sampler.stop()
    clocks["note"] = "sampled over warm-up + timed steps + identical untimed steps (>= 1.5 s of the same kernel)"
    assert np.array_equal(outs[0][0], f) and np.array_equal(outs[1][1], r), "counts changed during the timed steps"

    # blocking calls, for the per-call fixed cost (wall per call vs kernel time)
    barrier()
    t_b = time.perf_counter()
    kb = 0.0
    nb = 8
    for _ in range(nb):
        ctx.window_counts_resident(batch, WINDOW, QUERY, out=outs[0])
        kb += ctx.last_kernel_ms
    blocking = {"wall_ms_per_call": (time.perf_counter() - t_b) / nb * 1e3, "kernel_ms_per_call": kb / nb}
    # the same passes with the ASCII copy as the source (1 B per base from HBM, planes formed in registers): the kernel
    # round 1 reported, kept beside the default -- the bit-planes K1 builds once at upload (0.375 B per base)
    planes_ready, planes_ms, planes_bytes = batch.planes_info()
    os.environ["TIDK_PLANES_COUNT"] = "0"
    try:
        ctx.window_counts_resident(batch, WINDOW, QUERY, out=outs[0])
        ka = []
        for _ in range(8):
            ctx.window_counts_resident(batch, WINDOW, QUERY, out=outs[0])
            ka.append(ctx.last_kernel_ms)
        ascii_ms = float(np.median(ka))
        assert np.array_equal(outs[0][0], f) and np.array_equal(outs[0][1], r), "ASCII-source counts differ"
    finally:
        os.environ.pop("TIDK_PLANES_COUNT", None)

    # end to end through the C ABI with host buffers
    e2e_steps = args.e2e_steps or min(steps, 5)
    ctx.window_counts(pseq, off, WINDOW, QUERY)
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        ef, er = ctx.window_counts(pseq, off, WINDOW, QUERY)
    barrier()
    dt_e2e = time.perf_counter() - t1
    assert np.array_equal(ef, f) and np.array_equal(er, r), "host-pointer counts differ from the resident ones"
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
    del dst, src

    if dist is not None:
        t = torch.tensor([dt, dt_e2e, kernel_ms, ascii_ms, planes_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, kernel_ms, ascii_ms, planes_ms = (float(x) for x in t.tolist())
        tot = torch.tensor([float(bases), float(launches), float(f.sum()), float(r.sum()), float(h2d_bytes), float(d2h_bytes)],
                           device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_bases, launches = float(tot[0]), int(tot[1])
        fwd_total, rev_total, h2d_bytes, d2h_bytes = int(tot[2]), int(tot[3]), int(tot[4]), int(tot[5])
    else:
        total_bases = float(bases)
        fwd_total, rev_total = int(f.sum()), int(r.sum())

    peak, peak_src = measured_peak()
    extras = {}
    if not args.skip_extras:
        if world == 1:
            extras["roofline_explore"] = explore_block(ctx, peak)
            # explore --distance 0.01 on the resident genome (48 slices, 1 % of each chromosome end: launch-bound)
            lens = (off[1:] - off[:-1]).astype(np.float64)
            scanned = int(sum(2 * int(np.ceil(l * 0.01)) for l in lens.tolist()))
            best = None
            for _ in range(5):
                t2 = time.perf_counter()
                runs = ctx.explore_runs_resident(batch, 0.01, 5, 12, 100)
                wall = time.perf_counter() - t2
                sm, pm = ctx.explore_timings()
                if best is None or sm < best[0]:
                    best = (sm, pm, wall)
            extras["explore_c4_distance_0.01"] = {"scanned_bases": scanned, "scan_kernel_ms": best[0], "post_ms": best[1],
                                                  "resident_call_wall_ms": best[2] * 1e3, "runs": int(runs.size)}
        # FASTA in -> TSV out, fresh process (rank 0 drives `--gpus world`; the other ranks idle at the barrier)
        barrier()
        if rank == 0:
            try:
                if world == 1:
                    from telomeric_identifier_b200 import synth
                    ids = [f"chr{n}" for n in synth._HG38_NAMES]
                    extras["e2e_fasta"] = fasta_e2e((ids, pseq, goff), world, local)
                else:
                    from telomeric_identifier_b200 import synth
                    with ThreadPoolExecutor(max(1, ncpu // 2)) as ex:
                        parts = list(ex.map(lambda i: synth.config4_contig(i, SCALE)[1], range(len(goff) - 1)))
                    ids = [f"chr{n}" for n in synth._HG38_NAMES]
                    extras["e2e_fasta"] = fasta_e2e((ids, np.concatenate(parts), goff), world, local)
                    del parts
            except Exception as exc:  # never lose the bench line to the extra leg
                extras["e2e_fasta"] = {"unavailable": repr(exc)[:300]}
        barrier()

    if rank == 0:
        k_ms = kernel_ms / (steps * passes)  # max over ranks of the mean device time per pass
        per_gpu = total_bases / world
        achieved = per_gpu / (k_ms * 1e-3) / 1e9  # GB/s per GPU at 1 B per scanned base
        flat = planes_ready and os.environ.get("TIDK_FLAT", "1") != "0"
        kname = ("window_counts_flat_kernel<StaticMatcher<6,TTAGGG>,3>" if flat else
                 "window_counts_kernel<StaticMatcher<6,TTAGGG>,4,%s>" % ("PackedSource" if planes_ready else "AsciiSource"))
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kname, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": per_gpu, "peak_source": peak_src,
                "note": "per GPU: bases of one rank's range / slowest rank's mean kernel time per pass; `achieved` counts the "
                        "ALGORITHMIC 1 B per scanned base (SURVEY.md 8d)"}
        if planes_ready:
            # the resident batch is read as bit-planes (lo / hi: 0.25 B per base; the validity plane, 0.125 B per base, only in
            # the 4 096-base steps that hold anything but A C G T) that K1 built once behind the upload: frac above 1 means
            # less than one byte per base crosses HBM, not that the peak was beaten.  The estimate is the upper bound (all
            # three planes); `traffic` is what ncu measured.
            dram = 0.375 * per_gpu
            roof["source"] = {"what": "resident bit-planes, built once per upload by planes_build_kernel (K1); flat form (window_counts_flat.cuh): "
                                      "a warp walks 32 Ki consecutive bases in 4 096-base steps, windows are an accounting matter",
                              "dram_bytes_per_launch_upper_bound": dram, "k1_planes_build_ms": planes_ms,
                              "k1_bytes": "1 B read + 0.375 B written per base",
                              "k1_frac_of_peak": 1.375 * per_gpu / (planes_ms * 1e-3) / 1e9 / peak if planes_ms > 0 else None,
                              "planes_device_bytes": planes_bytes,
                              "limiter": "alu pipe (ncu: 85 % busy, issue slots 69 %), not HBM: see frac_of_peak_on_measured_traffic"}
            roof["ascii_source"] = {"kernel": "window_counts_kernel<StaticMatcher<6,TTAGGG>,4,AsciiSource>", "kernel_ms": ascii_ms,
                                    "achieved": per_gpu / (ascii_ms * 1e-3) / 1e9, "frac": per_gpu / (ascii_ms * 1e-3) / 1e9 / peak,
                                    "note": "same passes with TIDK_PLANES_COUNT=0: 1 B per base from HBM, planes formed in registers "
                                            "(blocking calls, median of 8)"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr) and world == 1 and SCALE == 1.0:
            try:
                roof["traffic"] = json.load(open(tr)).get("window_counts_kernel_c4_dram_bytes_per_launch")
                if roof["traffic"] and planes_ready:
                    roof["source"]["frac_of_peak_on_measured_traffic"] = roof["traffic"] / (k_ms * 1e-3) / 1e9 / peak
            except Exception:
                pass
        cb = None
        if world == 1 and not args.skip_extras:  # the CPU baseline is reported at N = 1 only
            t3 = time.perf_counter()
            cf, cr = cpu_counts(pseq, off, ncpu)
            cdt = time.perf_counter() - t3
            cb = {"value": bases / cdt / 1e9, "unit": UNIT, "cores": ncpu, "kind": "port",
                  "sample": f"the whole genome ({bases} bases), one pass, window ranges sharded over {ncpu} threads; oracle/tidk_oracle.c "
                            "restates tidk 0.2.7 (the Rust reference cannot be built here; the reference itself is single-threaded, search.rs:57)",
                  "all_windows_equal_to_gpu": bool(np.array_equal(cf, f) and np.array_equal(cr, r))}
        e2e_val = total_bases * e2e_steps / dt_e2e / 1e9
        line = {"metric": METRIC, "value": total_bases * passes * steps / dt / 1e9, "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "passes_per_step": passes, "timed_region_s": dt,
                "kernel_ms_per_pass": k_ms, "wall_ms_per_pass": dt / steps / passes * 1e3,
                "wall_over_kernel": dt / steps / passes * 1e3 / k_ms,
                "kernel_only_value": total_bases / (k_ms * 1e-3) / 1e9,
                "counts": {"fwd_total": fwd_total, "rev_total": rev_total}, "blocking_call": blocking, "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT,
                        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                        "host_bytes_per_step": int(total_bases),
                        "upload": ("two-ended: ASCII chunks by DMA from the front, host-packed lo/hi bit-planes (2 bits per base + validity "
                                   "exceptions) from the back, packed by %d host threads per rank inside the timed region"
                                   % min(32, ncpu // max(1, local_world))
                                   if host_packed else "ASCII, straight DMA from the caller's pinned buffer"),
                        "ms_per_step": dt_e2e / e2e_steps * 1e3,
                        "h2d_bound_gbs_rank0": h2d_gbs,
                        "frac_of_h2d_bound_rank0": (bases * e2e_steps / dt_e2e / 1e9) / h2d_gbs},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cb}
        line.update(extras)
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
