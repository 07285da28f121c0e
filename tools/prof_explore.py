"""Driver for timing / ncu: explore on C5-like reads (d = 0.5) and C3 (500 Mb genome, d = 0.01)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

n_reads = int(os.environ.get("TIDK_PROF_READS", "100000"))
steps = int(os.environ.get("TIDK_PROF_STEPS", "3"))
which = os.environ.get("TIDK_PROF_CFG", "c5,c3").split(",")
KMIN, KMAX = int(os.environ.get("TIDK_PROF_KMIN", "5")), int(os.environ.get("TIDK_PROF_KMAX", "12"))
with T.Context(0) as ctx:
    if "c5" in which:
        ids, seq, off = synth.config5(n_reads=n_reads)
        b = ctx.upload(seq, off)
        for _ in range(steps):
            t0 = time.perf_counter()
            runs = ctx.explore_runs_resident(b, 0.5, KMIN, KMAX, 100)
            dt = time.perf_counter() - t0
            scanned = sum(2 * int(np.ceil(float(l) * 0.5)) for l in (off[1:] - off[:-1]).tolist())
            sm, pm = ctx.explore_timings()
            hp, hq = ctx.explore_host_timings()
            print(f"C5 reads={n_reads} scanned={scanned} scan_ms={sm:.3f} post_ms={pm:.3f} "
                  f"scan GB/s={scanned / (sm * 1e-3) / 1e9:.1f} wall_ms={dt * 1e3:.1f} host_prep_ms={hp:.2f} post_wall_ms={hq:.2f} runs={runs.size}")
        b.free()
    if "c3" in which:
        ids, seq, off = synth.config2(scale=float(os.environ.get("TIDK_BENCH_SCALE", "1.0")))
        b = ctx.upload(seq, off)
        for _ in range(steps):
            t0 = time.perf_counter()
            runs = ctx.explore_runs_resident(b, 0.01, KMIN, KMAX, 100)
            dt = time.perf_counter() - t0
            scanned = sum(2 * int(np.ceil(float(l) * 0.01)) for l in (off[1:] - off[:-1]).tolist())
            sm, pm = ctx.explore_timings()
            print(f"C3 scanned={scanned} scan_ms={sm:.3f} post_ms={pm:.3f} scan GB/s={scanned / (sm * 1e-3) / 1e9:.1f} "
                  f"wall_ms={dt * 1e3:.1f} runs={runs.size}")
