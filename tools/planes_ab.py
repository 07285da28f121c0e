"""Resident planes (K1 at upload) vs the fused ASCII kernels: window counter on C4, explore on C5-shaped reads."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
import bench
from telomeric_identifier_b200 import synth
pieces, chrom, off, goff = bench.c4_shard(0, 1, 16)
ctx = T.Context(0)
pseq = ctx.pinned_empty(int(off[-1])); bench.fill_shard(pseq, pieces, chrom, off); chrom = None
batch = ctx.upload(pseq, off)
print("planes", batch.planes_info(), "bases", int(off[-1]))
ks = []
for i in range(12):
    f, r = ctx.window_counts_resident(batch, 10000, [b"TTAGGG"]); ks.append(round(ctx.last_kernel_ms, 4))
print("window counter C4", os.environ.get("TIDK_PLANES_COUNT", "planes"), ks, int(f.sum()), int(r.sum()))
for d in (0.5, 0.01):
    best = None
    for _ in range(5):
        runs = ctx.explore_runs_resident(batch, d, 5, 12, 100); sm, pm = ctx.explore_timings(); best = sm if best is None else min(best, sm)
    print("explore C4 genome d", d, "scan ms", round(best, 4), "runs", runs.size)
batch.free()
ids, seq, o5 = synth.config5(100_000)
b5 = ctx.upload(seq, o5)
print("planes C5", b5.planes_info(), "bases", seq.size)
for kmin, kmax in ((5, 12), (6, 6), (4, 15)):
    best = None
    for _ in range(7):
        runs = ctx.explore_runs_resident(b5, 0.5, kmin, kmax, 100); sm, pm = ctx.explore_timings(); best = sm if best is None else min(best, sm)
    print("explore C5 k", kmin, kmax, "scan ms", round(best, 4), "Gb/s", round(seq.size / best / 1e6, 1), "runs", runs.size, "sum", int(runs["end"].sum()), int(runs["start"].sum()))
