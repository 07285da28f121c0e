"""Vertical scan vs plane kernel over batch sizes (C5-shaped reads, d = 0.5) and on the C3 genome (d = 0.01): where the cross-over lies.
usage: python tools/explore_sweep.py   (TIDK_VERTICAL_MIN_TILES=0 forces the vertical scan, a huge value the plane kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

with T.Context(0) as ctx:
    for n_reads in (500, 2000, 5000, 10000, 20000, 50000):
        ids, seq, off = synth.config5(n_reads=n_reads)
        b = ctx.upload(seq, off)
        best = 1e9
        for _ in range(6):
            runs = ctx.explore_runs_resident(b, 0.5, 5, 12, 100)
            best = min(best, ctx.explore_timings()[0])
        print(f"C5 reads={n_reads} bases={seq.size} scan_ms={best:.4f} runs={runs.size}", flush=True)
        b.free()
    ids, seq, off = synth.config2(scale=1.0)
    b = ctx.upload(seq, off)
    for d in (0.01, 0.05, 0.2):
        best = 1e9
        for _ in range(6):
            runs = ctx.explore_runs_resident(b, d, 5, 12, 100)
            best = min(best, ctx.explore_timings()[0])
        print(f"C3 genome d={d} scan_ms={best:.4f} runs={runs.size}", flush=True)
