"""Explore on the C4-shaped genome (N caps, N gaps): how much do the byte-wise (non-ACGT) blocks cost?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
ids, seq, off = synth.config4(scale=float(os.environ.get("TIDK_BENCH_SCALE", "0.3")))
noN = seq.copy(); noN[noN == ord("N")] = ord("A")
with T.Context(0) as ctx:
    for label, sq in (("with N gaps (%.1f%% N)" % (100 * float((seq == 78).mean())), seq), ("N replaced by A", noN)):
        b = ctx.upload(sq, off)
        for d in (0.5, 0.01):
            for _ in range(3):
                runs = ctx.explore_runs_resident(b, d, 5, 12, 100)
                sm, pm = ctx.explore_timings()
            scanned = sum(2 * int(np.ceil(float(l) * d)) for l in (off[1:] - off[:-1]).tolist())
            print(f"{label:28s} d={d} scanned={scanned} scan_ms={sm:.3f} post_ms={pm:.3f} scan GB/s={scanned/(sm*1e-3)/1e9:.1f} runs={runs.size}")
        b.free()
