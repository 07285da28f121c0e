"""Explore scan rate on long slices (C2 genome, d = 0.5): mostly interior blocks, unlike read-sized slices."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
ids, seq, off = synth.config2(scale=1.0)
with T.Context(0) as ctx:
  for label, sq in (("soft-masked 2%", seq), ("upper-cased", np.where((seq >= 97) & (seq <= 122), seq - 32, seq).astype(np.uint8))):
    b = ctx.upload(sq, off)
    print(label)
    for d in (0.5, 0.1):
        for _ in range(3):
            runs = ctx.explore_runs_resident(b, d, 5, 12, 100)
            sm, pm = ctx.explore_timings()
            scanned = sum(2 * int(np.ceil(float(l) * d)) for l in (off[1:] - off[:-1]).tolist())
        print("  " + f"d={d} scanned={scanned} scan_ms={sm:.3f} post_ms={pm:.3f} scan GB/s={scanned/(sm*1e-3)/1e9:.1f} runs={runs.size}")
