"""Small driver for ncu: C2 (find AACCT, W=10000) resident, a few kernel launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

scale = float(os.environ.get("TIDK_BENCH_SCALE", "1.0"))
n = int(os.environ.get("TIDK_PROF_STEPS", "5"))
ids, seq, off = synth.config2(scale=scale)
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    ms = []
    for _ in range(n):
        f, r = ctx.window_counts_resident(b, 10000, [b"AACCT"])
        ms.append(ctx.last_kernel_ms)
    print("kernel_ms", ms, "GB/s", seq.size / (min(ms) * 1e-3) / 1e9, "fwd", int(f.sum()), "rev", int(r.sum()))
