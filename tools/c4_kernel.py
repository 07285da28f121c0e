"""Kernel time of the window counter on the C4-shaped genome (TTAGGG, W=10000, resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
ids, seq, off = synth.config4(scale=float(os.environ.get("TIDK_BENCH_SCALE", "0.5")))
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    ms = []
    for _ in range(15):
        f, r = ctx.window_counts_resident(b, 10000, [b"TTAGGG"])
        ms.append(ctx.last_kernel_ms)
    ms = sorted(ms[3:])
    print(f"C4 x{os.environ.get('TIDK_BENCH_SCALE', '0.5')} bases={seq.size} kernel min={ms[0]:.4f} med={ms[len(ms)//2]:.4f} ms GB/s(med)={seq.size / (ms[len(ms)//2] * 1e-3) / 1e9:.0f} sum={int(f.sum())},{int(r.sum())}")
