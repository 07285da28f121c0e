"""Kernel time vs window size at fixed input size: smaller items shorten the drain at the end of the queue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
rng = np.random.default_rng(1)
n = 500_000_000
seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]
off = np.array([0, n], dtype=np.uint64)
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    for W in (1000, 2500, 5000, 10000, 15000, 20000, 30000, 32768, 65536, 1000000):
        ms = []
        for _ in range(8):
            ctx.window_counts_resident(b, W, [b"AACCT"]); ms.append(ctx.last_kernel_ms)
        print(f"W={W:8d}  best {min(ms[2:])*1e3:8.1f} us  median {float(np.median(ms[2:]))*1e3:8.1f} us")
