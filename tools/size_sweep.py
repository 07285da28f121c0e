"""Kernel time vs input size (one GPU): separates the per-byte rate from the fixed cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
rng = np.random.default_rng(1)
base = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 1 << 28, dtype=np.uint8)]
with T.Context(0) as ctx:
    for mb in (16, 32, 64, 128, 256, 512, 1024, 2048):
        n = mb << 20
        seq = np.tile(base, (n + base.size - 1) // base.size)[:n]
        off = np.array([0, n], dtype=np.uint64)
        b = ctx.upload(seq, off)
        ms = []
        for _ in range(8):
            ctx.window_counts_resident(b, 10000, [b"AACCT"]); ms.append(ctx.last_kernel_ms)
        best, med = min(ms[2:]), float(np.median(ms[2:]))
        print(f"{mb:5d} MiB  best {best*1e3:8.1f} us  median {med*1e3:8.1f} us  {n/best/1e6:8.1f} GB/s(best)")
        b.free()
