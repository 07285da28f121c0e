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
    for qs in ([b"AACCT"], [b"AACCT", b"TTAGG"], [b"AACCT", b"TTAGG", b"TTAGGG"], [b"AACCT", b"TTAGG", b"TTAGGG", b"AACCC"]):
        ms = []
        for _ in range(8):
            ctx.window_counts_resident(b, 10000, qs); ms.append(ctx.last_kernel_ms)
        print(f"{len(qs)} launches: best {min(ms[2:])*1e3:8.1f} us  median {float(np.median(ms[2:]))*1e3:8.1f} us  launches={ctx.last_kernel_launches}")
