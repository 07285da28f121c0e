"""End-to-end timing of the host-pointer entry point on C2 (find AACCT, W = 10000): host buffer in, counts out,
upload inside the timed region.  Environment knobs of the library apply (TIDK_PACK_ISA=avx2, TIDK_PACK_THREADS=n)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

n = int(os.environ.get("TIDK_PROF_STEPS", "12"))
ids, seq, off = synth.config2(scale=float(os.environ.get("TIDK_BENCH_SCALE", "1.0")))
with T.Context(0) as ctx:
    pseq = ctx.pinned_empty(seq.size); pseq[:] = seq
    f0, r0 = ctx.window_counts(pseq, off, 10000, [b"AACCT"])
    ms = []
    for _ in range(n):
        t0 = time.perf_counter()
        f, r = ctx.window_counts(pseq, off, 10000, [b"AACCT"])
        ms.append((time.perf_counter() - t0) * 1e3)
        assert (f == f0).all() and (r == r0).all()
    h2d, d2h, packed = ctx.last_transfer()
    ms.sort()
    print(f"isa={os.environ.get('TIDK_PACK_ISA', 'auto')} threads={os.environ.get('TIDK_PACK_THREADS', 'auto')} "
          f"ms min {ms[0]:.2f} median {ms[len(ms) // 2]:.2f} -> {seq.size / ms[len(ms) // 2] / 1e6:.1f} Gbases/s (best {seq.size / ms[0] / 1e6:.1f}) "
          f"h2d_bytes={h2d} packed={packed} fwd={int(f.sum())}")
