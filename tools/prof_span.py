"""Flat window counter: kernel time and wall time per asynchronous pass vs input size (strong-scaling shares of the C4 genome).
usage: TIDK_CUDA_LIB=... python tools/prof_span.py 0.125 0.25 1.0"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

with T.Context(0) as ctx:
    for sc in [float(x) for x in sys.argv[1:]] or [0.125]:
        ids, seq, off = synth.config4(scale=sc)
        pseq = ctx.pinned_empty(seq.size); pseq[:] = seq
        b = ctx.upload(pseq, off)
        nwin = int(T.api.n_windows(off, 10000).sum())
        outs = []
        for _ in range(2):
            po = ctx.pinned_empty(2 * nwin * 8).view(np.uint64)
            outs.append((po[:nwin], po[nwin:]))
        ks = []
        for _ in range(10):
            ctx.window_counts_resident(b, 10000, [b"TTAGGG"], out=outs[0]); ks.append(ctx.last_kernel_ms)
        for _ in range(2):
            for p in range(32): ctx.window_counts_resident_async(b, 10000, [b"TTAGGG"], outs[p & 1])
            ctx.wait()
        t0 = time.perf_counter(); km = 0.0
        for _ in range(5):
            for p in range(32): ctx.window_counts_resident_async(b, 10000, [b"TTAGGG"], outs[p & 1])
            km += ctx.wait()[0]
        dt = time.perf_counter() - t0
        print(f"scale={sc} bases={seq.size} blocking kernel_ms min={min(ks[2:]):.4f} | async kernel_ms/pass={km / 160:.4f} wall_ms/pass={dt / 160 * 1e3:.4f} "
              f"Tbases/s wall={seq.size * 160 / dt / 1e12:.2f}", flush=True)
        b.free()
