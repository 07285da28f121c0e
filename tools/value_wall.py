"""Wall time per resident window-count call on C2 with the counts read back into pinned host arrays (the `value`
of bench.py) next to the kernel time.  TIDK_DIRECT_OUT=0 keeps the device-to-host copy behind the kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

ids, seq, off = synth.config2(scale=float(os.environ.get("TIDK_BENCH_SCALE", "1.0")))
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    nwin = int(T.api.n_windows(off, 10000).sum())
    pout = ctx.pinned_empty(2 * nwin * 8).view(np.uint64)
    pf, pr = pout[:nwin], pout[nwin:]
    ref_f, ref_r = ctx.window_counts_resident(b, 10000, [b"AACCT"])          # pageable arrays: always the copy
    for _ in range(10):
        ctx.window_counts_resident(b, 10000, [b"AACCT"], out=(pf, pr))
    assert (pf == ref_f).all() and (pr == ref_r).all()
    n, kms = 200, 0.0
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.window_counts_resident(b, 10000, [b"AACCT"], out=(pf, pr))
        kms += ctx.last_kernel_ms
    dt = time.perf_counter() - t0
    assert (pf == ref_f).all() and (pr == ref_r).all()
    print(f"direct_out={os.environ.get('TIDK_DIRECT_OUT', '1')} wall/call {dt / n * 1e3:.4f} ms = {seq.size / (dt / n) / 1e12:.2f} Tbases/s; "
          f"kernel {kms / n:.4f} ms = {seq.size / (kms / n * 1e-3) / 1e12:.2f} Tbases/s")
