"""Upload of a pageable (plain numpy) sequence batch: the multi-threaded staged copy against one cudaMemcpy
(TIDK_STAGED_UPLOAD=0), with a window count on the uploaded batch as the content check."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
ids, seq, off = synth.config2(scale=float(os.environ.get("TIDK_BENCH_SCALE", "1.0")))
with T.Context(0) as ctx:
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); b = ctx.upload(seq, off); f, r = ctx.window_counts_resident(b, 10000, [b"AACCT"]); ts.append(time.perf_counter() - t0); b.free()
    print(f"staged={os.environ.get('TIDK_STAGED_UPLOAD', '1')} upload+count of {seq.size} pageable bytes: best {min(ts) * 1e3:.1f} ms "
          f"({seq.size / min(ts) / 1e9:.1f} GB/s) fwd={int(f.sum())} rev={int(r.sum())}")
