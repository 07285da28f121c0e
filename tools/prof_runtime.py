"""Run-time-pattern kernels vs compile-time ones on C2 (resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
ids, seq, off = synth.config2(scale=float(os.environ.get("TIDK_BENCH_SCALE", "1.0")))
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    for qs in ([b"AACCT"], [b"GATTC"], [b"GATTACA"], [b"ACGTTGCAAC"], [b"ACGTACGTACGTACGTACGT"], [b"GATTACA", b"CATTAG"],
               [b"GATTACA", b"CATTAG", b"TTGACC", b"GGGTTT"], [b"TTANGG"], [b"A" * 40]):
        ms = []
        for _ in range(6):
            ctx.window_counts_resident(b, 10000, qs); ms.append(ctx.last_kernel_ms)
        print(f"{str([q.decode()[:12] for q in qs]):60s} best {min(ms[1:])*1e3:9.1f} us  launches={ctx.last_kernel_launches}")
