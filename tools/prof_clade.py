"""Timing of `find --clade X` (all repeats of the clade) on C2, resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import clades, synth
DB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "clades_min.csv")
ids, seq, off = synth.config2(scale=float(os.environ.get("TIDK_BENCH_SCALE", "1.0")))
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    for clade in ("Lepidoptera", "Hemiptera", "Coleoptera", "Hymenoptera"):
        reps = [r.encode() for r in clades.return_telomere_sequence(clade, DB)]
        ms = []
        for _ in range(6):
            ctx.window_counts_resident(b, 10000, reps); ms.append(ctx.last_kernel_ms)
        fused = min(ms[1:])
        ms2 = []
        for _ in range(4):
            t = 0.0
            for r in reps:
                ctx.window_counts_resident(b, 10000, [r]); t += ctx.last_kernel_ms
            ms2.append(t)
        print(f"{clade:12s} Q={len(reps):2d} fused {fused:.4f} ms ({ctx.last_kernel_launches} launch) = {seq.size/fused/1e6:7.1f} Gbases/s | one pass per repeat {min(ms2[1:]):.4f} ms")
