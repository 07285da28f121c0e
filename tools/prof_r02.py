"""ncu target of round 2: the explore scan on resident planes (C5-shaped reads, k = 5..12 and k = 6) and the window counter on
the C4-shaped genome from both sources (resident planes; ASCII), a few launches each.
usage: python tools/prof_r02.py     (TIDK_PROF_C4_SCALE=0.25 by default)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

with T.Context(0) as ctx:
    ids, seq, off = synth.config5(n_reads=int(os.environ.get("TIDK_PROF_READS", "100000")))
    b = ctx.upload(seq, off)
    for kmin, kmax in ((5, 12), (5, 12), (5, 12), (6, 6)):
        runs = ctx.explore_runs_resident(b, 0.5, kmin, kmax, 100)
        sm, pm = ctx.explore_timings()
        print(f"C5 k={kmin}..{kmax} bases={seq.size} scan_ms={sm:.4f} post_ms={pm:.4f} runs={runs.size}", flush=True)
    b.free()
    ids, seq, off = synth.config4(scale=float(os.environ.get("TIDK_PROF_C4_SCALE", "0.25")))
    b = ctx.upload(seq, off)
    for src in ("planes", "ascii"):
        if src == "ascii":
            os.environ["TIDK_PLANES_COUNT"] = "0"
        for _ in range(3):
            f, r = ctx.window_counts_resident(b, 10000, [b"TTAGGG"])
            print(f"C4 {src} bases={seq.size} kernel_ms={ctx.last_kernel_ms:.4f} sum={int(f.sum())},{int(r.sum())}", flush=True)
        for W in (1000, 2500):
            f, r = ctx.window_counts_resident(b, W, [b"TTAGGG"])
            f, r = ctx.window_counts_resident(b, W, [b"TTAGGG"])
            print(f"C4 {src} W={W} kernel_ms={ctx.last_kernel_ms:.4f}", flush=True)
        f, r = ctx.window_counts_resident(b, 10000, [b"GATTACAGATTA"])
        f, r = ctx.window_counts_resident(b, 10000, [b"GATTACAGATTA"])
        print(f"C4 {src} runtime 12-mer kernel_ms={ctx.last_kernel_ms:.4f}", flush=True)
    os.environ.pop("TIDK_PLANES_COUNT", None)
    # the fused clade kernel (find --clade Hymenoptera: 15 repeats in one pass, per-window form on the resident planes)
    from telomeric_identifier_b200 import clades
    db = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "clades_min.csv")
    hym = [r.encode() for r in clades.return_telomere_sequence("Hymenoptera", db)]
    for _ in range(2):
        f, r = ctx.window_counts_resident(b, 10000, hym)
        print(f"C4 find --clade Hymenoptera ({len(hym)} repeats) kernel_ms={ctx.last_kernel_ms:.4f} launches={ctx.last_kernel_launches}", flush=True)
    b.free()
