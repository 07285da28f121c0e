"""Window counter on the resident planes of the C4-shaped genome: flat form vs per-window form (TIDK_FLAT=0), W = 10000 / 2500 / 1000 / 100,
a compile-time motif and run-time patterns.   usage: python tools/prof_flat.py  (TIDK_PROF_C4_SCALE=0.25)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

ids, seq, off = synth.config4(scale=float(os.environ.get("TIDK_PROF_C4_SCALE", "0.25")))
with T.Context(0) as ctx:
    b = ctx.upload(seq, off)
    ref = {}
    for flat in ("0", "1"):
        os.environ["TIDK_FLAT"] = flat
        for q in (b"TTAGGG", b"GATTACAGATTA", b"GATTAC"):
            for W in (10000, 2500, 1000, 100):
                ms = []
                for _ in range(5):
                    f, r = ctx.window_counts_resident(b, W, [q])
                    ms.append(ctx.last_kernel_ms)
                key = (q, W)
                if flat == "0":
                    ref[key] = (f.copy(), r.copy())
                same = bool(np.array_equal(ref[key][0], f) and np.array_equal(ref[key][1], r))
                print(f"flat={flat} q={q.decode():14s} W={W:6d} kernel_ms min={min(ms[1:]):.4f} med={sorted(ms[1:])[2]:.4f} "
                      f"Gbases/s={seq.size / min(ms[1:]) / 1e6:8.0f} sum={int(f.sum())},{int(r.sum())} same_as_flat0={same}", flush=True)
    os.environ.pop("TIDK_FLAT", None)
    b.free()
