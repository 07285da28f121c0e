"""Explore from host memory (the host-pointer entry point) on C5-shaped reads: plain ASCII upload against the
two-ended STRICT-packed upload (planes expanded back to ASCII on the device).  usage: python tools/explore_upload.py [n_reads]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ids, seq, off = synth.config5(n_reads=n_reads)
out = {"reads": n_reads, "bases": int(seq.size), "host_threads": os.cpu_count()}
with T.Context(0) as ctx:
    pin = ctx.pinned_empty(seq.size)
    pin[:] = seq
    ref = None
    for name, env, src in (("ascii_pinned", "0", pin), ("packed_pinned", None, pin), ("packed_forced_pageable", "1", seq),
                           ("ascii_pageable", "0", seq)):
        if env is None: os.environ.pop("TIDK_EXPLORE_PACK", None)
        else: os.environ["TIDK_EXPLORE_PACK"] = env
        best = None
        for _ in range(5):
            t0 = time.perf_counter()
            runs = ctx.explore_runs(src, off, 0.5, 5, 12, 100)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        h2d, _, packed = ctx.last_transfer()
        if ref is None: ref = runs
        same = bool(runs.size == ref.size and (runs == ref).all())
        sm, pm = ctx.explore_timings()
        out[name] = {"call_ms": best * 1e3, "gbases_per_s": seq.size / best / 1e9, "h2d_bytes": int(h2d), "packed": bool(packed),
                     "runs": int(runs.size), "same_runs_as_ascii": same, "scan_ms": sm, "post_ms": pm}
        print(name, json.dumps(out[name]), flush=True)
    # a genome with soft-masked stretches (C2: 2 % of the bases lower-case in 1-5 kb blocks), d = 0.5
    ids, seq, off = synth.config2()
    pin = ctx.pinned_empty(seq.size)
    pin[:] = seq
    for name, env in (("c2_genome_ascii_pinned", "0"), ("c2_genome_packed_pinned", None)):
        if env is None: os.environ.pop("TIDK_EXPLORE_PACK", None)
        else: os.environ["TIDK_EXPLORE_PACK"] = env
        best = None
        for _ in range(5):
            t0 = time.perf_counter()
            runs = ctx.explore_runs(pin, off, 0.5, 5, 12, 100)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        h2d, _, packed = ctx.last_transfer()
        out[name] = {"call_ms": best * 1e3, "gbases_per_s": seq.size / best / 1e9, "h2d_bytes": int(h2d), "packed": bool(packed), "runs": int(runs.size)}
        print(name, json.dumps(out[name]), flush=True)
print(json.dumps(out))
