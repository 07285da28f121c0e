"""One-off differential stress run (not part of the suite): many seeded random batches through the C ABI
against the oracle -- window counts (all kernels, both upload paths) and explore (all block variants).
usage: python tests/parity_tools/stress.py [first_seed] [n_seeds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
from telomeric_identifier_b200.explore import run_labels
from oracle import oracle as O

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 150
ALPHA = [b"ACGT", b"ACGTacgt", b"ACGTN", b"ACGTNacgtn", b"ACGTNRYKMSWacgtn-", b"AC", b"Aa", b"ACGT" * 20 + b"N", b"ACGT" * 30 + b"acgt"]
KNOWN = [b"TTAGGG", b"AACCT", b"TTAGG", b"CCCTAA", b"AACCCT"]
t0 = time.time()
bad = 0
with T.Context(0) as ctx:
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        alpha = ALPHA[seed % len(ALPHA)]
        unit = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(2, 13))).tolist())
        big = seed % 7 == 0
        ids, seq, off = synth.random_records(seed, int(rng.integers(1, 8)), 400000 if big else 40000, alpha, plant=unit, p_empty=0.1)
        if seq.size and seed % 3 == 0:   # soft-masked stretch + an N gap somewhere
            a = int(rng.integers(0, seq.size)); seq[a:a + int(rng.integers(1, 3000))] |= 0x20
            b = int(rng.integers(0, seq.size)); seq[b:b + int(rng.integers(1, 2000))] = ord("N")
        # ---- explore
        d = float(rng.choice([0.5, 0.5, 0.25, 0.1, 0.01, 0.37]))
        kmin = int(rng.integers(1, 10)); kmax = kmin + int(rng.integers(0, 8))
        thr = int(rng.choice([0, 1, 2, 5, 100]))
        for fn in ("host", "resident"):
            if fn == "host":
                got = ctx.explore_runs(seq, off, d, kmin, kmax, thr)
            else:
                bt = ctx.upload(seq, off); got = ctx.explore_runs_resident(bt, d, kmin, kmax, thr); bt.free()
            want, wl = O.explore_runs(seq, off, d, kmin, kmax, thr, period_filter=False)
            ok = got.size == want.size and all((got[f] == want[f]).all() for f in ("record", "slice", "k", "first", "start", "end")) \
                and run_labels(seq, off, d, got) == wl
            if not ok:
                bad += 1; print("EXPLORE MISMATCH", seed, fn, d, kmin, kmax, thr, got.size, want.size)
        # ---- window counts
        qs = [unit, KNOWN[seed % len(KNOWN)]]
        if seed % 5 == 0:
            qs.append(bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 33))).tolist()))
        for W in (int(rng.integers(1, 3000)) if seq.size < 100000 else int(rng.integers(500, 3000)), 10000, int(rng.integers(20000, 70000))):
            f, r = ctx.window_counts(seq, off, W, qs)
            of, orr = O.window_counts(seq, off, W, qs, False)
            if not ((f == of).all() and (r == orr).all()):
                bad += 1; print("WINDOW MISMATCH", seed, W, qs)
            if seq.size:  # the resident path: K1 planes + the flat kernels (window_counts_flat.cuh)
                bt = ctx.upload(seq, off)
                f2, r2 = ctx.window_counts_resident(bt, W, qs)
                if not (np.array_equal(f2, of) and np.array_equal(r2, orr)):
                    bad += 1; print("RESIDENT WINDOW MISMATCH", seed, W, qs)
                bt.free()
    # ---- the upload paths on one bigger batch (forced packing, few threads so the DMA takes a share)
    ids, seq, off = synth.config4(scale=0.015)
    of, orr = O.window_counts(seq, off, 10000, [b"TTAGGG"], False, threads=4)
    for env in ({"TIDK_HOST_PACK": "1", "TIDK_PACK_THREADS": "2"}, {"TIDK_HOST_PACK": "1", "TIDK_PACK_THREADS": "16"},
                {"TIDK_HOST_PACK": "1", "TIDK_HYBRID_UPLOAD": "0"}, {"TIDK_HOST_PACK": "0"}):
        for k in ("TIDK_HOST_PACK", "TIDK_PACK_THREADS", "TIDK_HYBRID_UPLOAD"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            f, r = ctx.window_counts(seq, off, 10000, [b"TTAGGG"])
            if not ((f == of).all() and (r == orr).all()):
                bad += 1; print("UPLOAD MISMATCH", env, ctx.last_transfer())
print(f"stress: seeds {first}..{first + count - 1}, mismatches={bad}, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
