"""One process, one tidk_cuda context over N devices (tidk_cuda_create with n_devices = N, csrc/multi_device.cuh): the C4 search
(resident, asynchronous passes) and the C5-shaped explore (resident and from host memory) on 1 .. N GPUs of the box.
usage: python tools/multi_gpu.py N      -> gpurun_out/r02_multi_device.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
import bench

NMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 1
SCALE = float(os.environ.get("TIDK_BENCH_SCALE", "1.0"))
res = {"what": __doc__.split("\n")[0], "scale": SCALE, "runs": []}
pieces, chrom, off, goff = bench.c4_shard(0, 1, os.cpu_count() or 1)
n5 = max(1000, int(200_000 * SCALE))
ids5, seq5, off5 = synth.config5(n5)
lens5 = (off5[1:] - off5[:-1]).astype(np.int64)
scanned5 = int((2 * ((lens5 + 1) // 2)).sum())
base = {}
for n in [x for x in (1, 2, 4, 8) if x <= NMAX]:
    ctx = T.Context(list(range(n)))
    pseq = ctx.pinned_empty(int(off[-1])); bench.fill_shard(pseq, pieces, chrom, off)
    nwin = int(T.api.n_windows(off, 10000).sum())
    outs = []
    for _ in range(2):
        po = ctx.pinned_empty(2 * nwin * 8).view(np.uint64)
        outs.append((po[:nwin], po[nwin:]))
    t0 = time.perf_counter(); batch = ctx.upload(pseq, off); t_up = time.perf_counter() - t0
    f, r = ctx.window_counts_resident(batch, 10000, [b"TTAGGG"]); f, r = f.copy(), r.copy()
    if n == 1: base["c4"] = (f, r)
    P = 32
    def step():
        for p in range(P):
            ctx.window_counts_resident_async(batch, 10000, [b"TTAGGG"], outs[p & 1])
        return ctx.wait()
    for _ in range(3): step()
    t0 = time.perf_counter(); km = 0.0
    for _ in range(10):
        ms, nl = step(); km += ms
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(outs[0][0], base["c4"][0]) and np.array_equal(outs[1][1], base["c4"][1]))
    # end to end from host memory
    ctx.window_counts(pseq, off, 10000, [b"TTAGGG"])
    t1 = time.perf_counter()
    for _ in range(3): ef, er = ctx.window_counts(pseq, off, 10000, [b"TTAGGG"])
    de = (time.perf_counter() - t1) / 3
    batch.free()
    row = {"devices": n, "c4_search": {"bases": int(off[-1]), "upload_s": t_up, "wall_ms_per_pass": dt / 10 / P * 1e3, "gbases_per_s_wall": int(off[-1]) * P * 10 / dt / 1e9,
                                       "kernel_ms_per_pass_max_over_devices": km / 10 / P, "counts_equal_to_one_device": ok,
                                       "e2e_host_pointer_ms": de * 1e3, "e2e_gbases_per_s": int(off[-1]) / de / 1e9,
                                       "e2e_equal": bool(np.array_equal(ef, base["c4"][0]) and np.array_equal(er, base["c4"][1]))}}
    # explore
    p5 = ctx.pinned_empty(seq5.size); p5[:] = seq5
    b5 = ctx.upload(p5, off5)
    best = None
    for _ in range(5):
        t0 = time.perf_counter(); runs = ctx.explore_runs_resident(b5, 0.5, 5, 12, 100); w = time.perf_counter() - t0
        best = w if best is None else min(best, w)
    if n == 1: base["c5"] = runs.copy()
    e2e = None
    for _ in range(3):
        t0 = time.perf_counter(); r2 = ctx.explore_runs(p5, off5, 0.5, 5, 12, 100); w = time.perf_counter() - t0
        e2e = w if e2e is None else min(e2e, w)
    b5.free()
    row["c5_explore"] = {"reads": n5, "scanned_bases": scanned5, "resident_call_wall_ms": best * 1e3, "resident_gbases_per_s": scanned5 / best / 1e9,
                         "host_pointer_call_ms": e2e * 1e3, "host_pointer_gbases_per_s": seq5.size / e2e / 1e9, "runs": int(runs.size),
                         "runs_equal_to_one_device": bool(np.array_equal(runs, base["c5"]) and np.array_equal(r2, base["c5"]))}
    print(json.dumps(row), flush=True)
    res["runs"].append(row)
    ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_multi_device.json", "w"), indent=1)
