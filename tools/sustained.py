"""Does a long run of back-to-back launches keep the rate of a short one?  Per-group kernel times of the C4 window
counter issued asynchronously, a plain device copy loop beside it, nvidia-smi sampled every 20 ms."""
import os, subprocess, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import telomeric_identifier_b200 as T
import bench

Q = ("timestamp,clocks.sm,clocks.mem,power.draw,power.draw.instant,temperature.gpu,temperature.memory,clocks_event_reasons.active,"
     "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_power_brake_slowdown")
smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                       stdout=open("gpurun_out/sustained_smi.csv", "w"), stderr=subprocess.DEVNULL)
pieces, chrom, off, goff = bench.c4_shard(0, 1, 16)
ctx = T.Context(0)
pseq = ctx.pinned_empty(int(off[-1])); bench.fill_shard(pseq, pieces, chrom, off); chrom = None
batch = ctx.upload(pseq, off)
nwin = int(T.api.n_windows(off, 10000).sum())
outs = [ctx.pinned_empty(2 * nwin * 8).view(np.uint64) for _ in range(2)]
res = {}
for group, gap_ms in ((4, 0), (4, 0), (1, 0), (4, 2.0)):
    time.sleep(1.0)
    ks = []
    t0 = time.perf_counter()
    for g in range(300 // group):
        for p in range(group):
            o = outs[p & 1]
            ctx.window_counts_resident_async(batch, 10000, [b"TTAGGG"], (o[:nwin], o[nwin:]))
        ms, nl = ctx.wait()
        ks.append(ms / group)
        if gap_ms:
            time.sleep(gap_ms * 1e-3)
    res[f"group{group}_gap{gap_ms}"] = {"first5": ks[:5], "at_10pct": ks[len(ks) // 10], "median": float(np.median(ks)), "last5": ks[-5:],
                                       "wall_s": time.perf_counter() - t0}
time.sleep(1.0)
# plain copy, 1 GiB -> 1 GiB
a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(301)]
evs[0].record()
for i in range(300):
    b.copy_(a); evs[i + 1].record()
torch.cuda.synchronize()
cs = [2 * (1 << 30) / (evs[i].elapsed_time(evs[i + 1]) * 1e-3) / 1e9 for i in range(300)]
res["copy_gbs"] = {"first5": cs[:5], "median": float(np.median(cs)), "last5": cs[-5:]}
smi.terminate()
print(json.dumps(res, indent=1))
