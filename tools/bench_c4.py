"""C4 strong scaling: `tidk search --string TTAGGG --window 10000` on the synthetic 3.1 Gb /
24-chromosome genome with N gaps, cut into `world` contiguous window ranges (sharding.plan_shards:
whole chromosomes where possible, long ones split at multiples of W -- no halo, no collective).
Run under torchrun; every rank generates only the chromosomes its range touches.

Prints one JSON line on rank 0: aggregate Gbases/s (kernel-only from CUDA events, and wall per
step incl. the D2H of the counts), max over ranks.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    import torch
    import telomeric_identifier_b200 as T
    from telomeric_identifier_b200 import sharding, synth
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, Q = 10000, [b"TTAGGG"]
    lens = synth.config4_lengths(a.scale)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    plan = sharding.plan_shards(off, W, world)[rank]
    parts, plens = [], []
    cache = {}
    for rec, first, n in plan:
        if rec not in cache:
            cache = {rec: synth.config4_contig(rec, a.scale)[1]}  # one chromosome resident at a time
        s = cache[rec]
        lo, hi = first * W, min((first + n) * W, s.size)
        parts.append(s[lo:hi].copy()); plens.append(hi - lo)
    cache = None
    loff = np.zeros(len(plens) + 1, dtype=np.uint64); loff[1:] = np.cumsum(plens)
    lseq = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    parts = None
    ctx = T.Context(local)
    batch = ctx.upload(lseq, loff)
    f, r = ctx.window_counts_resident(batch, W, Q)
    for _ in range(a.warmup):
        ctx.window_counts_resident(batch, W, Q)
    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter(); kms = 0.0
    for _ in range(a.steps):
        f, r = ctx.window_counts_resident(batch, W, Q); kms += ctx.last_kernel_ms
    barrier()
    dt = time.perf_counter() - t0
    stats = torch.tensor([dt, kms / a.steps, float(lseq.size), float(f.sum()), float(r.sum())], device="cuda", dtype=torch.float64)
    if dist is not None:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = stats
    if rank == 0:
        total = float(sm[2])
        print(json.dumps({"workload": "C4 search TTAGGG W=10000, 3.1 Gb / 24 chr, N gaps", "scaling": "strong", "n_gpus": world,
                          "scale": a.scale, "total_bases": total, "steps": a.steps,
                          "kernel_ms_max_rank": float(mx[1]), "kernel_gbases_s": total / (float(mx[1]) * 1e-3) / 1e9,
                          "wall_ms_per_step": float(mx[0]) / a.steps * 1e3, "wall_gbases_s": total * a.steps / float(mx[0]) / 1e9,
                          "fwd_total": float(sm[3]), "rev_total": float(sm[4]),
                          "hbm_frac_per_gpu": (total / world) / (float(mx[1]) * 1e-3) / 1e9 / 6554.6}))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
