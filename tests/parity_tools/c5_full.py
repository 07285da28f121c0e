"""BASELINE config 5 at full size: `explore --minimum 5 --maximum 12 --threshold 100 --distance 0.5` on
2 000 000 HiFi-like reads (~30 Gb), streamed through the C ABI in batches of reads the way a FASTA reader
would hand them over (explore.rs:97-121 walks the records once per k; here every batch is scanned once for
all k and the per-k run lists are concatenated in batch order = file order).

Per batch: the host-pointer call (H2D inside the timed region), the resident call (scan / post device
times), both run lists compared with each other, and the runs of a subsample of reads (every read with a
planted array + random ones; the whole of batch 0) compared with the oracle.  At the end the period filter
and the estimator run over the runs of all batches (hostlogic.py) and the table is printed.

The random background of a batch is drawn on the GPU with torch (numpy needs ~15 s per 1.5 Gb batch) and
copied to host memory before anything is timed; lengths and planted arrays follow synth.config5.
Writes gpurun_out/c5_full.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth, hostlogic
from telomeric_identifier_b200.explore import run_labels
from oracle import oracle as O

N_READS = int(os.environ.get("TIDK_C5_READS", "2000000"))
BATCH = int(os.environ.get("TIDK_C5_BATCH", "100000"))
N_CHECK = int(os.environ.get("TIDK_C5_CHECK", "1500"))  # random reads per batch checked against the oracle
CHECK_ALL = os.environ.get("TIDK_C5_CHECK_ALL", "0") == "1"  # every read of every batch against the oracle (all host threads)
D, KMIN, KMAX, THR = 0.5, 5, 12, 100
THREADS = os.cpu_count() or 1


def make_batch(b: int, n_reads: int, host: np.ndarray, lut: torch.Tensor):
    rng = synth._rng(5, 1000 + b)
    lens = np.clip(rng.normal(15_000, 2_000, size=n_reads), 5_000, 25_000).astype(np.int64)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens).astype(np.uint64)
    total = int(off[-1])
    g = torch.Generator(device="cuda")
    g.manual_seed(50_000 + b)
    codes = torch.randint(0, 4, (total,), device="cuda", dtype=torch.uint8, generator=g)
    seq = host[:total]
    torch.from_numpy(seq).copy_(lut[codes.to(torch.int32)])
    del codes
    planted = np.flatnonzero(rng.random(n_reads) < 0.01)
    for r in planted.tolist():
        o, e = int(off[r]), int(off[r + 1])
        unit = b"TTAGGG" if rng.random() < 0.5 else b"CCCTAA"
        arr = synth._array(rng, unit, int(rng.integers(150, 1501)), 0.0005, 0.0)
        view = seq[o:e]
        if rng.random() < 0.5:
            synth._put(view, 0, arr)
        else:
            synth._put_right(view, view.size, arr)
    return seq, off, planted


def labels_of(seq, off, runs):
    return run_labels(seq, off, D, runs)


def same_runs(a, b, fields=("record", "slice", "k", "first", "start", "end", "label_pos")):
    return a.size == b.size and all((a[f] == b[f]).all() for f in fields)


# against the oracle the label is compared, not its position: the reference labels a run with the chunk of the pair
# that closes it (explore.rs:326-332), the library with the run's first chunk -- equal after upper-casing by
# construction of a run (explore.rs:324)
ORACLE_FIELDS = ("record", "slice", "k", "first", "start", "end")


def main():
    nb = (N_READS + BATCH - 1) // BATCH
    res = {"config": "C5: explore -m 5 -x 12 -t 100 --distance 0.5", "reads": N_READS, "batch_reads": BATCH,
           "batches": nb, "host_threads": THREADS}
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    import hashlib
    sha_gpu, sha_ora = hashlib.sha256(), hashlib.sha256()
    per_k_labels = {k: [] for k in range(KMIN, KMAX + 1)}
    per_k_counts = {k: [] for k in range(KMIN, KMAX + 1)}
    tot = dict(bases=0, scanned=0, runs=0, e2e_s=0.0, scan_ms=0.0, post_ms=0.0, res_wall_s=0.0, checked_reads=0,
               checked_runs=0, gen_s=0.0, oracle_s=0.0)
    with T.Context(0) as ctx:
        host = ctx.pinned_empty(int(BATCH * 15_400) + (1 << 20))
        for b in range(nb):
            n = min(BATCH, N_READS - b * BATCH)
            t0 = time.perf_counter()
            seq, off, planted = make_batch(b, n, host, lut)
            torch.cuda.synchronize()
            tot["gen_s"] += time.perf_counter() - t0
            lens = (off[1:] - off[:-1]).astype(np.int64)
            scanned = int((2 * np.ceil(lens.astype(np.float64) * D)).sum())
            # host-pointer entry point: upload + scan + run building + run list back
            t0 = time.perf_counter()
            runs = ctx.explore_runs(seq, off, D, KMIN, KMAX, THR)
            e2e = time.perf_counter() - t0
            # resident entry point
            buf = ctx.upload(seq, off)
            ctx.explore_runs_resident(buf, D, KMIN, KMAX, THR)
            t0 = time.perf_counter()
            runs_r = ctx.explore_runs_resident(buf, D, KMIN, KMAX, THR)
            rw = time.perf_counter() - t0
            sm, pm = ctx.explore_timings()
            buf.free()
            assert same_runs(runs, runs_r), f"batch {b}: host-pointer and resident run lists differ"
            # arrival order inside the batch: (k, record, slice, start)
            key = np.lexsort((runs["start"], runs["slice"], runs["record"], runs["k"]))
            assert (key == np.arange(runs.size)).all(), f"batch {b}: runs are not in arrival order"
            cnt = (runs["end"] - runs["start"]) // runs["k"]
            assert (cnt > THR).all()
            # oracle on a subsample of reads (runs of a read do not depend on the other reads)
            t0 = time.perf_counter()
            if CHECK_ALL or (b == 0 and os.environ.get("TIDK_C5_FULL_BATCH0", "1") == "1"):
                pick = np.arange(n)
            else:
                rng = synth._rng(5, 5000 + b)
                pick = np.unique(np.concatenate([planted, rng.integers(0, n, size=N_CHECK)]))
            sub_off = np.zeros(pick.size + 1, dtype=np.uint64)
            sub_off[1:] = np.cumsum(lens[pick]).astype(np.uint64)
            if pick.size == n:
                sub_seq = seq
            else:
                sub_seq = np.concatenate([seq[int(off[r]):int(off[r + 1])] for r in pick.tolist()])
            want, _ = O.explore_runs(sub_seq, sub_off, D, KMIN, KMAX, THR, period_filter=False, threads=THREADS)
            remap = np.full(n, -1, dtype=np.int64)
            remap[pick] = np.arange(pick.size)
            got = runs[remap[runs["record"]] >= 0].copy()
            got["record"] = remap[got["record"]].astype(np.uint32)
            assert same_runs(got, want, ORACLE_FIELDS), f"batch {b}: runs of the checked reads differ from the oracle ({got.size} vs {want.size})"
            assert labels_of(sub_seq, sub_off, got) == labels_of(sub_seq, sub_off, want), f"batch {b}: run labels differ from the oracle"
            for h, arr in ((sha_gpu, got), (sha_ora, want)):  # the run lists themselves, field by field, batch after batch
                for f in ORACLE_FIELDS:
                    h.update(np.ascontiguousarray(arr[f]).tobytes())
            tot["oracle_s"] += time.perf_counter() - t0
            tot["checked_reads"] += int(pick.size)
            tot["checked_runs"] += int(want.size)
            # host side of explore: labels, period filter (explore.rs:265-274), per-k arrival lists
            labs = labels_of(seq, off, runs)
            ks = runs["k"].tolist()
            cl = cnt.tolist()
            pcache = {}
            for lab, k, c in zip(labs, ks, cl):
                pr = pcache.get(lab)
                if pr is None:
                    pr = pcache[lab] = hostlogic.period(lab)
                if pr >= 3:
                    per_k_labels[k].append(lab)
                    per_k_counts[k].append(c)
            tot["bases"] += int(off[-1]); tot["scanned"] += scanned; tot["runs"] += int(runs.size)
            tot["e2e_s"] += e2e; tot["scan_ms"] += sm; tot["post_ms"] += pm; tot["res_wall_s"] += rw
            print(f"batch {b:2d}: reads={n} bases={int(off[-1])} runs={runs.size} e2e={e2e * 1e3:.1f} ms "
                  f"({int(off[-1]) / e2e / 1e9:.1f} Gbases/s) scan={sm:.3f} ms post={pm:.3f} ms resident wall={rw * 1e3:.2f} ms "
                  f"oracle-checked reads={pick.size} runs={want.size}", flush=True)
    t0 = time.perf_counter()
    labels = [l for k in range(KMIN, KMAX + 1) for l in per_k_labels[k]]
    counts = [c for k in range(KMIN, KMAX + 1) for c in per_k_counts[k]]
    table = hostlogic.estimates(labels, counts)
    est_s = time.perf_counter() - t0
    res.update(tot)
    res.update({
        "kept_runs_after_period_filter": len(labels), "estimator_s": est_s, "table_rows": len(table),
        "table_head": [[k.decode(), int(c)] for k, c in table[:8]],
        "e2e_gbases_per_s": tot["bases"] / tot["e2e_s"] / 1e9,
        "scan_gbases_per_s": tot["scanned"] / (tot["scan_ms"] * 1e-3) / 1e9,
        "scan_plus_post_gbases_per_s": tot["scanned"] / ((tot["scan_ms"] + tot["post_ms"]) * 1e-3) / 1e9,
        "resident_call_gbases_per_s": tot["scanned"] / tot["res_wall_s"] / 1e9,
        "parity": "host-pointer == resident run lists on every batch; oracle run lists identical on the checked reads",
        "run_lists_sha256": {"gpu": sha_gpu.hexdigest(), "oracle": sha_ora.hexdigest(), "equal": sha_gpu.hexdigest() == sha_ora.hexdigest(),
                             "of": "record, slice, k, first, start, end of every run of the checked reads, batch after batch"},
    })
    print(f"canonical_repeat_unit\tcount_repeat_runs_gt_{THR}")
    for k, c in table[:8]:
        print(f"{k.decode()}\t{c}")
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.environ.get("TIDK_C5_OUT", "gpurun_out/c5_full.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
