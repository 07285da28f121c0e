"""Full-size, full-output parity of BASELINE configs 1-4 (config 5: c5_full.py with TIDK_C5_CHECK_ALL=1): the whole text
the reference would write -- every row of the TSV / bedgraph, every line of the explore table, and every field of every
explore run -- produced through the C ABI on the GPU and by the oracle (all host threads), compared as sha256 of the
bytes.  Writes gpurun_out/r02_full_parity.json (copied to profiles/ when it says "equal" everywhere).

    python tests/parity_tools/full_parity.py            (TIDK_PARITY_SCALE=0.05 for a quick look)"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth, hostlogic
from telomeric_identifier_b200.search import format_rows, HEADER
from telomeric_identifier_b200.explore import run_labels
from telomeric_identifier_b200.clades import return_telomere_sequence
from oracle import oracle as O

SCALE = float(os.environ.get("TIDK_PARITY_SCALE", "1.0"))
THREADS = os.cpu_count() or 1
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(text):
    return hashlib.sha256(text if isinstance(text, bytes) else text.encode()).hexdigest()


def oracle_rows(ids, off, window, labels, fwd, rev, extension):
    """The oracle's own formatter (oracle.search_text / find_text row loops) on count arrays: written separately from the
    product's format_rows on purpose."""
    out = []
    p = 0
    for r, rid in enumerate(ids):
        n = int(off[r + 1] - off[r])
        nw = (n + window - 1) // window
        for lab in labels:
            for i in range(nw):
                end = min((i + 1) * window, n)
                if extension == "tsv":
                    out.append(f"{rid}\t{end}\t{fwd[p]}\t{rev[p]}\t{lab}\n")
                else:
                    out.append(f"{rid}\t{i * window}\t{end}\t{fwd[p] + rev[p]}\n")
                p += 1
    return "".join(out)


def counts_case(ctx, name, ids, seq, off, window, queries, uppercase, extensions):
    res = {}
    t0 = time.perf_counter()
    f, r = ctx.window_counts(seq, off, window, [q.encode() for q in queries])
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    of, orv = O.window_counts(seq, off, window, [q.encode() for q in queries], uppercase, THREADS)
    t_cpu = time.perf_counter() - t0
    for ext in extensions:
        g = (HEADER if ext == "tsv" else "") + format_rows(ids, off, window, queries, f, r, ext)
        o = (O.SEARCH_HEADER if ext == "tsv" else "") + oracle_rows(ids, off, window, queries, of.tolist(), orv.tolist(), ext)
        res[ext] = {"gpu_sha256": sha(g), "oracle_sha256": sha(o), "bytes": len(o), "rows": o.count("\n"), "equal": sha(g) == sha(o)}
    res["bases"] = int(off[-1]); res["gpu_call_s"] = t_gpu; res["oracle_s"] = t_cpu; res["oracle_threads"] = THREADS
    print(name, json.dumps(res), flush=True)
    return res


def explore_case(ctx, name, seq, off, d, kmin, kmax, thr):
    t0 = time.perf_counter()
    runs = ctx.explore_runs(seq, off, d, kmin, kmax, thr)
    batch = ctx.upload(seq, off)
    runs_res = ctx.explore_runs_resident(batch, d, kmin, kmax, thr)
    batch.free()
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    want, wlabels = O.explore_runs(seq, off, d, kmin, kmax, thr, False, THREADS)
    t_cpu = time.perf_counter() - t0
    fields = ("record", "slice", "k", "first", "start", "end")

    def runs_sha(a, labels):
        h = hashlib.sha256()
        for f in fields:
            h.update(np.ascontiguousarray(a[f]).tobytes())
        h.update(b"".join(labels))
        return h.hexdigest()

    glabels = run_labels(seq, off, d, runs)
    g, g2, o = runs_sha(runs, glabels), runs_sha(runs_res, run_labels(seq, off, d, runs_res)), runs_sha(want, wlabels)

    def table(a, labels):  # explore.rs:140-143 after the period filter (explore.rs:265-274)
        keep = [i for i, l in enumerate(labels) if hostlogic.period(l) >= 3]
        counts = ((a["end"] - a["start"]) // a["k"]).tolist()
        rows = hostlogic.estimates([labels[i] for i in keep], [counts[i] for i in keep])
        return f"canonical_repeat_unit\tcount_repeat_runs_gt_{thr}\n" + "".join(f"{k.decode()}\t{c}\n" for k, c in rows)

    def oracle_table(a, labels):
        keep = [i for i, l in enumerate(labels) if O.period(l) >= 3]
        counts = [(int(x["end"]) - int(x["start"])) // int(x["k"]) for x in a]
        rows = O.estimates([labels[i] for i in keep], [counts[i] for i in keep], literal=len(keep) <= 3000)
        return f"canonical_repeat_unit\tcount_repeat_runs_gt_{thr}\n" + "".join(f"{k.decode()}\t{c}\n" for k, c in rows)

    tg, to = table(runs, glabels), oracle_table(want, wlabels)
    res = {"runs": int(want.size), "run_list": {"gpu_host_pointer_sha256": g, "gpu_resident_sha256": g2, "oracle_sha256": o,
                                                "equal": g == o and g2 == o, "of": "record, slice, k, first, start, end, label of every run"},
           "table": {"gpu_sha256": sha(tg), "oracle_sha256": sha(to), "bytes": len(to), "equal": sha(tg) == sha(to), "text": to},
           "gpu_s": t_gpu, "oracle_s": t_cpu, "oracle_threads": THREADS}
    print(name, json.dumps(res), flush=True)
    return res


def main():
    out = {"scale": SCALE, "what": __doc__.split("\n\n")[0]}
    with T.Context(0) as ctx:
        ids, seq, off = synth.config1(SCALE)
        out["C1 search --string TTAGG --window 10000 (100 Mb, 10 contigs)"] = counts_case(ctx, "C1", ids, seq, off, 10000, ["TTAGG"], True, ("tsv", "bedgraph"))
        ids, seq, off = synth.config2(SCALE)
        lep = return_telomere_sequence("Lepidoptera", os.path.join(os.path.dirname(HERE), "golden", "clades_min.csv"))
        out["C2 find --clade Lepidoptera (500 Mb, 30 chromosomes)"] = counts_case(ctx, "C2", ids, seq, off, 10000, list(lep), False, ("tsv",))
        hym = return_telomere_sequence("Hymenoptera", os.path.join(os.path.dirname(HERE), "golden", "clades_min.csv"))
        out["C2b find --clade Hymenoptera (15 repeats, same genome)"] = counts_case(ctx, "C2b", ids, seq, off, 10000, list(hym), False, ("tsv",))
        out["C3 explore -m 5 -x 12 -t 100 --distance 0.01 (same genome)"] = explore_case(ctx, "C3", seq, off, 0.01, 5, 12, 100)
        out["C3b explore --distance 0.5 (same genome: slices of 4-12 Mb, vertical scan)"] = explore_case(ctx, "C3b", seq, off, 0.5, 5, 12, 100)
        del seq
        ids, seq, off = synth.config4(SCALE)
        out["C4 search --string TTAGGG --window 10000 (3.1 Gb, 24 chromosomes, N gaps)"] = counts_case(ctx, "C4", ids, seq, off, 10000, ["TTAGGG"], True, ("tsv", "bedgraph"))
    out["all_equal"] = all(v.get("equal", True) for c in out.values() if isinstance(c, dict) for v in c.values() if isinstance(v, dict))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_full_parity.json", "w") as f:
        json.dump(out, f, indent=1)
    print("all_equal", out["all_equal"])


if __name__ == "__main__":
    main()
