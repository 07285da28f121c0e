"""End to end through the C++ driver on a BASELINE config 5 read set (default 200 000 HiFi-like reads, ~3 Gb; the
full 2 000 000 reads are covered batch-wise by tests/parity_tools/c5_full.py): FASTA (page-cache warm) ->
`tidk explore --minimum 5 --maximum 12 --threshold 100 --distance 0.5` -> STDOUT table, compared with the oracle's
text (O(n) estimator, FASTA order).  usage: python tests/parity_tools/cli_e2e_c5.py [n_reads]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from telomeric_identifier_b200 import synth
from oracle import oracle as O
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TIDK = os.path.join(ROOT, "telomeric_identifier_b200", "bin", "tidk")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
d = "/tmp/tidk_cli_e2e_c5"; os.makedirs(d, exist_ok=True)
fa = os.path.join(d, "c5.fa")
t0 = time.perf_counter()
ids, seq, off = synth.config5(n_reads=n_reads)
with open(fa, "wb") as fh:
    step = 2000
    for r0 in range(0, n_reads, step):  # one sequence line per read (HiFi FASTA is usually unwrapped)
        parts = []
        for r in range(r0, min(n_reads, r0 + step)):
            parts.append(f">{ids[r]} len={int(off[r + 1] - off[r])}\n".encode())
            parts.append(seq[int(off[r]):int(off[r + 1])].tobytes())
            parts.append(b"\n")
        fh.write(b"".join(parts))
print(f"wrote {os.path.getsize(fa)} bytes ({seq.size} bases, {n_reads} reads) in {time.perf_counter() - t0:.1f}s", flush=True)
t0 = time.perf_counter()
runs, labels = O.explore_runs(seq, off, 0.5, 5, 12, 100, True, os.cpu_count() or 1)
counts = [(int(r["end"]) - int(r["start"])) // int(r["k"]) for r in runs]
rows = O.estimates(labels, counts, False)
want = "canonical_repeat_unit\tcount_repeat_runs_gt_100\n" + "".join(f"{k.decode()}\t{c}\n" for k, c in rows)
print(f"oracle: {len(labels)} runs, {len(rows)} rows in {time.perf_counter() - t0:.1f}s on {os.cpu_count()} threads", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    r = subprocess.run([TIDK, "explore", fa, "--minimum", "5", "--maximum", "12", "-t", "100", "--distance", "0.5", "--gpu-stats"],
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    stats = [l for l in r.stderr.splitlines() if l.startswith("{")]
    print(f"rep {rep}: rc={r.returncode} wall={dt:.3f}s  {seq.size / dt / 1e9:.3f} Gbases/s  {stats[-1] if stats else r.stderr[-300:]}  "
          f"table identical to the oracle: {r.stdout == want}", flush=True)
