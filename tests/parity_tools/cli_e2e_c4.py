"""End to end through the C++ driver on BASELINE config 4: a 3.1 Gb / 24-chromosome FASTA with N gaps (page-cache
warm) -> `tidk search --string TTAGGG --window 10000` -> TSV.  The rows of two chromosomes are checked against the
oracle's text.  usage: python tests/parity_tools/cli_e2e_c4.py [gpus ...]   (default: 1)"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from telomeric_identifier_b200 import synth
from oracle import oracle as O
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TIDK = os.path.join(ROOT, "telomeric_identifier_b200", "bin", "tidk")
scale = float(os.environ.get("TIDK_BENCH_SCALE", "1.0"))
d = "/tmp/tidk_cli_e2e_c4"; os.makedirs(d, exist_ok=True)
fa = os.path.join(d, "c4.fa")
t0 = time.perf_counter()
bases, keep = 0, {}
with open(fa, "wb") as fh:
    for i in range(len(synth.config4_lengths(scale))):
        rid, s = synth.config4_contig(i, scale)
        bases += s.size
        if i >= 20 and len(keep) < 2:  # two small chromosomes for the oracle check
            keep[rid] = s.tobytes()
        fh.write(f">{rid} synthetic\n".encode())
        full = (s.size // 60) * 60
        body = np.empty((full // 60, 61), dtype=np.uint8)
        body[:, :60] = s[:full].reshape(-1, 60); body[:, 60] = 10
        fh.write(body.tobytes())
        if full < s.size:
            fh.write(s[full:].tobytes() + b"\n")
        del s, body
print(f"wrote {os.path.getsize(fa)} bytes ({bases} bases) in {time.perf_counter() - t0:.1f}s", flush=True)
want = {rid: O.search_text([rid], [sq], b"TTAGGG", 10000).splitlines()[1:] for rid, sq in keep.items()}
for gpus in [int(x) for x in sys.argv[1:]] or [1]:
    for rep in range(3):
        out = os.path.join(d, f"out{gpus}")
        t0 = time.perf_counter()
        r = subprocess.run([TIDK, "search", fa, "-s", "TTAGGG", "-w", "10000", "-o", "c4", "-d", out, "--gpus", str(gpus), "--gpu-stats"],
                           capture_output=True, text=True)
        dt = time.perf_counter() - t0
        stats = [l for l in r.stderr.splitlines() if l.startswith("{")]
        print(f"gpus={gpus} rep {rep}: rc={r.returncode} wall={dt:.3f}s  {bases / dt / 1e9:.3f} Gbases/s  {stats[-1] if stats else r.stderr[-300:]}", flush=True)
    rows = {}
    for line in open(os.path.join(out, "c4_telomeric_repeat_windows.tsv")).read().splitlines()[1:]:
        rows.setdefault(line.split("\t", 1)[0], []).append(line)
    for rid, w in want.items():
        assert rows[rid] == w, f"{rid}: rows differ from the oracle"
    print(f"gpus={gpus}: rows of {', '.join(want)} identical to the oracle ({sum(len(w) for w in want.values())} windows); "
          f"{sum(len(v) for v in rows.values())} rows in all", flush=True)
