"""End to end through the C++ driver: FASTA file (page-cache warm) -> TSV file, on C2."""
import os, subprocess, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from telomeric_identifier_b200 import synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIDK = os.path.join(ROOT, "telomeric_identifier_b200", "bin", "tidk")
DB = os.path.join(ROOT, "tests", "golden", "clades_min.csv")
scale = float(os.environ.get("TIDK_BENCH_SCALE", "1.0"))
ids, seq, off = synth.config2(scale=scale)
d = "/tmp/tidk_cli_e2e"; os.makedirs(d, exist_ok=True)
fa = os.path.join(d, "c2.fa")
t0 = time.perf_counter()
with open(fa, "wb") as fh:
    nl = np.uint8(10)
    for i, rid in enumerate(ids):
        s = seq[int(off[i]):int(off[i + 1])]
        fh.write(f">{rid}\n".encode())
        full = (s.size // 60) * 60
        body = np.empty((full // 60, 61), dtype=np.uint8)
        body[:, :60] = s[:full].reshape(-1, 60); body[:, 60] = nl
        fh.write(body.tobytes())
        if full < s.size:
            fh.write(s[full:].tobytes() + b"\n")
print(f"wrote {os.path.getsize(fa)} bytes in {time.perf_counter() - t0:.1f}s", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    r = subprocess.run([TIDK, "find", fa, "-c", "Lepidoptera", "-o", "c2", "-d", d, "--database", DB, "--gpu-stats"],
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    stats = [l for l in r.stderr.splitlines() if l.startswith("{")]
    print(f"rep {rep}: rc={r.returncode} wall={dt:.3f}s  {seq.size / dt / 1e9:.3f} Gbases/s  {stats[-1] if stats else r.stderr[-300:]}", flush=True)
