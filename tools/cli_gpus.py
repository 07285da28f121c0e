"""`tidk search --gpus N` against `--gpus 1` on a C2-shaped FASTA: identical files, wall time of both."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from telomeric_identifier_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(os.environ.get("TIDK_BENCH_SCALE", "0.4"))
ids, seq, off = synth.config2(scale=scale)
fa = "/tmp/cli_gpus.fa"
with open(fa, "wb") as fh:
    for i, rid in enumerate(ids):
        fh.write(f">{rid}\n".encode())
        s = seq[int(off[i]):int(off[i + 1])]
        lines = np.full((s.size + 59) // 60 * 61, 10, dtype=np.uint8).reshape(-1, 61)
        pad = np.full(lines.shape[0] * 60, 10, dtype=np.uint8); pad[: s.size] = s
        lines[:, :60] = pad.reshape(-1, 60)
        out = lines.reshape(-1)
        fh.write(out[: s.size + (s.size + 59) // 60].tobytes() if s.size % 60 == 0 else np.concatenate([lines[:-1].reshape(-1), lines[-1][: s.size % 60], [10]]).astype(np.uint8).tobytes())
tidk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "telomeric_identifier_b200", "bin", "tidk")
res = {}
for g in (1, n):
    t0 = time.perf_counter()
    r = subprocess.run([tidk, "search", fa, "-s", "AACCT", "-o", f"g{g}", "-d", "/tmp/cli_gpus_out", "--gpus", str(g), "--gpu-stats"], capture_output=True, text=True)
    res[g] = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr
    print(g, "gpus:", f"{res[g]:.2f} s", r.stderr.strip().splitlines()[-1][:200])
a = open("/tmp/cli_gpus_out/g1_telomeric_repeat_windows.tsv", "rb").read()
b = open(f"/tmp/cli_gpus_out/g{n}_telomeric_repeat_windows.tsv", "rb").read()
print("identical:", a == b, len(a), "bytes")
