"""Host-side pieces of the product (no GPU): canonical key / period / estimator against the
oracle's literal restatement, FASTA parsing, clade lookup, row formatting."""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
from telomeric_identifier_b200 import clades, fasta, hostlogic as H
from telomeric_identifier_b200.search import format_rows

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
DB = os.path.join(HERE, "golden", "clades_min.csv")


def test_golden_vectors_host_logic():
    assert H.reverse_complement(VEC["utils.rs:178-181"]["in"].encode()) == VEC["utils.rs:178-181"]["revcomp"].encode()
    for a, want in VEC["utils.rs:219-227"]["lex_min"]:
        assert H.lex_min(a.encode()) == want.encode()
    for s, p in VEC["explore.rs:477-490"]["periods"]:
        assert H.period(s.encode()) == p
    assert H.estimates([b"AACCT", b"TAAAT", b"AACCT"], [2, 2, 2]) == [(b"AACCT", 4)]


@pytest.mark.parametrize("seed", range(60))
def test_estimator_vs_oracle_literal(seed):
    rng = random.Random(seed)
    alpha = rng.choice(["ACGT", "ACGT", "AC", "ACGTN", "ACGTR"])
    n = rng.randint(0, 40)
    labels, counts = [], []
    pool = []
    for _ in range(rng.randint(1, 6)):
        k = rng.randint(2, 7)
        pool.append("".join(rng.choice(alpha) for _ in range(k)))
    for _ in range(n):
        s = rng.choice(pool)
        if rng.random() < 0.5:
            r = rng.randrange(len(s))
            s = s[r:] + s[:r]
        if rng.random() < 0.3:
            s = H.reverse_complement(s.encode()).decode()
        labels.append(s.encode())
        counts.append(rng.randint(1, 500))
    assert H.estimates(labels, counts) == O.estimates(labels, counts, literal=True)


def test_revcomp_period_lexmin_fuzz():
    rng = random.Random(7)
    for _ in range(300):
        s = "".join(rng.choice("ACGTNacgtRY") for _ in range(rng.randint(1, 14))).encode()
        assert H.reverse_complement(s) == O.revcomp(s)
        assert H.period(s) == O.period(s)
        assert H.lex_min(s) == O.lex_min(s)


def test_fasta_parser(tmp_path):
    p = tmp_path / "x.fa"
    p.write_bytes(b">chr1 desc here\nACGT\nacgtNN  \r\n\n>chr2\n>chr3\tz\nTT\n")
    ids, seq, off = fasta.read_fasta(str(p))
    assert ids == ["chr1", "chr2", "chr3"]
    assert off.tolist() == [0, 10, 10, 12]
    assert seq.tobytes() == b"ACGTacgtNNTT"
    import gzip
    g = tmp_path / "x.fa.gz"
    g.write_bytes(gzip.compress(p.read_bytes()))
    ids2, seq2, off2 = fasta.read_fasta(str(g))
    assert ids2 == ids and (seq2 == seq).all() and (off2 == off).all()
    with pytest.raises(ValueError):
        fasta.parse_fasta(b"ACGT\n>x\nAC\n")


def test_clade_table():
    assert clades.return_telomere_sequence("Lepidoptera", DB) == ["AACCT"]
    hy = clades.return_telomere_sequence("Hymenoptera", DB)
    assert len(hy) == 15 and hy[0] == "AACCCAGACCT" and min(map(len, hy)) == 5 and max(map(len, hy)) == 14
    assert clades.return_telomere_sequence("Coleoptera", DB) == ["AACCC", "AACCT", "AACAGACCCG", "ACCTG"]
    cl = clades.get_clades(DB)
    assert "Lepidoptera" in cl and "" not in cl and len(cl) >= 52


def test_format_rows_matches_reference_layout():
    off = np.array([0, 52, 52, 75], dtype=np.uint64)
    # record 0: 3 windows, record 1: empty, record 2: 2 windows; 2 labels
    f = np.arange(10, dtype=np.uint64)
    r = np.arange(10, dtype=np.uint64) * 10
    txt = format_rows(["a", "b", "c"], off, 20, ["Q1", "Q2"], f, r)
    rows = txt.splitlines()
    assert rows[0] == "a\t20\t0\t0\tQ1" and rows[2] == "a\t52\t2\t20\tQ1" and rows[3] == "a\t20\t3\t30\tQ2"
    assert rows[6] == "c\t20\t6\t60\tQ1" and rows[-1] == "c\t23\t9\t90\tQ2" and len(rows) == 10
    bed = format_rows(["a"], np.array([0, 52], dtype=np.uint64), 20, ["Q"], f[:3], r[:3], "bedgraph").splitlines()
    assert bed == ["a\t0\t20\t0", "a\t20\t40\t11", "a\t40\t52\t22"]


def test_cpp_fasta_reader_matches_python(tmp_path):
    """The C++ driver's multi-threaded FASTA reader against the Python one on a file large enough to
    be split across threads, with ragged line widths, CRLF, blank lines, empty records and a
    header right at a chunk edge."""
    import subprocess
    import __graft_entry__ as g
    g.build()
    tidk = os.path.join(os.path.dirname(HERE), "telomeric_identifier_b200", "bin", "tidk")
    rng = np.random.default_rng(3)
    lines = []
    for r in range(40):
        lines.append(f">rec{r} some description {r}".encode())
        n = int(rng.integers(0, 200_000)) if r % 7 else 0
        s = np.frombuffer(b"ACGTNacgtn", dtype=np.uint8)[rng.integers(0, 10, n)].tobytes()
        w = int(rng.integers(20, 200))
        for j in range(0, n, w):
            tail = [b"", b"\r", b"  ", b" \t\r"][int(rng.integers(0, 4))] if r % 3 == 0 else b""
            lines.append(s[j:j + w] + tail)
        if r % 5 == 0:
            lines.append(b"")
    data = b"\n".join(lines) + (b"\n" if True else b"")
    p = tmp_path / "big.fa"
    p.write_bytes(data)
    ids, seq, off = fasta.read_fasta(str(p))
    h = 1469598103934665603
    want_len = [int(off[i + 1] - off[i]) for i in range(len(ids))]
    for threads in (1, 3, 16):
        out = subprocess.run([tidk, "fasta-check", str(p), "--threads", str(threads)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        rows = out.stdout.splitlines()
        n, total, _ = rows[0].split("\t")
        assert int(n) == len(ids) and int(total) == seq.size
        assert [r.split("\t")[0] for r in rows[1:]] == ids
        assert [int(r.split("\t")[1]) for r in rows[1:]] == want_len
    # same hash for every thread count (content, not only lengths)
    hs = {subprocess.run([tidk, "fasta-check", str(p), "--threads", str(t)], capture_output=True, text=True).stdout.splitlines()[0]
          for t in (1, 2, 5, 16)}
    assert len(hs) == 1
    fnv = 1469598103934665603
    for c in seq[:0].tolist():
        fnv = ((fnv ^ c) * 1099511628211) % (1 << 64)
    # full FNV in Python is slow; check it on a small file instead
    q = tmp_path / "small.fa"
    q.write_bytes(b">a x\nACGT\nAC  \n\n>b\n>c\nTT\r\n")
    out = subprocess.run([tidk, "fasta-check", str(q)], capture_output=True, text=True).stdout.splitlines()
    fnv = 1469598103934665603
    for c in b"ACGTACTT":
        fnv = ((fnv ^ c) * 1099511628211) % (1 << 64)
    assert out[0] == f"3\t8\t{fnv:016x}" and out[1:] == ["a\t6", "b\t0", "c\t2"]
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b"ACGT\n>x\nAC\n")
    assert subprocess.run([tidk, "fasta-check", str(bad)], capture_output=True, text=True).returncode == 1
