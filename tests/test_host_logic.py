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
