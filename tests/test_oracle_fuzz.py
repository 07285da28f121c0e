"""C oracle vs the independent pure-Python model on seeded random inputs that
stress what the reference's own tests do not pin: mixed case, N runs, planted
arrays, overlapping self-matches, odd windows, thresholds > 0."""
import random

import pytest

from oracle import oracle as O
from oracle import pymodel as M


def rand_seq(rng, n, alphabet="ACGT", plant=None):
    s = [rng.choice(alphabet) for _ in range(n)]
    if plant and n > 4 * len(plant):
        for _ in range(rng.randint(1, 3)):
            reps = rng.randint(2, max(2, n // (3 * len(plant))))
            p = rng.randint(0, n - 1)
            arr = (plant * reps)[: n - p]
            s[p:p + len(arr)] = list(arr)
    return "".join(s)


@pytest.mark.parametrize("seed", range(40))
def test_search_find_vs_model(seed):
    rng = random.Random(seed)
    alphabet = rng.choice(["ACGT", "ACGTacgtN", "ACGTNnacgtRYKM-*", "AT", "A"])
    q = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 9)))
    if seed % 7 == 0:
        q = q[0] * len(q)  # self-overlapping motif
    if seed % 11 == 0:
        q = q + "N"
    n = rng.randint(0, 400)
    seq = rand_seq(rng, n, alphabet, plant=q)
    w = rng.choice([1, 2, 3, 7, 20, 33, 64, 100, 1000])
    got = O.search_text(["r"], [seq.encode()], q.encode(), w).splitlines()[1:]
    assert got == M.search_rows(seq, "r", q, w)
    got = O.search_text(["r"], [seq.encode()], q.lower().encode(), w, "bedgraph").splitlines()
    assert got == M.search_rows(seq, "r", q.lower(), w, "bedgraph")
    reps = [q, M.reverse_complement(q), q.lower(), "AACCT"]
    got = O.find_text(["r"], [seq.encode()], [r.encode() for r in reps], w).splitlines()[1:]
    assert got == M.find_rows(seq, "r", reps, w)


@pytest.mark.parametrize("seed", range(40))
def test_explore_vs_model(seed):
    rng = random.Random(1000 + seed)
    alphabet = rng.choice(["ACGT", "ACGTacgt", "AC", "ACGTN", "Aa"])
    seqs = []
    for _ in range(rng.randint(1, 4)):
        unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(2, 8)))
        seqs.append(rand_seq(rng, rng.randint(0, 500), alphabet, plant=unit))
    d = rng.choice([0.5, 0.1, 0.01, 0.33, 0.25])
    thr = rng.choice([0, 0, 1, 2, 5])
    kmin = rng.randint(1, 6)
    kmax = kmin + rng.randint(0, 4)
    want = M.explore(seqs, d, kmin, kmax, thr)
    text = O.explore_text([s.encode() for s in seqs], d, kmin, kmax, thr, literal=True)
    got = [tuple(r.split("\t")) for r in text.splitlines()[1:]]
    assert got == [(k, str(c)) for k, c in want]
    # the O(n) estimator agrees with the literal pairwise loop on A/C/G/T/N labels
    if not any(c in a for a in alphabet for c in "RYKM"):
        assert O.explore_text([s.encode() for s in seqs], d, kmin, kmax, thr, literal=False) == text
    # threads only change who computes what, not the arrival order
    assert O.explore_text([s.encode() for s in seqs], d, kmin, kmax, thr, threads=3) == text


@pytest.mark.parametrize("seed", range(20))
def test_slice_runs_vs_model(seed):
    rng = random.Random(5000 + seed)
    unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 6)))
    s = rand_seq(rng, rng.randint(0, 300), rng.choice(["ACGT", "ACGTacgt", "AC"]), plant=unit)
    k = rng.randint(1, 7)
    thr = rng.choice([0, 1, 3])
    want = M.calculate_indexes(M.chunk_fasta(s, k), k, thr) or []
    got = [(a, b, l.decode()) for a, b, l, _ in O.slice_runs(s.encode(), k, thr)]
    assert got == want
    assert [(p, l.decode()) for p, l in O.chunk_fasta(s.encode(), k)] == M.chunk_fasta(s, k)
