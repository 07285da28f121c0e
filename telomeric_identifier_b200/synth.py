"""Deterministic synthetic genomes of the five BASELINE.json shapes.

Used by bench.py and the tests: there is no network and no real genome on the GPU
box, and the inputs are too large to ship, so they are generated in place from a
seed.  numpy's Philox bit generator is counter-based, so a (seed, contig) key gives
the same contig no matter how many are generated or in which order.

`scale` shrinks every length (and read count) proportionally for tests.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, stream]))


def _background(rng: np.random.Generator, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m] = _ACGT[rng.integers(0, 4, size=m, dtype=np.uint8)]
    return out


def _array(rng: np.random.Generator, unit: bytes, copies: int, sub: float, indel: float) -> np.ndarray:
    """(unit) x copies with substitutions and 1-bp indels at the given per-base rates."""
    a = np.tile(np.frombuffer(unit, dtype=np.uint8), copies)
    n = a.size
    if sub > 0:
        pos = np.flatnonzero(rng.random(n) < sub)
        a = a.copy()
        a[pos] = _ACGT[rng.integers(0, 4, size=pos.size)]
    if indel > 0:
        ev = np.flatnonzero(rng.random(n) < indel)
        if ev.size:
            keep = np.ones(n, dtype=bool)
            ins_at = []
            for p in ev.tolist():
                if rng.random() < 0.5:
                    keep[p] = False
                else:
                    ins_at.append(p)
            parts, last = [], 0
            for p in ins_at:
                parts.append(a[last:p][keep[last:p]])
                parts.append(_ACGT[rng.integers(0, 4, size=1)])
                last = p
            parts.append(a[last:][keep[last:]])
            a = np.concatenate(parts)
    return a


def _put(seq: np.ndarray, pos: int, arr: np.ndarray) -> None:
    pos = max(0, min(pos, seq.size))
    m = min(arr.size, seq.size - pos)
    seq[pos:pos + m] = arr[:m]


def _put_right(seq: np.ndarray, end: int, arr: np.ndarray) -> None:
    m = min(arr.size, end)
    seq[end - m:end] = arr[arr.size - m:]


def _softmask(rng: np.random.Generator, seq: np.ndarray, frac: float) -> None:
    n = seq.size
    target = int(n * frac)
    done = 0
    while done < target and n > 0:
        ln = int(rng.integers(1000, 5001))
        p = int(rng.integers(0, max(1, n - ln)))
        blk = seq[p:p + ln]
        letters = (blk >= 65) & (blk <= 90)
        blk[letters] |= 0x20
        done += ln


def _finish(ids: List[str], contigs: List[np.ndarray]) -> Tuple[List[str], np.ndarray, np.ndarray]:
    offsets = np.zeros(len(contigs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([c.size for c in contigs], dtype=np.uint64)
    seq = np.concatenate(contigs) if contigs else np.zeros(0, np.uint8)
    return ids, seq, offsets


def config1(scale: float = 1.0, seed: int = 1):
    """C1: 100 Mb / 10 contigs; `search --string TTAGG --window 10000`."""
    ids, contigs = [], []
    for i in range(10):
        rng = _rng(seed, i)
        n = max(64, int((10_000_000 + 1237 * i - 5567) * scale))
        s = _background(rng, n)
        reps = lambda: max(2, int(rng.integers(400, 2401) * min(1.0, scale * 10)))
        _put(s, 0, _array(rng, b"AACCT", reps(), 0.01, 0.002))
        _put_right(s, n, _array(rng, b"TTAGG", reps(), 0.01, 0.002))
        for _ in range(3):
            _put(s, int(rng.integers(0, n)), _array(rng, b"TTAGG", int(rng.integers(20, 61)), 0.01, 0.002))
        _softmask(rng, s, 0.02)
        ids.append(f"contig{i + 1}")
        contigs.append(s)
    return _finish(ids, contigs)


def config2(scale: float = 1.0, seed: int = 2):
    """C2 / C3: 500 Mb / 30 chromosomes; `find --clade Lepidoptera`, `explore 5..12 -t 100 --distance 0.01`."""
    ids, contigs = [], []
    for i in range(30):
        rng = _rng(seed, i)
        n = max(64, int((25_000_000 - i * (25_000_000 - 8_333_333) / 29.0) * scale))
        s = _background(rng, n)
        reps = lambda: max(2, int(rng.integers(400, 2401) * min(1.0, scale * 10)))
        off = 0 if i % 2 == 0 else 137  # exercises the first-run rule both ways (explore.rs:297)
        _put(s, off, _array(rng, b"AACCT", reps(), 0.0005, 0.0001))
        a = _array(rng, b"AGGTT", reps(), 0.0005, 0.0001)
        _put_right(s, n - off, a)
        for _ in range(3):
            _put(s, int(rng.integers(0, n)), _array(rng, b"AACCT", int(rng.integers(20, 61)), 0.0005, 0.0001))
        _softmask(rng, s, 0.02)
        ids.append(f"chr{i + 1}")
        contigs.append(s)
    return _finish(ids, contigs)


_HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
         133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
         58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


_HG38_NAMES = [str(i) for i in range(1, 23)] + ["X", "Y"]


def config4_lengths(scale: float = 1.0):
    return [max(256, int(L * scale)) for L in _HG38]


def config4_contig(i: int, scale: float = 1.0, seed: int = 4):
    """Chromosome i of C4 (keyed by (seed, i): any subset can be generated independently)."""
    rng = _rng(seed, i)
    n = max(256, int(_HG38[i] * scale))
    s = _background(rng, n)
    cap = max(1, int(10_000 * min(1.0, scale * 100)))
    reps = lambda: max(2, int(rng.integers(500, 2501) * min(1.0, scale * 10)))
    _put(s, cap, _array(rng, b"CCCTAA", reps(), 0.001, 0.0002))
    _put_right(s, n - cap, _array(rng, b"TTAGGG", reps(), 0.001, 0.0002))
    s[:cap] = ord("N")
    s[n - cap:] = ord("N")
    cen = int(3_000_000 * scale)
    p = int(rng.integers(n // 4, n // 2))
    s[p:p + cen] = ord("N")
    gap = max(1, int(50_000 * scale))
    for _ in range(20):
        p = int(rng.integers(0, max(1, n - gap)))
        s[p:p + gap] = ord("N")
    return f"chr{_HG38_NAMES[i]}", s


def config4(scale: float = 1.0, seed: int = 4):
    """C4: 3.1 Gb / 24 chromosomes with N gaps; `search --string TTAGGG --window 10000`."""
    ids, contigs = [], []
    for i in range(len(_HG38)):
        rid, s = config4_contig(i, scale, seed)
        ids.append(rid)
        contigs.append(s)
    return _finish(ids, contigs)


def config5(n_reads: int = 2_000_000, seed: int = 5):
    """C5: HiFi-like reads ~N(15 kb, 2 kb); `explore 5..12 -t 100 --distance 0.5`."""
    rng = _rng(seed, 0)
    lens = np.clip(rng.normal(15_000, 2_000, size=n_reads), 5_000, 25_000).astype(np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens).astype(np.uint64)
    seq = _background(rng, int(offsets[-1]))
    telo = np.flatnonzero(rng.random(n_reads) < 0.01)
    for r in telo.tolist():
        o, e = int(offsets[r]), int(offsets[r + 1])
        unit = b"TTAGGG" if rng.random() < 0.5 else b"CCCTAA"
        arr = _array(rng, unit, int(rng.integers(150, 1501)), 0.0005, 0.0)
        view = seq[o:e]
        if rng.random() < 0.5:
            _put(view, 0, arr)
        else:
            _put_right(view, view.size, arr)
    ids = [f"read_{i:07d}" for i in range(n_reads)]
    return ids, seq, offsets


def random_records(seed: int, n_records: int, max_len: int, alphabet: bytes = b"ACGT",
                   plant: bytes = b"", p_empty: float = 0.0):
    """Small ragged batches for parity tests."""
    rng = _rng(seed, 99)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    contigs, ids = [], []
    for r in range(n_records):
        n = 0 if rng.random() < p_empty else int(rng.integers(0, max_len + 1))
        s = alpha[rng.integers(0, alpha.size, size=n)].copy()
        if plant and n > 4 * len(plant):
            for _ in range(int(rng.integers(1, 4))):
                copies = int(rng.integers(2, max(3, n // (3 * len(plant)))))
                _put(s, int(rng.integers(0, n)), np.tile(np.frombuffer(plant, dtype=np.uint8), copies))
        ids.append(f"r{r}")
        contigs.append(s)
    return _finish(ids, contigs)
