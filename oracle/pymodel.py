"""Second, independent restatement of the reference hot path in pure Python.

TEST INFRASTRUCTURE ONLY.  Written line by line after the Rust (citations are
reference file:line) with Python strings/lists standing in for String/Vec, so
that the C oracle (tidk_oracle.c) can be fuzzed against a model that shares no
code with it.  Small inputs only.
"""
from __future__ import annotations

import math
from itertools import combinations
from typing import Dict, List, Optional, Tuple


# utils.rs:49-58
def switch_base(c: str) -> str:
    return {"A": "T", "C": "G", "T": "A", "G": "C", "N": "N"}.get(c, "N")


# utils.rs:37-46
def reverse_complement(dna: str) -> str:
    return "".join(switch_base(c) for c in dna)[::-1]


# utils.rs:19-34 (KMP / BOM both return every occurrence, overlapping included)
def find_motifs(motif: str, string: str) -> List[int]:
    out, i = [], string.find(motif)
    while i != -1:
        out.append(i)
        i = string.find(motif, i + 1)
    return out


# utils.rs:63-80 -- restated literally to show it is the identity
def remove_overlapping_indexes(indexes: List[int], pattern_length: int) -> List[int]:
    indexes = list(indexes)
    index = 0
    while True:
        vec_len = len(indexes)
        if not indexes or index == 0 or index == vec_len - 1:
            break
        while indexes[index + 1] < indexes[index] + pattern_length:  # pragma: no cover (unreachable)
            del indexes[index + 1]
            index += 1
    return indexes


# search.rs:83-151
def search_rows(seq: str, rid: str, telomeric_repeat: str, window_size: int, extension: str = "tsv") -> List[str]:
    fwd = telomeric_repeat.upper()
    rev = reverse_complement(fwd).upper()
    rows = []
    start, end = 0, window_size
    for i in range(0, math.ceil(len(seq) / window_size)):
        window = seq[i * window_size:(i + 1) * window_size].upper()
        f = len(remove_overlapping_indexes(find_motifs(fwd, window), len(fwd)))
        r = len(remove_overlapping_indexes(find_motifs(rev, window), len(fwd)))
        if i != 0:
            start += window_size
            end += window_size
        if end > len(seq):
            end = len(seq)
        if extension == "tsv":
            rows.append(f"{rid}\t{end}\t{f}\t{r}\t{fwd}")
        else:
            rows.append(f"{rid}\t{start}\t{end}\t{f + r}")
    return rows


# finder.rs:111-178
def find_rows(seq: str, rid: str, repeats: List[str], window_size: int) -> List[str]:
    rows = []
    for fwd in repeats:
        rev = reverse_complement(fwd)
        end = window_size
        for i in range(0, math.ceil(len(seq) / window_size)):
            window = seq[i * window_size:(i + 1) * window_size].upper()
            f = len(find_motifs(fwd, window))
            r = len(find_motifs(rev, window))
            if i != 0:
                end += window_size
            if end > len(seq):
                end = len(seq)
            rows.append(f"{rid}\t{end}\t{f}\t{r}\t{fwd}")
    return rows


# utils.rs:87-94
def string_rotation(s1: str, s2: str) -> bool:
    return len(s1) == len(s2) and s1 in (s2 + s2)


# utils.rs:100-127
def minimal_rotation(s: str) -> int:
    n = len(s)
    i, j = 0, 1
    while True:
        k = 0
        ci, cj = s[i % n], s[j % n]
        while k < n:
            ci, cj = s[(i + k) % n], s[(j + k) % n]
            if ci != cj:
                break
            k += 1
        if k == n:
            return min(i, j)
        if ci > cj:
            i += k + 1
            i += int(i == j)
        else:
            j += k + 1
            j += int(i == j)


# utils.rs:147-166
def lex_min(dna: str) -> str:
    r = reverse_complement(dna)
    jf, jr = minimal_rotation(dna), minimal_rotation(r)
    return min(dna[jf:] + dna[:jf], r[jr:] + r[:jr])


# utils.rs:133-141
def lms(r1: str, r2: str) -> str:
    assert (string_rotation(r1, r2) or string_rotation(r1, reverse_complement(r2))
            or string_rotation(reverse_complement(r1), r2))
    return lex_min(r1)


# explore.rs:438-445
def check_repeats(s: str) -> int:
    delays: Dict[int, int] = {}  # delay -> iterator position
    for i, c in enumerate(s):
        keep = {}
        for d, it in delays.items():
            if it < len(s) and s[it] == c:
                keep[d] = it + 1
        delays = keep
        delays[i + 1] = 0
    return min(delays)


# explore.rs:151-160
def split_seq_by_distance(seq: str, d: float) -> Tuple[str, str]:
    dist = int(math.ceil(len(seq) * d))
    return seq[0:dist], seq[len(seq) - dist:]


# explore.rs:178-233
def chunk_fasta(sequence: str, k: int) -> List[Tuple[int, str]]:
    if len(sequence) <= k:
        return []
    chunks = [sequence[i:i + k] for i in range(0, len(sequence), k)]
    indexes: List[Tuple[int, str]] = []
    pos, first = 0, True
    for a, b in zip(chunks, chunks[1:]):
        if a == b:
            if first:
                indexes.append((pos, a.upper()))
                pos += k
                indexes.append((pos, a.upper()))
                first = False
            else:
                pos += k
                indexes.append((pos, a.upper()))
        else:
            pos += k
            first = True
    return indexes


# explore.rs:289-347 (+ filter_by_frequency :265-274)
def calculate_indexes(indexes: List[Tuple[int, str]], k: int, frequency: int,
                      period_filter: bool = True) -> Optional[List[Tuple[int, int, str]]]:
    start = 0
    collection = []
    pairs = list(zip(indexes, indexes[1:]))
    for n, ((p1, s1), (p2, s2)) in enumerate(pairs):
        if n == len(pairs) - 1:
            collection.append((start, p2 + k, s1))
        elif s1 == s2 and (p2 - p1) == len(s1):
            continue
        else:
            collection.append((start, p1 + k, s1))
            start = p2
    if not collection:
        return None
    return [(s, e, q) for (s, e, q) in collection
            if (e - s) // len(q) > frequency and (not period_filter or not check_repeats(q) < 3)]


# explore.rs:359-371
def test_repeats(r1: str, r2: str) -> Tuple[bool, Optional[str]]:
    eq = (string_rotation(r1, r2) or string_rotation(reverse_complement(r1), r2)
          or string_rotation(r1, reverse_complement(r2)))
    return (eq, lms(r1, r2)) if eq else (eq, None)


# explore.rs:377-431 (+ filter_count_vec :455-464); ties sorted by key for determinism
def get_telomeric_repeat_estimates(runs: List[Tuple[int, int, str]]) -> List[Tuple[str, int]]:
    groups: Dict[int, List[Tuple[int, int, str]]] = {}
    for r in runs:
        groups.setdefault(len(r[2]), []).append(r)
    count = lambda r: (r[1] - r[0]) // len(r[2])
    m: Dict[str, int] = {}
    for g in groups.values():
        if len(g) == 1:
            m[lex_min(g[0][2])] = count(g[0])
            continue
        tracker: List[int] = []
        for a, b in combinations(range(len(g)), 2):
            ok, seq = test_repeats(g[a][2], g[b][2])
            if ok:
                if seq not in m:
                    m[seq] = count(g[a]) + count(g[b])
                if a not in tracker and b not in tracker:
                    m[seq] += count(g[a]) + count(g[b])
            tracker += [a, b]
    rows = sorted(m.items(), key=lambda kv: (-kv[1], kv[0]))
    return [(k, v) for k, v in rows if check_repeats(k) > 3]


# explore.rs:83-137 in FASTA file order
def explore(seqs: List[str], distance: float, kmin: int, kmax: int, threshold: int) -> List[Tuple[str, int]]:
    out: List[Tuple[int, int, str]] = []
    for k in range(kmin, kmax + 1):
        for s in seqs:
            for sl in split_seq_by_distance(s, distance):
                r = calculate_indexes(chunk_fasta(sl, k), k, threshold)
                if r:
                    out += r
    return get_telomeric_repeat_estimates(out)
