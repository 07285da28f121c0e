"""Host-side pieces of the path that stay on the CPU by design:
they touch a handful of k-byte labels per run, not the sequence.

  reverse_complement   utils.rs:37-58
  repeat period        explore.rs:438-445 (check_repeats)
  canonical key        utils.rs:147-166 (lex_min: min rotation of s and of revcomp(s))
  estimator            explore.rs:377-431 + filter_count_vec explore.rs:455-464
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

_COMP = bytes.maketrans(b"ACGT", b"TGCA")
_NONACGT = bytes(c for c in range(256) if c not in b"ACGT")
_TO_N = bytes.maketrans(_NONACGT, b"N" * len(_NONACGT))


def reverse_complement(s: bytes) -> bytes:
    """A<->T, C<->G, everything else (lower case included) -> 'N'; reversed."""
    return s.translate(_TO_N).translate(_COMP)[::-1]


def ascii_upper(s: bytes) -> bytes:
    return s.upper()


def period(s: bytes) -> int:
    """Shortest d >= 1 with s[j] == s[j+d] for all j < len(s)-d."""
    n = len(s)
    for d in range(1, n):
        if s[d:] == s[:n - d]:
            return d
    return n


def min_rotation(s: bytes) -> bytes:
    n = len(s)
    if n == 0:
        return s
    d = s + s
    return min(d[i:i + n] for i in range(n))


def lex_min(s: bytes) -> bytes:
    return min(min_rotation(s), min_rotation(reverse_complement(s)))


def _is_rotation(a: bytes, b: bytes) -> bool:
    return len(a) == len(b) and a in (b + b)


def _same_class(a: bytes, b: bytes) -> bool:  # explore.rs:359-371
    return (_is_rotation(a, b) or _is_rotation(reverse_complement(a), b)
            or _is_rotation(a, reverse_complement(b)))


_INVOLUTIVE = frozenset(b"ACGTN")


def estimates(labels: Sequence[bytes], counts: Sequence[int]) -> List[Tuple[bytes, int]]:
    """Runs in arrival order -> [(canonical unit, count)], sorted (count desc, key asc).

    The reference walks all pairs of a length group (explore.rs:395-423).  Its net
    effect per rotation/revcomp class is closed-form: the first two members
    (m1 < m2, arrival order) contribute count[m1] + count[m2], doubled iff
    (m1, m2) == (0, 1); classes with a single member in a group of >= 2 vanish; a
    group of one inserts its own count.  That O(n) form is used whenever every
    label is over A/C/G/T/N (revcomp is an involution there, so the pairwise
    relation is an equivalence); other alphabets take the literal pairwise loop.
    """
    groups: Dict[int, List[int]] = {}
    for i, l in enumerate(labels):
        groups.setdefault(len(l), []).append(i)
    out: Dict[bytes, int] = {}
    for k, idx in groups.items():
        if len(idx) == 1:
            out[lex_min(labels[idx[0]])] = int(counts[idx[0]])
            continue
        if all(set(labels[i]) <= _INVOLUTIVE for i in idx):
            first: Dict[bytes, Tuple[int, int]] = {}  # key -> (group index of m1, count)
            done = set()
            canon_cache: Dict[bytes, bytes] = {}
            for g, i in enumerate(idx):
                lab = labels[i]
                key = canon_cache.get(lab)
                if key is None:
                    key = canon_cache[lab] = lex_min(lab)
                if key in done:
                    continue
                if key not in first:
                    first[key] = (g, int(counts[i]))
                else:
                    g1, c1 = first[key]
                    v = c1 + int(counts[i])
                    if g1 == 0 and g == 1:
                        v *= 2
                    out[key] = v
                    done.add(key)
        else:
            tracker = set()
            for a in range(len(idx)):
                for b in range(a + 1, len(idx)):
                    la, lb = labels[idx[a]], labels[idx[b]]
                    if _same_class(la, lb):
                        key = lex_min(la)
                        add = int(counts[idx[a]]) + int(counts[idx[b]])
                        if key not in out:
                            out[key] = add
                        if a not in tracker and b not in tracker:
                            out[key] += add
                    tracker.add(a)
                    tracker.add(b)
    rows = sorted(out.items(), key=lambda kv: (-kv[1], kv[0]))
    return [(k, v) for k, v in rows if period(k) > 3]
