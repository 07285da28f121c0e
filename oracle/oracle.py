"""ctypes front-end of the CPU ORACLE (oracle/tidk_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product never imports it.

The text-producing helpers restate the reference's own row formats:
  search TSV / bedgraph  search.rs:49-54,132-147
  find TSV               finder.rs:75-78,169-172
  explore table          explore.rs:140-143
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile the oracle with gcc (idempotent)."""
    so = os.path.join(_HERE, "libtidk_oracle.so")
    src = os.path.join(_HERE, "tidk_oracle.c")
    hdr = os.path.join(_HERE, "tidk_oracle.h")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libtidk_oracle.so"])
    return so


class Run(C.Structure):
    _fields_ = [("record", C.c_uint32), ("slice", C.c_uint8), ("k", C.c_uint8),
                ("first_in_slice", C.c_uint8), ("pad", C.c_uint8),
                ("start", C.c_uint64), ("end", C.c_uint64), ("label_pos", C.c_uint64)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p = C.POINTER(C.c_uint8)
        u64p = C.POINTER(C.c_uint64)
        L.tidk_oracle_revcomp.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.tidk_oracle_revcomp.restype = None
        L.tidk_oracle_find_motifs.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, u64p, C.c_size_t]
        L.tidk_oracle_find_motifs.restype = C.c_size_t
        L.tidk_oracle_window_counts.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint32,
                                                C.c_int, C.c_void_p, C.c_void_p]
        L.tidk_oracle_window_counts.restype = C.c_int64
        L.tidk_oracle_window_counts_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                                      C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint32,
                                                      C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.tidk_oracle_window_counts_batch.restype = C.c_int
        L.tidk_oracle_period.argtypes = [C.c_char_p, C.c_size_t]
        L.tidk_oracle_period.restype = C.c_size_t
        L.tidk_oracle_string_rotation.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.tidk_oracle_string_rotation.restype = C.c_int
        L.tidk_oracle_lex_min.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.tidk_oracle_lex_min.restype = None
        L.tidk_oracle_split_dist.argtypes = [C.c_uint64, C.c_double]
        L.tidk_oracle_split_dist.restype = C.c_uint64
        L.tidk_oracle_chunk_fasta.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, u64p, u64p, C.c_size_t]
        L.tidk_oracle_chunk_fasta.restype = C.c_int64
        L.tidk_oracle_slice_runs.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32,
                                             C.c_uint8, C.POINTER(C.POINTER(Run)), u64p, u64p]
        L.tidk_oracle_slice_runs.restype = C.c_int
        L.tidk_oracle_explore_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_uint32,
                                               C.c_uint32, C.c_uint64, C.c_int, C.c_int,
                                               C.POINTER(C.POINTER(Run)), u64p]
        L.tidk_oracle_explore_runs.restype = C.c_int
        L.tidk_oracle_estimates.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.POINTER(C.c_int64),
                                            C.c_uint64, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_int64), C.c_size_t]
        L.tidk_oracle_estimates.restype = C.c_int64
        L._libc = C.CDLL(None)
        L._libc.free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error {code}")
        self.code = code


# ------------------------------------------------------------------ utils
def revcomp(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s))
    lib().tidk_oracle_revcomp(s, len(s), out)
    return out.raw


def find_motifs(motif: bytes, text: bytes) -> List[int]:
    n = lib().tidk_oracle_find_motifs(motif, len(motif), text, len(text), None, 0)
    buf = (C.c_uint64 * max(n, 1))()
    lib().tidk_oracle_find_motifs(motif, len(motif), text, len(text), buf, n)
    return list(buf[:n])


def period(s: bytes) -> int:
    return lib().tidk_oracle_period(s, len(s))


def string_rotation(a: bytes, b: bytes) -> bool:
    return bool(lib().tidk_oracle_string_rotation(a, len(a), b, len(b)))


def lex_min(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s))
    lib().tidk_oracle_lex_min(s, len(s), out)
    return out.raw


def split_dist(n: int, d: float) -> int:
    return lib().tidk_oracle_split_dist(n, d)


# ---------------------------------------------------------- search / find
def concat_records(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offsets[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    cat = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return cat, offsets


def n_windows(offsets: np.ndarray, window: int) -> np.ndarray:
    lens = (offsets[1:] - offsets[:-1]).astype(np.uint64)
    return ((lens + np.uint64(window - 1)) // np.uint64(window)).astype(np.uint64)


def window_counts(seq: np.ndarray, offsets: np.ndarray, window: int, queries: Sequence[bytes],
                  uppercase_query: bool, threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Counts in the C-ABI layout [record][query][window]."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if window <= 0:
        raise OracleError(-1)
    total = int(n_windows(offsets, window).sum()) * len(queries)
    fwd = np.zeros(max(total, 1), dtype=np.uint64)
    rev = np.zeros(max(total, 1), dtype=np.uint64)
    qarr = (C.c_char_p * len(queries))(*queries)
    qlen = (C.c_uint32 * len(queries))(*[len(q) for q in queries])
    rc = lib().tidk_oracle_window_counts_batch(seq.ctypes.data, offsets.ctypes.data, len(offsets) - 1, window,
                                               qarr, qlen, len(queries), int(uppercase_query),
                                               fwd.ctypes.data, rev.ctypes.data, threads)
    if rc != 0:
        raise OracleError(rc)
    return fwd[:total], rev[:total]


def _upper(b: bytes) -> bytes:
    return bytes(c - 32 if 97 <= c <= 122 else c for c in b)


SEARCH_HEADER = "id\twindow\tforward_repeat_number\treverse_repeat_number\ttelomeric_repeat\n"


def search_text(ids: Sequence[str], seqs: Sequence[bytes], query: bytes, window: int,
                extension: str = "tsv") -> str:
    """Full file content written by search::search (search.rs:49-54, 57-70, 132-147)."""
    cat, off = concat_records(seqs)
    fwd, rev = window_counts(cat, off, window, [query], True)
    label = _upper(query).decode()
    out = [SEARCH_HEADER] if extension == "tsv" else []
    p = 0
    for rid, s in zip(ids, seqs):
        n = len(s)
        nw = (n + window - 1) // window
        for i in range(nw):
            end = min((i + 1) * window, n)
            if extension == "tsv":
                out.append(f"{rid}\t{end}\t{fwd[p]}\t{rev[p]}\t{label}\n")
            else:
                out.append(f"{rid}\t{i * window}\t{end}\t{fwd[p] + rev[p]}\n")
            p += 1
    return "".join(out)


def find_text(ids: Sequence[str], seqs: Sequence[bytes], repeats: Sequence[bytes], window: int) -> str:
    """Full file content written by finder::finder (finder.rs:75-78, 85-98, 169-172):
    rows ordered record -> repeat -> window."""
    cat, off = concat_records(seqs)
    fwd, rev = window_counts(cat, off, window, list(repeats), False)
    out = [SEARCH_HEADER]
    p = 0
    for rid, s in zip(ids, seqs):
        n = len(s)
        nw = (n + window - 1) // window
        for r in repeats:
            for i in range(nw):
                end = min((i + 1) * window, n)
                out.append(f"{rid}\t{end}\t{fwd[p]}\t{rev[p]}\t{r.decode()}\n")
                p += 1
    return "".join(out)


# --------------------------------------------------------------- explore
def chunk_fasta(slice_: bytes, k: int) -> List[Tuple[int, bytes]]:
    n = lib().tidk_oracle_chunk_fasta(slice_, len(slice_), k, None, None, 0)
    if n < 0:
        raise OracleError(n)
    pos = (C.c_uint64 * max(n, 1))()
    lab = (C.c_uint64 * max(n, 1))()
    lib().tidk_oracle_chunk_fasta(slice_, len(slice_), k, pos, lab, n)
    return [(pos[i], _upper(slice_[lab[i]:lab[i] + k])) for i in range(n)]


def slice_runs(slice_: bytes, k: int, threshold: int, period_filter: bool = True):
    """calculate_indexes on one slice -> [(start, end, label, first_in_slice)]"""
    runs = C.POINTER(Run)()
    n = C.c_uint64(0)
    cap = C.c_uint64(0)
    rc = lib().tidk_oracle_slice_runs(slice_, len(slice_), k, threshold, int(period_filter), 0, 0,
                                      C.byref(runs), C.byref(n), C.byref(cap))
    if rc:
        raise OracleError(rc)
    out = [(runs[i].start, runs[i].end, _upper(slice_[runs[i].label_pos:runs[i].label_pos + k]),
            runs[i].first_in_slice) for i in range(n.value)]
    if n.value or cap.value:
        lib()._libc.free(runs)
    return out


def explore_runs(seq: np.ndarray, offsets: np.ndarray, distance: float, kmin: int, kmax: int,
                 threshold: int, period_filter: bool = True, threads: int = 1):
    """All retained runs in reference arrival order (k, record, slice, position).
    Returns a numpy structured array plus the list of upper-cased labels."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    runs = C.POINTER(Run)()
    n = C.c_uint64(0)
    rc = lib().tidk_oracle_explore_runs(seq.ctypes.data, offsets.ctypes.data, len(offsets) - 1, distance,
                                        kmin, kmax, threshold, int(period_filter), threads,
                                        C.byref(runs), C.byref(n))
    if rc:
        raise OracleError(rc)
    dt = np.dtype([("record", "<u4"), ("slice", "u1"), ("k", "u1"), ("first", "u1"), ("pad", "u1"),
                   ("start", "<u8"), ("end", "<u8"), ("label_pos", "<u8")])
    if n.value:
        arr = np.ctypeslib.as_array(C.cast(runs, C.POINTER(C.c_uint8)), shape=(n.value * C.sizeof(Run),))
        arr = arr.view(dt).copy()
        lib()._libc.free(runs)
    else:
        arr = np.zeros(0, dtype=dt)
    return arr, run_labels(seq, offsets, distance, arr)


def run_labels(seq: np.ndarray, offsets: np.ndarray, distance: float, runs: np.ndarray) -> List[bytes]:
    labels = []
    for r in runs:
        o, e = int(offsets[r["record"]]), int(offsets[r["record"] + 1])
        n = e - o
        dist = split_dist(n, distance)
        base = o if r["slice"] == 0 else e - dist
        p = base + int(r["label_pos"])
        labels.append(_upper(seq[p:p + int(r["k"])].tobytes()))
    return labels


def estimates(labels: Sequence[bytes], counts: Sequence[int], literal: bool = True) -> List[Tuple[bytes, int]]:
    n = len(labels)
    lp = (C.c_char_p * max(n, 1))(*labels)
    kl = (C.c_uint32 * max(n, 1))(*[len(l) for l in labels])
    ct = (C.c_int64 * max(n, 1))(*[int(c) for c in counts])
    keycap = sum(len(l) for l in labels) + 1
    keys = C.create_string_buffer(keycap)
    oklen = (C.c_uint32 * max(n, 1))()
    ocnt = (C.c_int64 * max(n, 1))()
    m = lib().tidk_oracle_estimates(lp, kl, ct, n, int(literal), keys, keycap, oklen, ocnt, max(n, 1))
    if m < 0:
        raise OracleError(m)
    out, off = [], 0
    for i in range(m):
        out.append((keys.raw[off:off + oklen[i]], ocnt[i]))
        off += oklen[i]
    return out


def explore_text(seqs: Sequence[bytes], distance: float, kmin: int, kmax: int, threshold: int,
                 literal: bool = True, threads: int = 1) -> str:
    """STDOUT of explore::explore (explore.rs:140-143), rows sorted (count desc, key asc)."""
    cat, off = concat_records(seqs)
    runs, labels = explore_runs(cat, off, distance, kmin, kmax, threshold, True, threads)
    counts = [(int(r["end"]) - int(r["start"])) // int(r["k"]) for r in runs]
    rows = estimates(labels, counts, literal)
    out = [f"canonical_repeat_unit\tcount_repeat_runs_gt_{threshold}\n"]
    out += [f"{k.decode()}\t{c}\n" for k, c in rows]
    return "".join(out)
