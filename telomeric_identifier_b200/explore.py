"""Host mirror of explore::explore (explore.rs:20-149).  The GPU returns every run
above the threshold in FASTA arrival order (tidk_cuda_explore_runs); the period
filter, canonical key and estimator run here on the run labels."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .api import Context
from .hostlogic import estimates, period


def split_dist(n: int, d: float) -> int:  # explore.rs:156
    v = math.ceil(float(n) * d)
    return int(v) if v > 0 else 0


def run_labels(seq: np.ndarray, offsets: np.ndarray, distance: float, runs: np.ndarray) -> List[bytes]:
    """Upper-cased label of every run (slice[label_pos .. label_pos + k)), gathered per k in one indexing step:
    a read set yields hundreds of thousands of runs, a per-run Python loop would dominate the call."""
    if runs.size == 0:
        return []
    offsets = np.asarray(offsets)
    rec = runs["record"].astype(np.int64)
    o = offsets[rec].astype(np.int64)
    e = offsets[rec + 1].astype(np.int64)
    dist = np.maximum(np.ceil((e - o).astype(np.float64) * distance), 0).astype(np.int64)  # explore.rs:156
    p = np.where(runs["slice"] == 0, o, e - dist) + runs["label_pos"].astype(np.int64)
    out = np.empty(runs.size, dtype=object)
    for k in np.unique(runs["k"]).tolist():
        sel = np.flatnonzero(runs["k"] == k)
        m = seq[p[sel][:, None] + np.arange(k, dtype=np.int64)[None, :]]
        m = np.where((m >= 97) & (m <= 122), m - 32, m).astype(np.uint8)
        out[sel] = np.ascontiguousarray(m).view(f"S{k}").ravel().tolist()
    labels = out.tolist()
    # numpy's S dtype strips trailing NULs; a label can only end in NUL if the sequence holds NUL bytes
    return [l if len(l) == int(kk) else l + b"\x00" * (int(kk) - len(l)) for l, kk in zip(labels, runs["k"].tolist())]


def explore_table(ctx: Context, seq: np.ndarray, offsets: np.ndarray, length: int = 0, minimum: int = 5,
                  maximum: int = 12, threshold: int = 100, distance: float = 0.01) -> List[Tuple[bytes, int]]:
    if distance > 0.5:
        raise ValueError("Distance from chromosome end as a proportion can't be more than 0.5.")  # explore.rs:41-43
    kmin, kmax = (length, length) if length > 0 else (minimum, maximum)  # explore.rs:50,83-88
    runs = ctx.explore_runs(seq, offsets, distance, kmin, kmax, threshold)
    labels = run_labels(seq, offsets, distance, runs)
    ok = {l: period(l) >= 3 for l in set(labels)}  # filter_by_frequency, explore.rs:265-274 (once per distinct label)
    keep = [i for i, l in enumerate(labels) if ok[l]]
    counts = ((runs["end"] - runs["start"]) // runs["k"]).tolist()
    return estimates([labels[i] for i in keep], [counts[i] for i in keep])


def explore(ctx: Context, fasta: str, length: int = 0, minimum: int = 5, maximum: int = 12, threshold: int = 100,
            distance: float = 0.01) -> str:
    """`tidk explore`: returns what the reference prints to STDOUT (explore.rs:140-143)."""
    from .fasta import read_fasta
    _, seq, offsets = read_fasta(fasta)
    rows = explore_table(ctx, seq, offsets, length, minimum, maximum, threshold, distance)
    return f"canonical_repeat_unit\tcount_repeat_runs_gt_{threshold}\n" + "".join(
        f"{k.decode()}\t{c}\n" for k, c in rows)
