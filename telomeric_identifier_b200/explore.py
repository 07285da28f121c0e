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
    labels = []
    for r in runs:
        o, e = int(offsets[r["record"]]), int(offsets[r["record"] + 1])
        dist = split_dist(e - o, distance)
        base = o if r["slice"] == 0 else e - dist
        p = base + int(r["label_pos"])
        labels.append(seq[p:p + int(r["k"])].tobytes().upper())
    return labels


def explore_table(ctx: Context, seq: np.ndarray, offsets: np.ndarray, length: int = 0, minimum: int = 5,
                  maximum: int = 12, threshold: int = 100, distance: float = 0.01) -> List[Tuple[bytes, int]]:
    if distance > 0.5:
        raise ValueError("Distance from chromosome end as a proportion can't be more than 0.5.")  # explore.rs:41-43
    kmin, kmax = (length, length) if length > 0 else (minimum, maximum)  # explore.rs:50,83-88
    runs = ctx.explore_runs(seq, offsets, distance, kmin, kmax, threshold)
    labels = run_labels(seq, offsets, distance, runs)
    keep = [i for i, l in enumerate(labels) if period(l) >= 3]  # filter_by_frequency, explore.rs:265-274
    counts = [(int(runs[i]["end"]) - int(runs[i]["start"])) // int(runs[i]["k"]) for i in keep]
    return estimates([labels[i] for i in keep], counts)


def explore(ctx: Context, fasta: str, length: int = 0, minimum: int = 5, maximum: int = 12, threshold: int = 100,
            distance: float = 0.01) -> str:
    """`tidk explore`: returns what the reference prints to STDOUT (explore.rs:140-143)."""
    from .fasta import read_fasta
    _, seq, offsets = read_fasta(fasta)
    rows = explore_table(ctx, seq, offsets, length, minimum, maximum, threshold, distance)
    return f"canonical_repeat_unit\tcount_repeat_runs_gt_{threshold}\n" + "".join(
        f"{k.decode()}\t{c}\n" for k, c in rows)
