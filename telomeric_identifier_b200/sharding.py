"""Multi-GPU sharding of the window counter: windows are independent
(an occurrence never spans two windows, search.rs:98,105-110), so the global window sequence is cut
into `world` contiguous, equally long ranges -- whole records where possible, long records split at
multiples of W -- and every rank counts its range with no halo and no data-path collective.  Per-rank
counts (16 B per window) are gathered on rank 0 and re-assembled into [record][query][window].
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

Piece = Tuple[int, int, int]  # (record, first window, number of windows)


def windows_per_record(offsets: np.ndarray, window: int) -> np.ndarray:
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    return (lens + window - 1) // window


def plan_shards(offsets: np.ndarray, window: int, world: int) -> List[List[Piece]]:
    """Contiguous window ranges of (almost) equal size, one list of pieces per rank."""
    nw = windows_per_record(offsets, window)
    total = int(nw.sum())
    base = np.concatenate([[0], np.cumsum(nw)])
    plans: List[List[Piece]] = []
    for r in range(world):
        lo, hi = total * r // world, total * (r + 1) // world
        pieces: List[Piece] = []
        rec = int(np.searchsorted(base, lo, side="right") - 1) if total else 0
        while lo < hi:
            while base[rec + 1] <= lo:
                rec += 1
            first = lo - int(base[rec])
            n = min(hi, int(base[rec + 1])) - lo
            pieces.append((rec, first, n))
            lo += n
        plans.append(pieces)
    return plans


def shard_batch(seq: np.ndarray, offsets: np.ndarray, pieces: Sequence[Piece], window: int):
    """The byte ranges of `pieces` as a local batch (one sub-record per piece)."""
    parts, lens = [], []
    for rec, first, n in pieces:
        o, e = int(offsets[rec]), int(offsets[rec + 1])
        a = o + first * window
        b = min(o + (first + n) * window, e)
        parts.append(seq[a:b])
        lens.append(b - a)
    loc_off = np.zeros(len(pieces) + 1, dtype=np.uint64)
    if lens:
        loc_off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    loc_seq = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return loc_seq, loc_off


def merge_counts(plans: Sequence[Sequence[Piece]], per_rank: Sequence[Tuple[np.ndarray, np.ndarray]],
                 offsets: np.ndarray, window: int, nq: int):
    """Per-rank [piece][query][window] counts -> global [record][query][window]."""
    nw = windows_per_record(offsets, window)
    base = np.concatenate([[0], np.cumsum(nw)]) * nq
    fwd = np.zeros(int(nw.sum()) * nq, dtype=np.uint64)
    rev = np.zeros_like(fwd)
    for pieces, (f, r) in zip(plans, per_rank):
        p = 0
        for rec, first, n in pieces:
            for q in range(nq):
                dst = int(base[rec]) + q * int(nw[rec]) + first
                fwd[dst:dst + n] = f[p:p + n]
                rev[dst:dst + n] = r[p:p + n]
                p += n
    return fwd, rev


def sharded_window_counts(count_fn: Callable, seq: np.ndarray, offsets: np.ndarray, window: int,
                          queries: Sequence[bytes], rank: int, world: int, dist=None):
    """Count this rank's shard with `count_fn(seq, offsets, window, queries) -> (fwd, rev)` (normally
    Context.window_counts) and gather on rank 0.  Returns the global arrays on rank 0, None elsewhere."""
    plans = plan_shards(offsets, window, world)
    loc_seq, loc_off = shard_batch(seq, offsets, plans[rank], window)
    f, r = count_fn(loc_seq, loc_off, window, queries)
    if world == 1 or dist is None:
        return merge_counts(plans, [(f, r)], offsets, window, len(queries))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((np.asarray(f), np.asarray(r)), gathered, dst=0)
    if rank != 0:
        return None
    return merge_counts(plans, gathered, offsets, window, len(queries))
