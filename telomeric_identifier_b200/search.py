"""Host mirror of search::search (search.rs:10-79): same arguments, same file
content; the per-window counting (search.rs:83-151) runs on the GPU."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

from .api import Context, n_windows

HEADER = "id\twindow\tforward_repeat_number\treverse_repeat_number\ttelomeric_repeat\n"


def format_rows(ids: Sequence[str], offsets: np.ndarray, window: int, labels: Sequence[str],
                fwd: np.ndarray, rev: np.ndarray, extension: str = "tsv") -> str:
    """Rows in the reference order record -> repeat -> window (search.rs:132-147, finder.rs:169-172)."""
    out: List[str] = []
    nw = n_windows(offsets, window)
    p = 0
    for r, rid in enumerate(ids):
        n = int(offsets[r + 1] - offsets[r])
        k = int(nw[r])
        ends = np.minimum((np.arange(k, dtype=np.uint64) + 1) * np.uint64(window), np.uint64(n))
        for lab in labels:
            f, v = fwd[p:p + k], rev[p:p + k]
            if extension == "tsv":
                out.extend(f"{rid}\t{e}\t{a}\t{b}\t{lab}\n" for e, a, b in zip(ends.tolist(), f.tolist(), v.tolist()))
            else:
                starts = (np.arange(k, dtype=np.uint64) * np.uint64(window)).tolist()
                out.extend(f"{rid}\t{s}\t{e}\t{a + b}\n" for s, e, a, b in zip(starts, ends.tolist(), f.tolist(), v.tolist()))
            p += k
    return "".join(out)


def search_rows(ctx: Context, ids: Sequence[str], seq: np.ndarray, offsets: np.ndarray, telomeric_repeat: str,
                window: int = 10000, extension: str = "tsv") -> str:
    """write_window_counts for every record (search.rs:57-70)."""
    fwd_q = telomeric_repeat.upper()  # search.rs:93
    fwd, rev = ctx.window_counts(seq, offsets, window, [fwd_q.encode()])
    return format_rows(ids, offsets, window, [fwd_q], fwd, rev, extension)


def search(ctx: Context, fasta: str, string: str, output: str, dir: str, window: int = 10000,
           extension: str = "tsv") -> str:
    """`tidk search`: writes {dir}/{output}_telomeric_repeat_windows.{extension} (search.rs:35-54)."""
    from .fasta import read_fasta
    ids, seq, offsets = read_fasta(fasta)
    os.makedirs(dir, exist_ok=True)
    path = f"{dir}/{output}_telomeric_repeat_windows.{extension}"
    with open(path, "w") as fh:
        if extension == "tsv":
            fh.write(HEADER)
        fh.write(search_rows(ctx, ids, seq, offsets, string, window, extension))
    return path
