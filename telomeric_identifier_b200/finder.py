"""Host mirror of finder::finder (finder.rs:13-107): the clade's repeats are
looked up in the table (clades.rs:119-141) and counted verbatim (finder.rs:129-136);
rows are ordered record -> repeat -> window (finder.rs:121,144)."""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from .api import Context
from .clades import return_telomere_sequence
from .search import HEADER, format_rows


def finder_rows(ctx: Context, ids: Sequence[str], seq: np.ndarray, offsets: np.ndarray,
                repeats: Sequence[str], window: int = 10000) -> str:
    fwd, rev = ctx.window_counts(seq, offsets, window, [r.encode() for r in repeats])
    return format_rows(ids, offsets, window, list(repeats), fwd, rev, "tsv")


def finder(ctx: Context, fasta: str, clade: str, output: str, dir: str, window: int = 10000,
           database: Optional[str] = None) -> str:
    """`tidk find`: writes {dir}/{output}_telomeric_repeat_windows.tsv (finder.rs:66-78)."""
    from .fasta import read_fasta
    repeats = return_telomere_sequence(clade, database)
    ids, seq, offsets = read_fasta(fasta)
    os.makedirs(dir, exist_ok=True)
    path = f"{dir}/{output}_telomeric_repeat_windows.tsv"
    with open(path, "w") as fh:
        fh.write(HEADER)
        fh.write(finder_rows(ctx, ids, seq, offsets, repeats, window))
    return path
