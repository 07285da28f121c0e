"""FASTA ingest -- stays on the host, as in the reference (lib.rs:32-49 opens plain
or .gz by extension and hands the stream to bio::io::fasta::Reader).

Semantics follow rust-bio 2.0.3's reader as far as the reference relies on them
(unpinned by the reference's own tests): a record starts at a line
beginning with '>', id = first whitespace-delimited token of that line, sequence
= the following lines with trailing whitespace removed, concatenated.
"""
from __future__ import annotations

import gzip
from typing import List, Tuple

import numpy as np


def read_fasta(path: str) -> Tuple[List[str], np.ndarray, np.ndarray]:
    """-> (ids, concatenated sequence bytes, offsets[n+1])."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fh:
        data = fh.read()
    return parse_fasta(data)


def parse_fasta(data: bytes) -> Tuple[List[str], np.ndarray, np.ndarray]:
    ids: List[str] = []
    parts: List[bytes] = []
    lens: List[int] = []
    cur: List[bytes] = []
    started = False
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            if started:
                s = b"".join(cur)
                parts.append(s)
                lens.append(len(s))
            head = line[1:].rstrip().split(None, 1)
            ids.append(head[0].decode() if head else "")
            cur = []
            started = True
        elif started:
            cur.append(line.rstrip())
        elif line.strip():
            raise ValueError("Expected > at record start.")
    if started:
        s = b"".join(cur)
        parts.append(s)
        lens.append(len(s))
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    if lens:
        offsets[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if parts else np.zeros(0, np.uint8)
    return ids, seq, offsets
