"""Clade table lookup -- stays on the host (clades.rs:96-141).  The table is the
CSV that `tidk build` stores at <data_dir>/tidk/tidk_database.csv (build.rs:30-40);
columns used: "Order" and "Telomeric repeat"."""
from __future__ import annotations

import csv
import os
from typing import List, Optional


def database_path() -> str:
    base = os.environ.get("XDG_DATA_HOME") or os.path.join(os.path.expanduser("~"), ".local", "share")
    return os.path.join(base, "tidk", "tidk_database.csv")


def _rows(path: Optional[str]):
    with open(path or database_path(), newline="") as fh:
        for row in csv.DictReader(fh):
            yield row


def get_clades(path: Optional[str] = None) -> List[str]:
    """clades.rs:96-115: the Order column, consecutive duplicates and empties removed."""
    out: List[str] = []
    for row in _rows(path):
        o = row["Order"]
        if not out or out[-1] != o:
            out.append(o)
    return [o for o in out if o]


def return_telomere_sequence(clade: str, path: Optional[str] = None) -> List[str]:
    """clades.rs:119-141: repeats of every row with Order == clade, first occurrence order."""
    seqs: List[str] = []
    for row in _rows(path):
        if row["Order"] == clade and row["Telomeric repeat"] not in seqs:
            seqs.append(row["Telomeric repeat"])
    return seqs
