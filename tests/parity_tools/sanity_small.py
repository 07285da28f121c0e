"""Tiny end-to-end run for compute-sanitizer: one window-count call per kernel family and one explore call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import telomeric_identifier_b200 as T
from telomeric_identifier_b200 import synth
from oracle import oracle as O
ids, seq, off = synth.random_records(5, 3, 20000, b"ACGTNacgt", plant=b"TTAGGG")
with T.Context(0) as ctx:
    for qs, W in (([b"TTAGGG"], 1000), ([b"GATTACA"], 777), ([b"AACCC", b"AACCT", b"AACAGACCCG", b"ACCTG"], 5000),
                  ([b"TTANG", b"ACGTT", b"GGGTTA"], 40000)):
        f, r = ctx.window_counts(seq, off, W, qs)
        of, orr = O.window_counts(seq, off, W, qs, False)
        assert (f == of).all() and (r == orr).all()
    runs = ctx.explore_runs(seq, off, 0.5, 4, 9, 1)
    want, _ = O.explore_runs(seq, off, 0.5, 4, 9, 1, period_filter=False)
    assert runs.size == want.size and (runs["end"] == want["end"]).all()
    runs = ctx.explore_runs(seq, off, 0.2, 17, 18, 0)
print("sanity ok")
