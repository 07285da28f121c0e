"""The arithmetic of the vertical explore scan (csrc/explore_vertical_core.hpp), compiled for the host: the events a lane
computes tile by tile (transposed planes, sliding ANDs, chunk-end masks, head flip, closing parity) against the events
of the definition (explore.rs:178-233), on slices whose lengths sit around every tile boundary."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("vshim") / "vertical_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "vertical_shim.cpp")])
    L = C.CDLL(so)
    for f in (L.vt_events, L.vt_events_ref):
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
    L.vt_transpose_ok.argtypes = [C.c_void_p]
    return L


def events(L, fn, seq, beg, m, kmin, kmax):
    out = np.zeros(1 << 16, dtype=np.uint64)
    n = fn(seq.ctypes.data, seq.size, beg, m, kmin, kmax, out.ctypes.data, out.size)
    assert n <= out.size
    return np.sort(out[:n])


def make(rng, n, repeats=True):
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    if repeats:
        for _ in range(int(rng.integers(0, 6))):
            k = int(rng.integers(1, 17))
            unit = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=k)
            p = int(rng.integers(0, max(1, n - 1)))
            ln = int(rng.integers(1, 700))
            arr = np.tile(unit, ln // k + 2)[:ln]
            seq[p:p + ln] = arr[: max(0, min(ln, n - p))]
    return seq


def test_transpose(shim):
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = rng.integers(0, 2**32, size=32, dtype=np.uint64).astype(np.uint32)
        assert shim.vt_transpose_ok(x.ctypes.data) == 1


@pytest.mark.parametrize("seed", range(6))
def test_slices_around_tile_boundaries(shim, seed):
    rng = np.random.default_rng(100 + seed)
    lens = [0, 1, 2, 5, 9, 10, 11, 23, 24, 25, 31, 32, 33, 63, 64, 65, 959, 960, 961, 1023, 1024, 1025, 1087, 1088, 1089,
            1983, 1984, 1985, 2047, 2048, 2049, 2943, 2944, 2945, 4001, 7500]
    for m in lens:
        pre, post = int(rng.integers(0, 70)), int(rng.integers(0, 70))
        seq = make(rng, pre + m + post)
        got = events(shim, shim.vt_events, seq, pre, m, 1, 16)
        want = events(shim, shim.vt_events_ref, seq, pre, m, 1, 16)
        assert got.size == want.size and (got == want).all(), (m, pre)


def test_all_a_heads_and_periodic_tails(shim):
    rng = np.random.default_rng(7)
    for m in (40, 700, 1024, 1500, 2500):
        for unit in (b"A", b"AC", b"AAC", b"TTAGGG", b"AACCT", b"ACGTACGTACGA"):
            body = np.tile(np.frombuffer(unit, dtype=np.uint8), m // len(unit) + 2)[:m]
            for pre_fill in (b"A", b"C", unit):
                pre = np.tile(np.frombuffer(pre_fill, dtype=np.uint8), 40)[:37]
                post = np.tile(np.frombuffer(unit, dtype=np.uint8), 40)[:45]  # the repeat continues behind the slice
                seq = np.concatenate([pre, body, post])
                for cut in (0, 1, 3):
                    got = events(shim, shim.vt_events, seq, pre.size, m - cut, 1, 16)
                    want = events(shim, shim.vt_events_ref, seq, pre.size, m - cut, 1, 16)
                    assert got.size == want.size and (got == want).all(), (m, unit, pre_fill, cut)
    # random sequence starting with runs of A of every length
    for na in range(0, 40):
        seq = make(rng, 1500, repeats=False)
        seq[5:5 + na] = ord("A")
        got = events(shim, shim.vt_events, seq, 5, 1400, 1, 16)
        want = events(shim, shim.vt_events_ref, seq, 5, 1400, 1, 16)
        assert (got.size == want.size) and (got == want).all(), na
