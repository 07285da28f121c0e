"""The host packers of the two-ended upload (host_pack.hpp; scalar, AVX2 and AVX-512 forms) against a numpy model: lo = ASCII bit 1,
hi = ASCII bit 2, valid = byte is one of ACGTacgt, 32 bases per word, LSB first; the two-plane variant
reports every word whose validity plane is not all ones as (global word index << 32 | va)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "host_pack_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host_pack_shim.cpp")])
    L = C.CDLL(so)
    L.shim_pack2.restype = C.c_uint64
    L.shim_pack2_strict.restype = C.c_uint64
    return L


def model(buf: np.ndarray, n_words: int):
    pad = np.zeros(n_words * 32, dtype=np.uint8)
    pad[: buf.size] = buf
    exists = np.zeros(n_words * 32, dtype=bool)
    exists[: buf.size] = True
    w = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    lo = (((pad >> 1) & 1).astype(np.uint64) * exists).reshape(n_words, 32) @ w
    hi = (((pad >> 2) & 1).astype(np.uint64) * exists).reshape(n_words, 32) @ w
    up = pad | 0x20
    ok = ((up == ord("a")) | (up == ord("c")) | (up == ord("g")) | (up == ord("t"))) & exists
    va = ok.astype(np.uint64).reshape(n_words, 32) @ w
    return lo.astype(np.uint32), hi.astype(np.uint32), va.astype(np.uint32)


def _aligned(n_words):  # 64-byte aligned, like the pinned staging buffers (the AVX-512 form needs it)
    raw = np.zeros(n_words + 32, dtype=np.uint32)
    shift = (-raw.ctypes.data // 4) % 16
    return raw[shift: shift + n_words]


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("scalar", [0, 1, 2])  # dispatch (AVX-512 if present) / scalar / AVX2
def test_packers_match_model(shim, seed, scalar):
    rng = np.random.default_rng(seed)
    alphabet = [b"ACGT", b"ACGTacgt", b"ACGTN", b"ACGTNRYKMacgtn-*", b"ACGT\xc3", b"N"][seed]
    n = int(rng.integers(1, 5000))
    buf = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()
    if seed == 2:
        buf[100:400] = ord("N")  # an N run: one exception per word
    n_words = (n + 31) // 32 + int(rng.integers(0, 20))      # words beyond the data pack as invalid
    lo, hi, va = model(buf, n_words)
    flag = C.c_uint32(0)
    # three planes
    l3, h3, v3 = _aligned(n_words), _aligned(n_words), _aligned(n_words)
    shim.shim_pack3(buf.ctypes.data, C.c_uint64(n), C.c_uint64(n_words), l3.ctypes.data, h3.ctypes.data, v3.ctypes.data,
                    C.byref(flag), int(scalar == 1))
    assert (l3 == lo).all() and (h3 == hi).all() and (v3 == va).all()
    assert bool(flag.value) == bool((buf >= 0x80).any())
    # two planes + exceptions
    l2, h2 = _aligned(n_words), _aligned(n_words)
    exc = np.zeros(n_words + 1, dtype=np.uint64)
    flag2 = C.c_uint32(0)
    word0 = int(rng.integers(0, 1 << 20))
    ne = shim.shim_pack2(buf.ctypes.data, C.c_uint64(n), C.c_uint64(n_words), l2.ctypes.data, h2.ctypes.data, C.c_uint64(word0),
                         exc.ctypes.data, C.c_uint64(exc.size), C.byref(flag2), scalar)
    assert (l2 == lo).all() and (h2 == hi).all()
    want = [((word0 + i) << 32) | int(v) for i, v in enumerate(va) if v != 0xFFFFFFFF]
    assert ne == len(want) and exc[:ne].tolist() == want      # in word order, nothing else
    assert bool(flag2.value) == bool((buf >= 0x80).any())
    # the device's view: a validity plane of ones with the exceptions scattered over it
    rebuilt = np.full(n_words, 0xFFFFFFFF, dtype=np.uint32)
    for e in exc[:ne].tolist():
        rebuilt[(e >> 32) - word0] = e & 0xFFFFFFFF
    assert (rebuilt == va).all()


def test_empty_and_exact_multiples(shim):
    for n in (0, 32, 64, 256, 257):
        buf = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)[:n].copy()
        n_words = max(1, (n + 31) // 32)
        lo, hi, va = model(buf, n_words)
        l2, h2 = _aligned(n_words), _aligned(n_words)
        exc = np.zeros(n_words + 1, dtype=np.uint64)
        flag = C.c_uint32(0)
        ne = shim.shim_pack2(buf.ctypes.data if n else None, C.c_uint64(n), C.c_uint64(n_words), l2.ctypes.data, h2.ctypes.data,
                             C.c_uint64(0), exc.ctypes.data, C.c_uint64(exc.size), C.byref(flag), 0)
        assert (l2 == lo).all() and (h2 == hi).all()
        assert ne == int((va != 0xFFFFFFFF).sum())


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("scalar", [0, 1, 2])
def test_strict_packers_for_the_explore_upload(shim, seed, scalar):
    """STRICT validity (explore upload): only upper-case A C G T are valid, so an exception-free word is
    reproduced exactly from its lo / hi planes; every other word is reported."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(600, 6000))
    if seed == 0:
        buf = rng.integers(0, 128, n).astype(np.uint8)             # every ASCII byte, spaces and NUL included
    else:
        alphabet = [b"", b"ACGT", b"ACGTacgt", b"ACGTN", b"ACGT -@`\x00\x7f"][seed]
        buf = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()
    buf[64:64 + 512] = np.frombuffer(b"ACGT" * 128, dtype=np.uint8)  # a clean stretch: words without exceptions
    n_words = (n + 31) // 32 + 3
    lo, hi, _ = model(buf, n_words)
    pad = np.zeros(n_words * 32, dtype=np.uint8); pad[:n] = buf
    ok = np.isin(pad, np.frombuffer(b"ACGT", dtype=np.uint8)); ok[n:] = False
    w = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    va = (ok.astype(np.uint64).reshape(n_words, 32) @ w).astype(np.uint32)
    l2, h2 = _aligned(n_words), _aligned(n_words)
    exc = np.zeros(n_words + 1, dtype=np.uint64)
    flag = C.c_uint32(0)
    ne = shim.shim_pack2_strict(buf.ctypes.data, C.c_uint64(n), C.c_uint64(n_words), l2.ctypes.data, h2.ctypes.data, C.c_uint64(7),
                                exc.ctypes.data, C.c_uint64(exc.size), C.byref(flag), scalar)
    assert (l2 == lo).all() and (h2 == hi).all()
    want = [((7 + i) << 32) | int(v) for i, v in enumerate(va) if v != 0xFFFFFFFF]
    assert ne == len(want) and exc[:ne].tolist() == want
    # what the device does with it: unpack (hi, lo) -> A C T G and overwrite the exception words with the raw bytes
    lut = np.frombuffer(b"ACTG", dtype=np.uint8)
    bits_lo = ((l2[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)
    bits_hi = ((h2[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)
    out = lut[bits_hi * 2 + bits_lo]
    for e in exc[:ne].tolist():
        i = (e >> 32) - 7
        out[i * 32:(i + 1) * 32] = pad[i * 32:(i + 1) * 32]
    assert (out[:n] == buf).all()
