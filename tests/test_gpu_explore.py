"""Parity of the CUDA explore path (through the C ABI) with the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

GENOME = b"AACCTAACCTAACATATCGTAACCTAACCTAACCTAACCTAACATATCGTAACCTAACCT"
GENOME_2 = b"AACCTAACCTTAAATTAAATAACCTAACCTAACCTAACCTTAAATTAAATAACCTAACCT"


@pytest.fixture(scope="module")
def ctx():
    import telomeric_identifier_b200 as T
    c = T.Context(0)
    yield c
    c.close()


def _runs_equal(ctx, seq, off, d, kmin, kmax, thr):
    from telomeric_identifier_b200.explore import run_labels
    got = ctx.explore_runs(seq, off, d, kmin, kmax, thr)
    want, wlabels = O.explore_runs(seq, off, d, kmin, kmax, thr, period_filter=False)
    assert got.size == want.size, (got.size, want.size)
    for f in ("record", "slice", "k", "first", "start", "end"):
        assert (got[f] == want[f]).all(), f
    assert run_labels(seq, off, d, got) == wlabels
    # the same batch resident: bit-planes built at upload; once on the plane kernel (small batches stay there), once with
    # the vertical scan forced (explore_vertical.cuh: plain tiles in vertical form, the rest handed to the plane kernel)
    import os
    batch = ctx.upload(np.ascontiguousarray(seq, dtype=np.uint8), np.ascontiguousarray(off, dtype=np.uint64))
    try:
        for min_tiles in ("1000000000", "0"):
            os.environ["TIDK_VERTICAL_MIN_TILES"] = min_tiles
            res = ctx.explore_runs_resident(batch, d, kmin, kmax, thr)
            assert res.size == want.size, (min_tiles, res.size, want.size)
            for f in ("record", "slice", "k", "first", "start", "end", "label_pos"):
                assert (res[f] == got[f]).all(), (min_tiles, f)
    finally:
        os.environ.pop("TIDK_VERTICAL_MIN_TILES", None)
        batch.free()


def test_reference_golden_index_left(ctx):  # explore.rs:592-612
    seq = np.frombuffer(GENOME, dtype=np.uint8)
    runs = ctx.explore_runs(seq, np.array([0, seq.size], dtype=np.uint64), 0.5, 5, 5, 0)
    left = runs[runs["slice"] == 0]
    assert [(int(a), int(b)) for a, b in zip(left["start"], left["end"])] == [(0, 10), (20, 30)]
    right = runs[runs["slice"] == 1]
    assert [(int(a), int(b)) for a, b in zip(right["start"], right["end"])] == [(0, 10), (20, 30)]


def test_reference_golden_estimates(ctx):  # explore.rs:622-626 (left half of GENOME_2 alone)
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(GENOME_2[:30], dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    runs = ctx.explore_runs(seq, off, 0.5, 5, 5, 0)
    # d = 0.5 on a 30-mer gives two 15-base slices; use the explicit one-slice form instead
    from telomeric_identifier_b200.hostlogic import estimates
    seq60 = np.frombuffer(GENOME_2, dtype=np.uint8)
    runs = ctx.explore_runs(seq60, np.array([0, 60], dtype=np.uint64), 0.5, 5, 5, 0)
    left = runs[runs["slice"] == 0]
    labels = [GENOME_2[int(p):int(p) + 5] for p in left["label_pos"]]
    counts = [(int(e) - int(s)) // 5 for s, e in zip(left["start"], left["end"])]
    assert labels == [b"AACCT", b"TAAAT", b"AACCT"]
    assert estimates(labels, counts) == [(b"AACCT", 4)]


@pytest.mark.parametrize("seed", range(16))
def test_random_runs(ctx, seed):
    from telomeric_identifier_b200 import synth
    rng = np.random.default_rng(seed)
    alphabet = [b"ACGT", b"ACGTacgt", b"AC", b"ACGTN", b"Aa", b"ACGTNRYacgtn"][seed % 6]
    unit = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(2, 9))).tolist())
    ids, seq, off = synth.random_records(100 + seed, int(rng.integers(1, 6)), 200000 if seed % 4 == 0 else 5000,
                                         alphabet, plant=unit, p_empty=0.15)
    d = [0.5, 0.1, 0.01, 0.33][seed % 4]
    thr = [0, 0, 1, 3, 100][seed % 5]
    kmin = int(rng.integers(1, 7))
    _runs_equal(ctx, seq, off, d, kmin, kmin + int(rng.integers(0, 6)), thr)


@pytest.mark.parametrize("seed", range(8))
def test_explore_table_vs_oracle(ctx, seed):
    import telomeric_identifier_b200 as T
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.random_records(300 + seed, 4, 20000, b"ACGTacgt" if seed % 2 else b"ACGT",
                                         plant=[b"TTAGGG", b"AACCT", b"TTTAGGG", b"AACCCTG"][seed % 4])
    thr = [0, 2, 10, 50][seed % 4]
    rows = T.explore_table(ctx, seq, off, 0, 4, 9, thr, 0.5)
    seqs = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ids))]
    want = O.explore_text(seqs, 0.5, 4, 9, thr, literal=(thr >= 2)).splitlines()[1:]
    assert [f"{k.decode()}\t{c}" for k, c in rows] == want


def test_config3_scaled(ctx):
    import telomeric_identifier_b200 as T
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.config2(scale=0.02)
    _runs_equal(ctx, seq, off, 0.01, 5, 12, 100)
    _runs_equal(ctx, seq, off, 0.05, 5, 12, 20)
    rows = T.explore_table(ctx, seq, off, 0, 5, 12, 20, 0.05)
    seqs = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ids))]
    want = O.explore_text(seqs, 0.05, 5, 12, 20, literal=True).splitlines()[1:]
    assert [f"{k.decode()}\t{c}" for k, c in rows] == want
    assert any(k == b"AACCT" for k, _ in rows)


def test_config5_scaled(ctx):
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.config5(n_reads=3000)
    _runs_equal(ctx, seq, off, 0.5, 5, 12, 100)


def test_config5_in_batches(ctx):
    """C5's batched launch path: a read set handed over in batches of records.  A read's runs do not depend on
    the other reads and every call returns (k, record, slice, start) order, so per k the batches' run lists
    concatenated in batch order are the run list of the whole set (explore.rs:88-121 walks the records once
    per k); the table from the merged runs is the oracle's table."""
    from telomeric_identifier_b200 import synth
    from telomeric_identifier_b200.explore import run_labels
    from telomeric_identifier_b200.hostlogic import estimates, period
    ids, seq, off = synth.config5(n_reads=1500)
    want, wlabels = O.explore_runs(seq, off, 0.5, 5, 12, 100, period_filter=False)
    parts = []
    for r0, r1 in ((0, 400), (400, 401), (401, 1100), (1100, 1500)):
        lo = off[r0:r1 + 1] - off[r0]
        got = ctx.explore_runs(seq[int(off[r0]):int(off[r1])], lo, 0.5, 5, 12, 100)
        got["record"] += np.uint32(r0)
        parts.append(got)
    merged = np.concatenate(parts)
    merged = merged[np.argsort(merged["k"], kind="stable")]
    assert merged.size == want.size
    for f in ("record", "slice", "k", "first", "start", "end"):
        assert (merged[f] == want[f]).all(), f
    labels = run_labels(seq, off, 0.5, merged)
    assert labels == wlabels
    keep = [i for i, l in enumerate(labels) if period(l) >= 3]
    counts = [(int(merged[i]["end"]) - int(merged[i]["start"])) // int(merged[i]["k"]) for i in keep]
    rows = estimates([labels[i] for i in keep], counts)
    seqs = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ids))]
    text = O.explore_text(seqs, 0.5, 5, 12, 100, literal=False).splitlines()[1:]
    assert [f"{k.decode()}\t{c}" for k, c in rows] == text


def test_explore_errors(ctx):
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(b"ACGT" * 50, dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    with pytest.raises(T.TidkCudaError) as e:
        ctx.explore_runs(seq, off, 0.6, 5, 12, 100)
    assert e.value.code == -1
    with pytest.raises(T.TidkCudaError):
        ctx.explore_runs(seq, off, 0.1, 0, 3, 100)
    assert ctx.explore_runs(np.zeros(0, np.uint8), np.array([0], dtype=np.uint64), 0.1, 5, 12, 100).size == 0


def test_large_k_uses_bytewise_kernel(ctx):
    """k > 16 is outside the plane kernel: the byte-wise kernel must give the same answer."""
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.random_records(77, 3, 60000, b"ACGTacgtN", plant=b"TTAGGGTTAGGGTTAGGCA")
    _runs_equal(ctx, seq, off, 0.5, 15, 20, 0)
    _runs_equal(ctx, seq, off, 0.25, 19, 19, 2)


def test_tile_and_block_boundaries(ctx):
    """Arrays, N runs and soft-masked stretches placed across 1 KiB block and 16 KiB tile edges."""
    rng = np.random.default_rng(5)
    n = 200_000
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
    for pos in (1000, 16380, 32700, 49152 - 7, 65536 + 3, 99990):
        unit = np.frombuffer(b"TTAGGG", dtype=np.uint8)
        arr = np.tile(unit, 40)
        s[pos:pos + arr.size] = arr
    s[20000:23000] = ord("N")           # N run: equal chunks of N are runs too (and can be the first run)
    s[40000:40100] |= 0x20              # soft-masked stretch inside random sequence
    s[65530:65600] |= 0x20              # case change inside an array
    s[150000:150012] = np.frombuffer(b"acgtacACGTAC", dtype=np.uint8)  # equal after upper-casing only
    off = np.array([0, n], dtype=np.uint64)
    for d, k0, k1, thr in ((0.5, 1, 16, 0), (0.5, 5, 12, 3), (0.3, 6, 6, 0), (0.085, 2, 9, 1)):
        _runs_equal(ctx, s, off, d, k0, k1, thr)


def test_many_tiny_records(ctx):
    """Records shorter than k, than 2k, than one block; offsets at every alignment."""
    rng = np.random.default_rng(9)
    parts = []
    for i in range(400):
        ln = int(rng.integers(0, 90))
        unit = np.frombuffer(b"ACGTT"[: int(rng.integers(1, 6))], dtype=np.uint8)
        parts.append(np.tile(unit, ln // unit.size + 1)[:ln] if i % 3 else np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, ln)])
    off = np.zeros(len(parts) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in parts])
    seq = np.concatenate(parts).astype(np.uint8)
    _runs_equal(ctx, seq, off, 0.5, 1, 8, 0)
    _runs_equal(ctx, seq, off, 0.2, 3, 5, 1)


def test_host_and_device_run_building_agree(ctx):
    """Runs are normally built on the device (sorted 64-bit event keys); batches that do not fit the key
    layout fall back to struct events + host stitching.  Force the fallback and compare."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, telomeric_identifier_b200 as T\n"
        "from telomeric_identifier_b200 import synth\n"
        "from oracle import oracle as O\n"
        "ids, seq, off = synth.random_records(55, 5, 80000, b'ACGTacgtN', plant=b'TTAGGG')\n"
        "with T.Context(0) as ctx:\n"
        "    got = ctx.explore_runs(seq, off, 0.5, 2, 9, 1)\n"
        "    b = ctx.upload(seq, off)\n"
        "    res = ctx.explore_runs_resident(b, 0.5, 2, 9, 1)\n"
        "want, _ = O.explore_runs(seq, off, 0.5, 2, 9, 1, period_filter=False)\n"
        "for g in (got, res):\n"
        "    assert g.size == want.size and all((g[f] == want[f]).all() for f in ('record','slice','k','first','start','end'))\n"
        "print('ok', got.size)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # ... and (TIDK_EXPLORE_POST_MAX) the path of batches with more events than CUB's 32-bit counts take: packed keys,
    # stitched on the host -- with the vertical scan forced, whose events carry no type bits
    for env_extra in ({}, {"TIDK_EXPLORE_HOST_POST": "1"}, {"TIDK_EXPLORE_POST_MAX": "10", "TIDK_VERTICAL_MIN_TILES": "0"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root,
                           env={**os.environ, **env_extra})
        assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr[-2000:]


def test_dense_events_overflow_and_retry(ctx):
    """Worst case for the event list: 'AACC' repeated makes a raw-run boundary at almost every chunk for
    k = 1 and k = 2, far more than the first-guess buffer (scanned / 64) holds -- the call must notice
    the overflow, re-run with the exact size and still agree with the oracle."""
    n = 3_000_000
    seq = np.frombuffer((b"AACC" * (n // 4 + 1))[:n], dtype=np.uint8).copy()
    seq[1_000_000:1_000_400] = ord("G")  # one long run in the middle
    off = np.array([0, n], dtype=np.uint64)
    _runs_equal(ctx, seq, off, 0.5, 1, 3, 0)
    _runs_equal(ctx, seq, off, 0.5, 1, 2, 50)


@pytest.mark.parametrize("case", ["upper", "softmask", "n_caps"])
def test_slice_ends_around_block_and_tile_edges(ctx, case):
    """Slices whose length falls just below / at / just above multiples of the 1 KiB block, the 4 KiB and
    the 16 KiB tile (the slice can end in the last block of an item or in the one before it), at every
    start alignment, with arrays that touch the slice start, the slice end (closing event beyond it) or
    cover the slice whole.  Exercises the one-ballot plain path next to the edge blocks."""
    rng = np.random.default_rng(31)
    parts = []
    unit6, unit5 = np.frombuffer(b"TTAGGG", dtype=np.uint8), np.frombuffer(b"AACCT", dtype=np.uint8)
    for base in (1024, 2048, 4096, 8192, 16384, 32768):
        for delta in (-33, -12, -7, -1, 0, 1, 5, 11, 12, 13, 31, 32, 40):
            dist = base + delta
            n = 2 * dist - int(rng.integers(0, 2))                   # ceil(n / 2) == dist
            s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
            mode = int(rng.integers(0, 4))
            if mode == 0:    # arrays at both record ends: first run of the left slice, last run of the right one
                a = np.tile(unit6, 60)
                s[:a.size] = a
                s[n - a.size:] = a
            elif mode == 1:  # array across the middle: ends the left slice, starts the right one
                a = np.tile(unit5, 200)
                p0 = max(0, n // 2 - 500)
                s[p0:p0 + a.size] = a[: n - p0]
            elif mode == 2:  # one array over the whole record
                s[:] = np.tile(unit6, n // 6 + 1)[:n]
            if case == "softmask":
                p0 = int(rng.integers(0, max(1, n - 300)))
                s[p0:p0 + 300] |= 0x20
            elif case == "n_caps":
                s[:37] = ord("N")
                s[n - 41:] = ord("N")
            parts.append(s)
            parts.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(rng.integers(0, 40)))])  # shifts the alignment
    off = np.zeros(len(parts) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in parts])
    seq = np.concatenate(parts).astype(np.uint8)
    _runs_equal(ctx, seq, off, 0.5, 5, 12, 0)
    _runs_equal(ctx, seq, off, 0.5, 1, 16, 2)
    _runs_equal(ctx, seq, off, 0.5, 5, 12, 100)


# ---- two-ended (host-packed) upload of the explore entry point ---------------------------------------
def _packed_case(ctx, monkeypatch, seq, off, d, pinned, hybrid):
    """The host-pointer call with the STRICT-packed upload forced on must return the run list of the resident
    call (plain ASCII upload) and of the oracle."""
    monkeypatch.setenv("TIDK_EXPLORE_PACK", "1")
    monkeypatch.setenv("TIDK_PACK_THREADS", "4")
    monkeypatch.setenv("TIDK_HYBRID_UPLOAD", "1" if hybrid else "0")
    src = seq
    if pinned:
        src = ctx.pinned_empty(seq.size)
        src[:] = seq
    got = ctx.explore_runs(src, off, d, 5, 12, 3)
    h2d, _, packed = ctx.last_transfer()
    assert packed == 1
    monkeypatch.setenv("TIDK_EXPLORE_PACK", "0")
    plain = ctx.explore_runs(seq, off, d, 5, 12, 3)
    assert ctx.last_transfer()[2] == 0
    assert got.size == plain.size and (got == plain).all()
    want, _ = O.explore_runs(seq, off, d, 5, 12, 3, period_filter=False)
    assert got.size == want.size
    for f in ("record", "slice", "k", "first", "start", "end"):
        assert (got[f] == want[f]).all(), f
    return h2d


def _reads(rng, n_reads, alphabet, mean=6000):
    parts, off = [], [0]
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    for _ in range(n_reads):
        n = int(rng.integers(mean // 2, mean * 2))
        r = lut[rng.integers(0, lut.size, n)].copy()
        if rng.random() < 0.3:  # a planted array at one end
            unit = np.frombuffer([b"TTAGGG", b"AACCT", b"TTAGGCATG"][int(rng.integers(0, 3))], dtype=np.uint8)
            m = int(rng.integers(50, 600))
            arr = np.tile(unit, m // unit.size + 1)[:m]
            if rng.random() < 0.5: r[:m] = arr
            else: r[n - m:] = arr
        parts.append(r)
        off.append(off[-1] + n)
    return np.concatenate(parts), np.array(off, dtype=np.uint64)


@pytest.mark.parametrize("pinned,hybrid", [(False, False), (True, True), (True, False)])
def test_packed_upload_clean_reads(ctx, monkeypatch, pinned, hybrid):
    """Upper-case A/C/G/T reads over several 4 MiB chunks (the last one partial): everything packs; the bus
    carries about a quarter of the bases when nothing goes up as ASCII."""
    rng = np.random.default_rng(11)
    seq, off = _reads(rng, 1500, b"ACGT")
    assert seq.size > 2 * (4 << 20)
    h2d = _packed_case(ctx, monkeypatch, seq, off, 0.5, pinned, hybrid)
    if not hybrid:
        assert h2d < 0.5 * seq.size


@pytest.mark.parametrize("seed", range(4))
def test_packed_upload_exceptions_and_fallback_chunks(ctx, monkeypatch, seed):
    """Sparse exceptions (N, lower case, IUPAC codes, a space) travel as raw words; chunks dense in them go up as
    ASCII; slices straddle chunk borders."""
    rng = np.random.default_rng(200 + seed)
    seq, off = _reads(rng, 900, b"ACGT", mean=8000)
    n = seq.size
    # sparse single-byte exceptions in the first part
    for ch in b"NnacgtRY -":
        pos = rng.integers(0, n // 2, 200)
        seq[pos] = ch
    # a soft-masked stretch and an N gap across the second chunk border
    b = 2 * (4 << 20)
    if n > b + 300000:
        seq[b - 200000:b + 100000] |= 0x20
        seq[b + 150000:b + 250000] = ord("N")
    if seed % 2:  # mixed-case tandem array over a chunk border
        unit = np.frombuffer(b"TTAGGGttagggTTAGGG", dtype=np.uint8)
        b1 = 4 << 20
        seq[b1 - 3000:b1 + 3000] = np.tile(unit, 6000 // unit.size + 1)[:6000]
    _packed_case(ctx, monkeypatch, seq, off, [0.5, 0.45, 0.5, 0.41][seed], bool(seed & 1), bool(seed & 2))


def test_packed_upload_small_and_nonascii(ctx, monkeypatch):
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(GENOME * 3, dtype=np.uint8).copy()
    off = np.array([0, 60, 180], dtype=np.uint64)
    monkeypatch.setenv("TIDK_EXPLORE_PACK", "1")
    monkeypatch.setenv("TIDK_PACK_THREADS", "2")
    got = ctx.explore_runs(seq, off, 0.5, 5, 6, 0)
    want, _ = O.explore_runs(seq, off, 0.5, 5, 6, 0, period_filter=False)
    assert got.size == want.size and (got["start"] == want["start"]).all() and (got["end"] == want["end"]).all()
    bad = seq.copy()
    bad[70] = 0xC3
    with pytest.raises(T.TidkCudaError) as e:
        ctx.explore_runs(bad, off, 0.5, 5, 6, 0)
    assert e.value.code == -4  # TIDK_ENONASCII


# ---- vertical scan (explore_vertical.cuh): resident batches of upper-case A C G T ---------------------------------
@pytest.mark.parametrize("seed", range(4))
def test_vertical_heads_tails_and_tile_edges(ctx, seed):
    """Records whose slices start with 0..40 A's (the head flip), end inside a repeat (the closing parity), and whose
    lengths sit around the tile boundaries of the vertical scan (1024, 1024 + 960 j), for every k the scan supports."""
    rng = np.random.default_rng(900 + seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lens = [1, 7, 31, 47, 64, 65, 959, 960, 1023, 1024, 1025, 1087, 1088, 1983, 1984, 1985, 2944, 2945, 3904, 4000, 7777]
    recs = []
    for i in range(160):
        n = 2 * lens[(i + seed) % len(lens)] + int(rng.integers(0, 2))
        r = rng.choice(acgt, size=n)
        na = (i * 7 + seed) % 41
        r[:na] = ord("A")
        r[n - (n // 2): n - (n // 2) + (na // 2)] = ord("A")      # the right slice starts with A's too
        if i % 3 == 0:  # a repeat that runs into the end of the record (and of the left slice)
            k = int(rng.integers(1, 17))
            unit = rng.choice(acgt, size=k)
            ln = min(n, int(rng.integers(k, 40 * k)))
            r[n - ln:] = np.tile(unit, ln // k + 1)[:ln]
            h = (n + 1) // 2
            l2 = min(h, int(rng.integers(k, 40 * k)))
            r[h - l2: h] = np.tile(unit, l2 // k + 1)[:l2]
        recs.append(r)
    seq = np.concatenate(recs)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in recs])
    kmin = [1, 5, 9, 13][seed]
    _runs_equal(ctx, seq, off, 0.5, kmin, kmin + 3, [0, 1, 2, 0][seed])
    _runs_equal(ctx, seq, off, 0.5, 5, 12, 3)


def test_vertical_mixed_plain_and_other_tiles(ctx):
    """Soft-masked stretches, N gaps and IUPAC codes inside otherwise plain records: the tiles that hold them go to the
    plane kernel, their neighbours stay on the vertical scan, and the seams do not show."""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for i in range(40):
        n = int(rng.integers(3000, 30000))
        r = rng.choice(acgt, size=n)
        unit = rng.choice(acgt, size=int(rng.integers(2, 13)))
        p = int(rng.integers(0, n - 2000))
        r[p:p + 1500] = np.tile(unit, 1500 // unit.size + 1)[:1500]
        for _ in range(3):
            q = int(rng.integers(0, n - 300))
            ln = int(rng.integers(1, 300))
            what = int(rng.integers(0, 3))
            if what == 0:
                r[q:q + ln] |= 0x20  # lower case
            elif what == 1:
                r[q:q + ln] = ord("N")
            else:
                r[q] = ord("R")
        recs.append(r)
    seq = np.concatenate(recs)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in recs])
    _runs_equal(ctx, seq, off, 0.5, 5, 12, 10)
    _runs_equal(ctx, seq, off, 0.37, 2, 7, 0)
