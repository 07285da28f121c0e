"""Parity of the CUDA window counter (through the C ABI) with the CPU oracle.
Bit-exact: every output is an integer count."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import telomeric_identifier_b200 as T
    c = T.Context(0)
    yield c
    c.close()


def _cmp(ctx, seq, off, W, queries, upper=False):
    qs = [q.upper() for q in queries] if upper else list(queries)
    f, r = ctx.window_counts(seq, off, W, qs)
    of, orr = O.window_counts(seq, off, W, qs, False)
    assert f.shape == of.shape
    bad = np.flatnonzero((f != of) | (r != orr))
    assert bad.size == 0, f"W={W} q={qs}: first mismatch at {bad[:5]} got {f[bad[:5]]},{r[bad[:5]]} want {of[bad[:5]]},{orr[bad[:5]]}"


def test_reference_golden_search(ctx):  # search.rs:183-199
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(b"TTAGGTTAGGTTAGGCAGCATCACACTGATCATCTGATTAGGTTAGGTTAGG", dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    rows = T.search_rows(ctx, ["test1"], seq, off, "TTAGG", 20).splitlines()
    assert rows == ["test1\t20\t3\t0\tTTAGG", "test1\t40\t0\t0\tTTAGG", "test1\t52\t2\t0\tTTAGG"]
    bed = T.search_rows(ctx, ["test1"], seq, off, "ttagg", 20, "bedgraph").splitlines()
    assert bed == ["test1\t0\t20\t3", "test1\t20\t40\t0", "test1\t40\t52\t2"]


def test_reference_golden_finder(ctx):  # finder.rs:213-235
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(b"AAACCCTAAACCCTAAACCCTTGAGAGAGGGGGTGTGGGGAGGGGTTGAGAAACCCT", dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    rows = T.finder_rows(ctx, ["test1"], seq, off, ["AAACCCT"], 20).splitlines()
    assert rows == ["test1\t20\t2\t0\tAAACCCT", "test1\t40\t0\t0\tAAACCCT", "test1\t57\t1\t0\tAAACCCT"]


def test_reference_golden_motifs1(ctx):  # utils.rs:231-238, as a one-window count
    hay = np.frombuffer(b"AACCTAACCTAACCTAACCTAACCTAACCTAACTAACCT", dtype=np.uint8)
    f, r = ctx.window_counts(hay, np.array([0, hay.size], dtype=np.uint64), 1000, [b"AACCT"])
    assert (int(f[0]), int(r[0])) == (7, 0)


@pytest.mark.parametrize("seed", range(24))
def test_random_batches(ctx, seed):
    from telomeric_identifier_b200 import synth
    rng = np.random.default_rng(seed)
    alphabet = [b"ACGT", b"ACGTacgt", b"ACGTNacgtn", b"ACGTNRYKMSWBDHVacgtn-*.", b"AT", b"A"][seed % 6]
    L = int(rng.integers(1, 13))
    q = bytes(rng.choice(list(b"ACGT"), size=L).tolist())
    if seed % 5 == 0:
        q = q[:1] * L  # self-overlapping
    ids, seq, off = synth.random_records(seed, int(rng.integers(1, 6)), 30000, alphabet, plant=q, p_empty=0.2)
    for W in [int(rng.integers(1, 40)), int(rng.integers(40, 2000)), 10000, int(rng.integers(33000, 80000))]:
        if W < 8 and seq.size > 20000:
            continue
        _cmp(ctx, seq, off, W, [q])


@pytest.mark.parametrize("W", [1, 2, 31, 32, 33, 1000, 1024, 1056, 10000, 32768, 32769, 65537, 200000])
def test_window_alignments(ctx, W):
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.random_records(7, 4, 150000, b"ACGTacgtN", plant=b"TTAGGG")
    if W < 4:
        seq, off = seq[:4000].copy(), np.array([0, 1500, 1500, 4000], dtype=np.uint64)
    _cmp(ctx, seq, off, W, [b"TTAGGG"])
    _cmp(ctx, seq, off, W, [b"CCCTAA", b"TTAGG"])


def test_multi_query_mixed_paths(ctx):
    """Hymenoptera-like table (lengths 5..14), a lower-case repeat (never matches, finder.rs:129),
    a query with N (literal compare), a 40-mer (beyond the plane kernel) and a 70-mer (BOM branch)."""
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.random_records(11, 3, 60000, b"ACGTNacgt", plant=b"TTAGGTTGGG")
    long40 = bytes(seq[100:140]).upper()
    long70 = bytes(seq[500:570]).upper()
    qs = [b"AACCT", b"AACCC", b"AACCCT", b"AACCCAGACGC", b"AACCCAGACCC", b"AACCCCAACCT", b"AACCCAGACCT",
          b"AACCCTAACCTAAC", b"TTAGGTTGGG", b"aacct", b"TTANG", b"NN", long40, long70]
    _cmp(ctx, seq, off, 10000, qs)
    _cmp(ctx, seq, off, 777, qs[:9])
    _cmp(ctx, seq, off, 50000, qs)


def test_search_uppercases_query_find_does_not(ctx):
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(b"ttaggTTAGGttagg" * 10, dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    assert T.search_rows(ctx, ["a"], seq, off, "ttagg", 1000) == O.search_text(["a"], [seq.tobytes()], b"ttagg", 1000)[len(O.SEARCH_HEADER):]
    assert T.finder_rows(ctx, ["a"], seq, off, ["ttagg", "TTAGG"], 1000) == O.find_text(["a"], [seq.tobytes()], [b"ttagg", b"TTAGG"], 1000)[len(O.SEARCH_HEADER):]


def test_long_pattern_up_to_32(ctx):
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.random_records(3, 2, 50000, b"ACGT")
    for L in [13, 16, 17, 31, 32]:
        q = bytes(seq[1000:1000 + L])
        seq[5000:5000 + L] = seq[1000:1000 + L]
        _cmp(ctx, seq, off, 10000, [q])
        _cmp(ctx, seq, off, 4999, [q])


def test_errors(ctx):
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(b"ACGT" * 100, dtype=np.uint8).copy()
    off = np.array([0, seq.size], dtype=np.uint64)
    with pytest.raises(T.TidkCudaError) as e:
        ctx.window_counts(seq, off, 0, [b"ACG"])
    assert e.value.code == -1
    with pytest.raises(T.TidkCudaError) as e:
        ctx.window_counts(seq, off, 10, [b""])
    assert e.value.code == -1
    seq[57] = 0xC3
    with pytest.raises(T.TidkCudaError) as e:
        ctx.window_counts(seq, off, 64, [b"ACG"])
    assert e.value.code == -4
    with pytest.raises(O.OracleError):
        O.window_counts(seq, off, 64, [b"ACG"], False)
    # the ctx stays usable after an error
    seq[57] = ord("A")
    _cmp(ctx, seq, off, 64, [b"ACG"])


def test_empty_inputs(ctx):
    f, r = ctx.window_counts(np.zeros(0, np.uint8), np.array([0], dtype=np.uint64), 100, [b"ACG"])
    assert f.size == 0 and r.size == 0
    f, r = ctx.window_counts(np.zeros(0, np.uint8), np.array([0, 0, 0], dtype=np.uint64), 100, [b"ACG"])
    assert f.size == 0


def test_resident_matches_host_path(ctx):
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.config2(scale=0.004)
    b = ctx.upload(seq, off)
    f1, r1 = ctx.window_counts_resident(b, 10000, [b"AACCT"])
    assert ctx.last_kernel_launches == 2 and ctx.last_kernel_ms > 0  # item table + counting kernel
    f1, r1 = ctx.window_counts_resident(b, 10000, [b"AACCT"])
    assert ctx.last_kernel_launches == 1  # item table cached for (batch, window)
    f2, r2 = ctx.window_counts(seq, off, 10000, [b"AACCT"])
    of, orr = O.window_counts(seq, off, 10000, [b"AACCT"], False)
    assert (f1 == of).all() and (r1 == orr).all() and (f2 == of).all() and (r2 == orr).all()
    assert int(of.sum()) > 0 and int(orr.sum()) > 0
    b.free()


@pytest.mark.parametrize("cfg", ["config1", "config2", "config4"])
def test_configs_scaled_text_identical(ctx, cfg):
    """BASELINE configs at reduced scale: full output text byte-identical to the oracle's."""
    import telomeric_identifier_b200 as T
    from telomeric_identifier_b200 import synth
    ids, seq, off = getattr(synth, cfg)(scale=0.01)
    seqs = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ids))]
    if cfg == "config1":
        for ext in ("tsv", "bedgraph"):
            got = T.search_rows(ctx, ids, seq, off, "TTAGG", 10000, ext)
            want = O.search_text(ids, seqs, b"TTAGG", 10000, ext)
            assert got == (want[len(O.SEARCH_HEADER):] if ext == "tsv" else want)
    elif cfg == "config2":
        hymenoptera = ["AACCT", "AACCC", "AACCCT", "AACCCAGACGC", "AACCCAGACCC", "AACCCCAACCT"]
        for reps in (["AACCT"], hymenoptera):
            got = T.finder_rows(ctx, ids, seq, off, reps, 10000)
            assert got == O.find_text(ids, seqs, [r.encode() for r in reps], 10000)[len(O.SEARCH_HEADER):]
    else:
        got = T.search_rows(ctx, ids, seq, off, "TTAGGG", 10000)
        assert got == O.search_text(ids, seqs, b"TTAGGG", 10000)[len(O.SEARCH_HEADER):]


def test_full_size_properties(ctx):
    """BASELINE config 2 at full size (500 Mb): size-independent properties instead of the oracle.
    (a) swapping the query for its reverse complement swaps fwd and rev;
    (b) counts with W are >= 0 and the sum over windows with W = 2W' never exceeds ... (monotone):
        every occurrence counted at W' = 5000 is also inside a W = 10000 window, so
        sum(W=10000) >= sum(W=5000), and the difference is the number of occurrences that
        straddle the extra boundaries;
    (c) a window-aligned sub-range recomputed alone gives the same counts (oracle-checked)."""
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.config2(scale=1.0)
    b = ctx.upload(seq, off)
    f, r = ctx.window_counts_resident(b, 10000, [b"AACCT"])
    f2, r2 = ctx.window_counts_resident(b, 10000, [b"AGGTT"])
    assert (f == r2).all() and (r == f2).all()
    f5, r5 = ctx.window_counts_resident(b, 5000, [b"AACCT"])
    assert int(f.sum()) >= int(f5.sum()) and int(r.sum()) >= int(r5.sum())
    assert int(f.sum()) - int(f5.sum()) < f5.size
    # chromosome 7, windows 100..140 against the oracle
    o7 = int(off[7])
    sub = seq[o7 + 100 * 10000: o7 + 140 * 10000]
    of, orr = O.window_counts(sub, np.array([0, sub.size], dtype=np.uint64), 10000, [b"AACCT"], False)
    nw = ((off[1:] - off[:-1] + np.uint64(9999)) // np.uint64(10000)).astype(np.int64)
    base = int(nw[:7].sum())
    assert (f[base + 100: base + 140] == of).all() and (r[base + 100: base + 140] == orr).all()
    # the planted telomere arrays are seen at both ends of chromosome 1
    assert int(f[0]) > 100 and int(r[int(nw[0]) - 1]) > 100
    b.free()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_ranks_emulated_on_one_gpu(ctx, world):
    """the ranks' shards computed one after another on this GPU and merged on the host
    equal the single-GPU (and oracle) result -- no halo is needed because shards cut at window multiples."""
    from telomeric_identifier_b200 import sharding, synth
    ids, seq, off = synth.config4(scale=0.004)
    qs = [b"TTAGGG"]
    plans = sharding.plan_shards(off, 10000, world)
    per_rank = []
    for r in range(world):
        ls, lo = sharding.shard_batch(seq, off, plans[r], 10000)
        per_rank.append(ctx.window_counts(ls, lo, 10000, qs))
    f, r = sharding.merge_counts(plans, per_rank, off, 10000, 1)
    of, orr = O.window_counts(seq, off, 10000, qs, False)
    assert (f == of).all() and (r == orr).all()


@pytest.mark.parametrize("clade", ["Coleoptera", "Hemiptera", "Hymenoptera"])
def test_fused_clade_kernels(ctx, clade):
    """Multi-repeat clades run as one fused pass; any other order / subset takes the per-query path.
    Both must equal the oracle."""
    import os
    from telomeric_identifier_b200 import clades, synth
    db = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clades_min.csv")
    reps = [r.encode() for r in clades.return_telomere_sequence(clade, db)]
    ids, seq, off = synth.random_records(21, 4, 120000, b"ACGTNacgt", plant=reps[0] + reps[-1])
    for r in reps:  # plant every repeat somewhere, both strands
        p = 1000 + 37 * len(r)
        seq[p:p + 4 * len(r)] = np.frombuffer(r * 4, dtype=np.uint8)
        rc = O.revcomp(r)
        seq[p + 500:p + 500 + 3 * len(r)] = np.frombuffer(rc * 3, dtype=np.uint8)
    b = ctx.upload(seq, off)
    for W in (10000, 777, 40000):
        f, r = ctx.window_counts_resident(b, W, reps)
        assert ctx.last_kernel_launches <= 2  # (item table +) ONE counting kernel
        of, orr = O.window_counts(seq, off, W, reps, False)
        assert (f == of).all() and (r == orr).all()
    _cmp(ctx, seq, off, 10000, reps[::-1])
    _cmp(ctx, seq, off, 10000, reps[:-1] if len(reps) > 2 else reps + [b"TTAGG"])
    b.free()


@pytest.mark.parametrize("cfg,query", [("config1", b"TTAGG"), ("config4", b"TTAGGG")])
def test_full_size_c1_c4_properties(ctx, cfg, query):
    """BASELINE configs 1 (100 Mb) and 4 (3.1 Gb, N gaps) at FULL size.  The oracle needs ~10 s per Gb,
    so it checks sampled stretches (both ends of a chromosome, an N-gap edge, a soft-masked block);
    the whole output is checked through size-independent properties."""
    from telomeric_identifier_b200 import synth
    ids, seq, off = getattr(synth, cfg)(scale=1.0)
    W = 10000
    b = ctx.upload(seq, off)
    f, r = ctx.window_counts_resident(b, W, [query])
    f2, r2 = ctx.window_counts_resident(b, W, [O.revcomp(query)])
    assert (f == r2).all() and (r == f2).all()                      # strand swap
    fh, rh = ctx.window_counts_resident(b, W // 2, [query])
    assert int(f.sum()) >= int(fh.sum()) and int(f.sum()) - int(fh.sum()) < fh.size  # halving loses only straddlers
    nw = ((off[1:] - off[:-1] + np.uint64(W - 1)) // np.uint64(W)).astype(np.int64)
    base = np.concatenate([[0], np.cumsum(nw)])
    lens = (off[1:] - off[:-1]).astype(np.int64)
    assert f.size == int(nw.sum())
    rng = np.random.default_rng(0)
    spots = []
    for rec in (0, len(ids) // 2, len(ids) - 1):
        spots += [(rec, 0), (rec, max(0, int(nw[rec]) - 30))]       # both chromosome ends (last partial window too)
        spots += [(rec, int(rng.integers(0, max(1, int(nw[rec]) - 30))))]
    # a stretch that crosses into an N gap / a soft-masked block, found from the data itself
    o0 = int(off[1])
    sub = seq[o0:o0 + int(lens[1])]
    odd = np.flatnonzero((sub == ord("N")) | (sub >= 97))
    if odd.size:
        spots.append((1, max(0, int(odd[0]) // W - 5)))
    for rec, w0 in spots:
        w1 = min(w0 + 30, int(nw[rec]))
        a = int(off[rec]) + w0 * W
        e = min(int(off[rec]) + w1 * W, int(off[rec + 1]))
        piece = seq[a:e]
        of, orr = O.window_counts(piece, np.array([0, piece.size], dtype=np.uint64), W, [query], False)
        s = int(base[rec]) + w0
        assert (f[s:s + of.size] == of).all() and (r[s:s + orr.size] == orr).all(), (rec, w0)
    assert int(f.sum()) + int(r.sum()) > 1000                       # the planted arrays are seen
    b.free()


def test_host_packed_upload_path(ctx, monkeypatch):
    """tidk_cuda_window_counts packs the sequence into bit-planes on the host when every query has a
    compile-time kernel (2 bits per base + validity exceptions cross PCIe instead of 8).  Same counts as the oracle; queries
    without a compile-time kernel, and small inputs, keep taking the plain upload."""
    import os
    import telomeric_identifier_b200 as T
    from telomeric_identifier_b200 import clades, synth
    db = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clades_min.csv")
    ids, seq, off = synth.config4(scale=0.012)          # ~37 Mb with N gaps, above the 32 MiB threshold
    seq = seq.copy()
    seq[1_000_000:1_200_000] |= 0x20                     # a soft-masked stretch
    seq[5_000_000:5_000_100] = np.frombuffer(b"RYKMSWBDHV" * 10, dtype=np.uint8)
    hym = [r.encode() for r in clades.return_telomere_sequence("Hymenoptera", db)]
    monkeypatch.setenv("TIDK_HOST_PACK", "1")
    # hybrid (two-ended) upload: ASCII chunks by DMA from the front, packed chunks from the back, the
    # windows split between the two sources wherever the two meet; "0" = everything packed
    # The DMA front reads the caller's buffer directly, so it is taken for PINNED sources only (a pageable source --
    # a plain numpy array, a malloc'ed FASTA buffer -- travels all-packed whatever the knob says): both are run.
    pinned = ctx.pinned_empty(seq.size)
    pinned[:] = seq
    splits = set()
    for hybrid, threads, src in (("1", "2", pinned), ("1", "16", pinned), ("0", "16", pinned), ("1", "16", seq)):   # 2 packers: the DMA front takes a real share
        monkeypatch.setenv("TIDK_HYBRID_UPLOAD", hybrid)
        monkeypatch.setenv("TIDK_PACK_THREADS", threads)
        for qs, W in (([b"TTAGGG"], 10000), ([b"CCCTAA", b"AACCT"], 4321), (hym, 10000), ([b"TTAGGG"], 50000),
                      ([b"TTAGGG"], 32768)):
            f, r = ctx.window_counts(src, off, W, qs)
            h2d, d2h, was_packed = ctx.last_transfer()
            assert was_packed
            if src is pinned:
                splits.add((hybrid, threads, h2d))
            else:
                assert h2d < seq.size // 2, "a pageable source must not take the DMA front"
            of, orr = O.window_counts(seq, off, W, qs, False, threads=4)
            assert (f == of).all() and (r == orr).all(), (qs, W, hybrid, threads)
    monkeypatch.delenv("TIDK_HYBRID_UPLOAD")
    monkeypatch.delenv("TIDK_PACK_THREADS")
    assert len({h for hy, th, h in splits if hy == "1"}) > 1 or any(h > seq.size // 2 for hy, th, h in splits if hy == "1")
    # non-ASCII near the front (the part that may travel as ASCII): caught by the kernel or by the packer
    bad0 = seq.copy()
    bad0[77] = 0xC3
    with pytest.raises(T.TidkCudaError) as e0:
        ctx.window_counts(bad0, off, 10000, [b"TTAGGG"])
    assert e0.value.code == -4
    # a query without a compile-time kernel falls back to the plain upload, same answer
    f, r = ctx.window_counts(seq, off, 10000, [b"GATTACA", b"TTAGGG"])
    of, orr = O.window_counts(seq, off, 10000, [b"GATTACA", b"TTAGGG"], False, threads=4)
    assert (f == of).all() and (r == orr).all()
    # non-ASCII is caught by the host packer
    bad = seq.copy()
    bad[123_457] = 0xE2
    with pytest.raises(T.TidkCudaError) as e:
        ctx.window_counts(bad, off, 10000, [b"TTAGGG"])
    assert e.value.code == -4
    monkeypatch.setenv("TIDK_HOST_PACK", "0")
    f2, r2 = ctx.window_counts(seq, off, 10000, [b"TTAGGG"])
    monkeypatch.setenv("TIDK_HOST_PACK", "1")
    f3, r3 = ctx.window_counts(seq, off, 10000, [b"TTAGGG"])
    assert (f2 == f3).all() and (r2 == r3).all()


def test_many_runtime_queries_fused_kernel(ctx):
    """Nine or more queries without a compile-time kernel share one pass of the fused run-time kernel
    (mixed lengths 3..32, consecutive and non-consecutive query indices)."""
    from telomeric_identifier_b200 import synth
    rng = np.random.default_rng(42)
    ids, seq, off = synth.random_records(31, 3, 90000, b"ACGTNacgt")
    qs = []
    for L in (3, 4, 6, 7, 9, 12, 13, 17, 21, 25, 31, 32):
        p = int(rng.integers(0, seq.size - 40))
        q = bytes(seq[p:p + L]).upper().replace(b"N", b"A")
        qs.append(q)
        seq[p + 100:p + 100 + L] = np.frombuffer(q, dtype=np.uint8)
    _cmp(ctx, seq, off, 10000, qs)
    _cmp(ctx, seq, off, 999, qs[:5] + [b"TTAGGG"] + qs[5:])  # a compile-time motif in the middle splits the run
    _cmp(ctx, seq, off, 40000, qs + qs[:8])                   # 20 run-time queries: two fused launches


# ---- asynchronous entry point (tidk_cuda_window_counts_resident_async + tidk_cuda_wait) ----
def test_async_equals_blocking(ctx):
    from telomeric_identifier_b200 import synth
    ids, seq, off = synth.config4(scale=0.002)
    batch = ctx.upload(seq, off)
    want = {}
    for W, q in ((10000, b"TTAGGG"), (777, b"CCCTAA"), (50000, b"TTAGGG"), (10000, b"ACGTNA")):
        f, r = ctx.window_counts_resident(batch, W, [q])
        of, orr = O.window_counts(seq, off, W, [q], False)
        assert np.array_equal(f, of) and np.array_equal(r, orr)
        want[(W, q)] = (f.copy(), r.copy())
    # several calls in flight, different windows / queries back to back, more calls than slots
    outs = []
    for rep in range(3):
        for (W, q), (f, r) in want.items():
            po = ctx.pinned_empty(2 * f.size * 8).view(np.uint64)
            po[:] = 0xDEADBEEF
            o = (po[:f.size], po[f.size:])
            ctx.window_counts_resident_async(batch, W, [q], o)
            outs.append(((W, q), o))
    ms, nl = ctx.wait()
    assert nl >= len(outs) and ms > 0
    for key, o in outs:
        assert np.array_equal(o[0], want[key][0]) and np.array_equal(o[1], want[key][1]), key
    # a blocking call behind asynchronous ones collects them first
    po = ctx.pinned_empty(2 * want[(10000, b"TTAGGG")][0].size * 8).view(np.uint64)
    n = want[(10000, b"TTAGGG")][0].size
    ctx.window_counts_resident_async(batch, 10000, [b"TTAGGG"], (po[:n], po[n:]))
    f, r = ctx.window_counts_resident(batch, 777, [b"CCCTAA"])
    assert np.array_equal(po[:n], want[(10000, b"TTAGGG")][0])
    assert np.array_equal(f, want[(777, b"CCCTAA")][0]) and np.array_equal(r, want[(777, b"CCCTAA")][1])
    assert ctx.wait()[1] >= 1  # the asynchronous call's launches are still reported by the next wait
    batch.free()


def test_async_reports_non_ascii_at_wait(ctx, monkeypatch):
    """A byte >= 0x80 (the reference fails in str::from_utf8, search.rs:107).  Counting from the ASCII copy, the kernel finds
    it and the asynchronous call can only report it when it is collected; counting from the resident planes, K1 has seen it
    at upload and the call itself refuses."""
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(b"ACGT" * 5000, dtype=np.uint8).copy()
    seq[12345] = 0xC3
    off = np.array([0, seq.size], dtype=np.uint64)
    batch = ctx.upload(seq, off)
    po = ctx.pinned_empty(2 * 2 * 8).view(np.uint64)
    monkeypatch.setenv("TIDK_PLANES_COUNT", "0")
    ctx.window_counts_resident_async(batch, 10000, [b"ACGT"], (po[:2], po[2:]))
    with pytest.raises(T.TidkCudaError) as ei:
        ctx.wait()
    assert ei.value.code == T.api.TIDK_ENONASCII
    ctx.wait()  # the error is reported once
    monkeypatch.delenv("TIDK_PLANES_COUNT")
    for q in (b"ACGT", b"TTAGGG"):  # run-time and compile-time flat kernels
        with pytest.raises(T.TidkCudaError) as ei:
            ctx.window_counts_resident_async(batch, 10000, [q], (po[:2], po[2:]))
        assert ei.value.code == T.api.TIDK_ENONASCII
        with pytest.raises(T.TidkCudaError) as ei:
            ctx.window_counts_resident(batch, 10000, [q])
        assert ei.value.code == T.api.TIDK_ENONASCII
    ctx.wait()
    batch.free()


# ---------------------------------------------------------------- flat form on resident planes (window_counts_flat.cuh)
@pytest.mark.parametrize("seed", range(8))
def test_flat_resident_random(ctx, seed, monkeypatch):
    """Resident batches are counted by the flat kernels (spans of 32 Ki bases, windows by accounting): ragged and empty
    records, N / IUPAC / lower case, windows from 1 base to several spans, compile-time motifs, their reverse
    complements and run-time patterns of every length class -- against the oracle and against the per-window kernels."""
    from telomeric_identifier_b200 import synth
    rng = np.random.default_rng(9000 + seed)
    alpha = [b"ACGT", b"ACGTN", b"ACGTacgtNRY", b"AC"][seed % 4]
    plant = [b"TTAGGG", b"AACCT", b"CCCTAA", b"GATTACA"][seed % 4]
    ids, seq, off = synth.random_records(700 + seed, 1 + 5 * (seed % 3), [90_000, 300_000, 40_000][seed % 3], alpha, plant=plant, p_empty=0.15)
    if seq.size == 0:
        pytest.skip("empty batch")
    b = ctx.upload(seq, off)
    try:
        assert b.planes_info()[0]
        windows = [1, 7, 31, 32, 33, 100, 1000, 4095, 4096, 4097, 10000, 32768, 32769, 70001, 10 ** 9]
        queries = [plant, b"TTAGGG", b"CCCTAA", b"A", b"AC", b"GATTACAGATTA", b"ACGTACGTACGTACGTACGTACGTACGTACGT",
                   bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(2, 32))))]
        for W in [windows[i] for i in rng.choice(len(windows), size=6, replace=False)]:
            qs = [queries[i] for i in rng.choice(len(queries), size=3, replace=False)]
            f, r = ctx.window_counts_resident(b, W, qs)
            f, r = f.copy(), r.copy()
            of, orr = O.window_counts(seq, off, W, qs, False)
            assert np.array_equal(f, of) and np.array_equal(r, orr), (seed, W, qs)
            monkeypatch.setenv("TIDK_FLAT", "0")
            f0, r0 = ctx.window_counts_resident(b, W, qs)
            monkeypatch.delenv("TIDK_FLAT")
            assert np.array_equal(f0, of) and np.array_equal(r0, orr), (seed, W, qs, "per-window kernels")
    finally:
        b.free()


def test_flat_window_straddles_and_span_edges(ctx):
    """Occurrences that straddle a window boundary are lost (search.rs:98 cuts the windows first), also when the
    boundary coincides with a 4 096-base step or a 32 Ki-base span edge of the flat kernel, or with a record edge."""
    unit = b"TTAGGG"
    n = 3 * 32768 + 500
    seq = np.frombuffer((unit * (n // 6 + 1))[:n], dtype=np.uint8).copy()
    for cut in (4096, 32768, 65536):           # a second record that starts exactly on a step / span edge
        off = np.array([0, cut, n], dtype=np.uint64)
        b = ctx.upload(seq, off)
        try:
            for W in (4096, 4098, 8192, 32768, 32770, 6, 5, 11):
                for q in (unit, b"GGGTTA", b"TTAGGGTTAGGG"):
                    f, r = ctx.window_counts_resident(b, W, [q])
                    of, orr = O.window_counts(seq, off, W, [q], False)
                    assert np.array_equal(f, of) and np.array_equal(r, orr), (cut, W, q)
        finally:
            b.free()
