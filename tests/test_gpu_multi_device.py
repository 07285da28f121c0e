"""One context over several devices (tidk_cuda_create with n_devices > 1, csrc/multi_device.cuh) against the single-device
context and the oracle.  The same device id is listed two or three times, which a one-GPU box can run: every part is an
ordinary single-device context, only the cutting and the placing of the results is under test."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[2, 3])
def mctx(request):
    import os
    import telomeric_identifier_b200 as T
    # TIDK_TEST_DEVICES="0,1,2": distinct devices on a multi-GPU box (cycled to the requested count); default: device 0 listed n times
    pool = [int(x) for x in os.environ.get("TIDK_TEST_DEVICES", "0").split(",")]
    c = T.Context([pool[i % len(pool)] for i in range(request.param)])
    yield c
    c.close()


def _records(seed, n, max_len, plant=b"TTAGGG"):
    from telomeric_identifier_b200 import synth
    return synth.random_records(seed, n, max_len, b"ACGTacgtN", plant=plant, p_empty=0.1)


@pytest.mark.parametrize("seed,window", [(1, 1000), (2, 977), (3, 10000), (4, 64)])
def test_window_counts_host_pointers(mctx, seed, window):
    ids, seq, off = _records(300 + seed, 1 + 3 * seed, 60000)
    queries = [b"TTAGGG", b"AACCT", b"CCCTAACCCTAA"]
    f, r = mctx.window_counts(seq, off, window, queries)
    wf, wr = O.window_counts(seq, off, window, queries, False)
    assert np.array_equal(f, wf) and np.array_equal(r, wr)


def test_window_counts_resident_blocking_and_async(mctx):
    ids, seq, off = _records(41, 9, 120000)
    batch = mctx.upload(seq, off)
    try:
        for window, queries in ((5000, [b"TTAGGG"]), (1234, [b"AACCT", b"TTAGG"])):
            f, r = mctx.window_counts_resident(batch, window, queries)
            wf, wr = O.window_counts(seq, off, window, queries, False)
            assert np.array_equal(f, wf) and np.array_equal(r, wr)
            n = f.size
            outs = [(mctx.pinned_empty(8 * max(n, 1)).view(np.uint64), mctx.pinned_empty(8 * max(n, 1)).view(np.uint64)) for _ in range(2)]
            for i in range(4):
                mctx.window_counts_resident_async(batch, window, queries, outs[i & 1])
            ms, launches = mctx.wait()
            assert launches > 0
            for o in outs:
                assert np.array_equal(o[0][:n], wf) and np.array_equal(o[1][:n], wr)
        ready, ms, nbytes = batch.planes_info()
        assert ready and nbytes > 0
    finally:
        batch.free()


@pytest.mark.parametrize("seed,d", [(5, 0.5), (6, 0.1), (7, 0.33)])
def test_explore_host_and_resident(mctx, seed, d):
    ids, seq, off = _records(500 + seed, 2 + 2 * seed, 90000, plant=b"AACCT")
    want, _ = O.explore_runs(seq, off, d, 3, 9, 2, period_filter=False)
    got = mctx.explore_runs(seq, off, d, 3, 9, 2)
    batch = mctx.upload(seq, off)
    try:
        res = mctx.explore_runs_resident(batch, d, 3, 9, 2)
    finally:
        batch.free()
    for g in (got, res):
        assert g.size == want.size
        for f in ("record", "slice", "k", "first", "start", "end"):
            assert (g[f] == want[f]).all(), f


def test_errors_are_the_single_device_ones(mctx):
    import telomeric_identifier_b200 as T
    ids, seq, off = _records(9, 3, 5000)
    with pytest.raises(T.TidkCudaError) as e:
        mctx.window_counts(seq, off, 0, [b"TTAGGG"])
    assert e.value.code == -1
    with pytest.raises(T.TidkCudaError) as e:
        mctx.explore_runs(seq, off, 0.7, 5, 6, 0)
    assert e.value.code == -1
    bad = seq.copy()
    bad[off[1] // 2] = 0xC3
    with pytest.raises(T.TidkCudaError) as e:
        mctx.window_counts(bad, off, 1000, [b"TTAGGG"])
    assert e.value.code == -4


def test_more_devices_than_records(mctx):
    seq = np.frombuffer(b"TTAGGG" * 500, dtype=np.uint8).copy()
    off = np.array([0, seq.size], dtype=np.uint64)
    f, r = mctx.window_counts(seq, off, 3000, [b"TTAGGG"])
    wf, wr = O.window_counts(seq, off, 3000, [b"TTAGGG"], False)
    assert np.array_equal(f, wf) and np.array_equal(r, wr)
    got = mctx.explore_runs(seq, off, 0.5, 5, 7, 1)
    want, _ = O.explore_runs(seq, off, 0.5, 5, 7, 1, period_filter=False)
    assert got.size == want.size and (got["end"] == want["end"]).all()
