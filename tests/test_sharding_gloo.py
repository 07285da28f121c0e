"""Host-side multi-GPU logic on CPU: plan/merge, and a world_size-2 gloo run in which each rank
counts its shard (the oracle stands in for the device here -- it is the checker, and this test is
about sharding, not about the kernel) and rank 0 re-assembles the global table."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle as O
from telomeric_identifier_b200 import sharding, synth


def _oracle_count(seq, off, window, queries):
    return O.window_counts(seq, off, window, list(queries), False)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("window", [1000, 10000, 7])
def test_plan_and_merge_roundtrip(world, window):
    ids, seq, off = synth.random_records(5, 7, 40000 if window > 7 else 300, b"ACGTacgtN", plant=b"TTAGGG", p_empty=0.2)
    plans = sharding.plan_shards(off, window, world)
    nw = sharding.windows_per_record(off, window)
    assert sum(n for p in plans for _, _, n in p) == int(nw.sum())
    sizes = [sum(n for _, _, n in p) for p in plans]
    assert max(sizes) - min(sizes) <= 1  # balanced by windows (== bases up to one window)
    per_rank = []
    for r in range(world):
        ls, lo = sharding.shard_batch(seq, off, plans[r], window)
        per_rank.append(_oracle_count(ls, lo, window, [b"TTAGGG", b"AACCT"]))
    f, r = sharding.merge_counts(plans, per_rank, off, window, 2)
    wf, wr = _oracle_count(seq, off, window, [b"TTAGGG", b"AACCT"])
    assert (f == wf).all() and (r == wr).all()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids, seq, off = synth.config2(scale=0.002)
    out = sharding.sharded_window_counts(_oracle_count, seq, off, 10000, [b"AACCT"], rank, world, dist)
    if rank == 0:
        wf, wr = _oracle_count(seq, off, 10000, [b"AACCT"])
        q.put(bool((out[0] == wf).all() and (out[1] == wr).all() and int(wf.sum()) > 0))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_counts():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
