"""Host logic of the multi-GPU path on CPU: byte-balanced sharding, and the
final gather over a 2-rank gloo group.  The matcher itself is stood in for by
the oracle here (test infrastructure) - what is under test is that the ranges
tile the corpus and that the gathered results equal the single-process ones."""

import os
import sys

import numpy as np
import pytest

from conftest import ROOT, unz


def test_byte_balanced_ranges_tile_and_balance():
    from one_b200 import shard
    rng = np.random.RandomState(1)
    for n, world in ((0, 2), (1, 4), (7, 8), (1000, 2), (100000, 8), (5, 1)):
        lens = rng.randint(0, 300, size=n)
        ranges = shard.byte_balanced_ranges(lens, world)
        assert len(ranges) == world
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        for (a, b), (c, d) in zip(ranges, ranges[1:]):
            assert a <= b == c <= d
        if n >= 1000:
            per = [int(lens[a:b].sum()) for a, b in ranges]
            assert max(per) - min(per) <= 2 * 300, per
    # a pathological corpus: one huge line
    lens = np.array([10, 10, 10**6, 10, 10])
    ranges = shard.byte_balanced_ranges(lens, 4)
    assert sum(b - a for a, b in ranges) == 5


def _worker(rank, world, port, blob, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from one_b200 import shard, synth
    from oracle.loader import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text, starts, lens = synth.make_corpus(6000, seed=77)  # same corpus on every rank
    a, b = shard.byte_balanced_ranges(lens, world)[rank]
    res, _, ends = Oracle(blob).batch(text, starts[a:b], lens[a:b], "match", "instant")
    counts = torch.from_numpy(np.bincount(res, minlength=65).astype(np.int64))
    allc = shard.gather_counts(counts)
    idx = torch.from_numpy(np.nonzero(res > 0)[0].astype(np.int64))
    lines = shard.gather_matched_lines(idx, a)
    if rank == 0:
        q.put((torch.stack(allc).sum(0).numpy(), lines.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(vectors, oracle_cls):
    import torch.multiprocessing as mp
    from one_b200 import synth
    blob = unz(next(c for c in vectors["cases"] if c["name"] == "union64")["program"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blob, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, lines = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    text, starts, lens = synth.make_corpus(6000, seed=77)
    res, _, _ = oracle_cls(blob).batch(text, starts, lens, "match", "instant")
    assert np.array_equal(counts, np.bincount(res, minlength=65))
    assert np.array_equal(lines, np.nonzero(res > 0)[0])


def test_periodic_ranges_equal_the_explicit_ones():
    """bench.py cuts a pool of lines repeated many times without building the
    repeated array: same cuts as byte_balanced_ranges on the array itself."""
    import numpy as np
    from one_b200 import shard
    rng = np.random.RandomState(3)
    for _ in range(400):
        m, reps, world = rng.randint(1, 14), rng.randint(1, 9), rng.randint(1, 9)
        pool = rng.randint(0, 6, m).astype(np.uint32)
        assert shard.byte_balanced_ranges_periodic(pool, reps, world) == \
            shard.byte_balanced_ranges(np.tile(pool, reps), world), (pool, reps, world)
