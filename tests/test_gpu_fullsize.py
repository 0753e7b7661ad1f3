"""BASELINE.json's full sizes on the GPU, checked through properties that do
not depend on the size: the bench corpus is a pool of 200 000 distinct lines
tiled to 100 M lines (configs[1]) / to a 16 GB stream (configs[3]), so the
outcomes must repeat with the pool's period, and the first period must equal
the oracle's answer on the pool.  GPU only."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_100M_lines_are_periodic_and_match_the_oracle(gpu_red, oracle_cls):
    import torch
    import bench
    pool = bench.host_pool(1234)
    blob = bench.load_program_blob("union64")
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    dev = torch.device("cuda", 0)
    text, starts, lens, n = bench.device_corpus(pool, 100_000_000, dev)
    assert n == 100_000_000 and text.numel() > 10_000_000_000
    P = bench.POOL_LINES
    for verb, sty, want_start in (("match", "instant", False), ("match", "last", True), ("check", "full", False)):
        res = torch.empty(n, dtype=torch.int32, device=dev)
        ends = torch.empty(n, dtype=torch.int32, device=dev)
        sts = torch.empty(n, dtype=torch.int32, device=dev) if want_start else None
        prog.match_batch_device(text, starts, lens, verb, sty, out=(res, sts, ends))
        torch.cuda.synchronize()
        eres, ems, eme = orc.batch(pool[0], pool[1], pool[2], verb, sty)
        assert np.array_equal(res[:P].cpu().numpy(), eres), (verb, sty)
        if verb == "match":
            assert np.array_equal(ends[:P].cpu().numpy().view(np.uint32), eme), (verb, sty)
        if want_start:
            assert np.array_equal(sts[:P].cpu().numpy().view(np.uint32), ems), (verb, sty)
        # every later period repeats the first
        assert bool((res.view(-1, P) == res[:P]).all()), (verb, sty)
        if verb == "match":
            assert bool((ends.view(-1, P) == ends[:P]).all()), (verb, sty)
        if want_start:
            assert bool((sts.view(-1, P) == sts[:P]).all()), (verb, sty)
        assert int((res > 0).sum().item()) == int((eres > 0).sum()) * (n // P)
        del res, ends, sts
    # compaction at full size: ascending, complete
    res = torch.empty(n, dtype=torch.int32, device=dev)
    prog.match_batch_device(text, starts, lens, "match", "instant", out=(res, None, None))
    idx = gpu_red.compact_matches_device(res)
    assert idx.numel() == int((res > 0).sum().item())
    assert bool((idx[1:] > idx[:-1]).all()) and bool((res[idx] > 0).all())
    # line index at full size reproduces the corpus index
    s2, l2 = gpu_red.split_lines_device(text, capacity=n + 1)
    assert s2.numel() == n and bool((s2 == starts).all()) and bool((l2 == lens).all())
    prog.close()


def test_16GB_stream_shifts_with_the_period(gpu_red, oracle_cls):
    """configs[3]: the DFA forgets at every newline, so on R copies of a text the
    last match is the last match of one copy moved by R - 1 periods, and the
    first match is the first match of one copy."""
    import torch
    import bench
    pool = bench.host_pool(1234)
    blob = bench.load_program_blob("union64")
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    dev = torch.device("cuda", 0)
    period = torch.from_numpy(pool[0].copy()).to(dev)
    raw = pool[0].tobytes()
    assert raw.endswith(b"\n")
    R = int(16e9 // period.numel()) + 1
    big = period.repeat(R)
    assert big.numel() >= 16_000_000_000
    shift = (R - 1) * period.numel()
    for sty in ("instant", "first", "tangent", "last", "full"):
        one = orc.run(raw, "match", sty, False)
        got = tuple(prog.match_stream_device(big, "match", sty))
        if sty in ("last", "full"):
            exp = (one[0], one[1] + shift, one[2] + shift) if one[0] else (0, 0, 0)
        else:
            exp = one
        assert got == exp, (sty, got, exp)
        assert prog.match_stream_device(big, "check", sty).result == exp[0]
    prog.close()
