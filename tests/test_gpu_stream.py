"""One long byte stream on the GPU (chunked, speculated start states, stitched)
against the oracle's match() over the same text.  GPU only."""

import numpy as np
import pytest

from conftest import unz

pytestmark = pytest.mark.gpu

STYLES = ["instant", "first", "tangent", "last", "full"]


def to_dev(raw):
    import torch
    buf = bytearray(raw) + bytearray(64)
    return torch.frombuffer(buf, dtype=torch.uint8).cuda()[:len(raw)]


def check_stream(gpu_red, oracle_cls, blob, raw, what):
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    dt = to_dev(raw)
    for sty in STYLES:
        exp = orc.run(raw, "match", sty, False)
        got = tuple(prog.match_stream_device(dt, "match", sty))
        assert got == exp, (what, sty, got, exp)
        assert prog.match_stream_device(dt, "check", sty).result == exp[0]
    prog.close()


def test_stream_union64_on_log_text(gpu_red, vectors, oracle_cls):
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    text, _, _ = synth.make_corpus(40000, seed=21)
    raw = text.tobytes()
    check_stream(gpu_red, oracle_cls, blob, raw, "union64")
    # no match anywhere: only noise lines
    noise = b"\n".join(l for l in synth.make_lines(30000, seed=5) if b"CRON" in l or b"ntpd" in l)
    check_stream(gpu_red, oracle_cls, blob, noise, "union64-noise")
    # a single match at the very end / very start
    check_stream(gpu_red, oracle_cls, blob, noise + b" Quarantined", "tail")
    check_stream(gpu_red, oracle_cls, blob, b"Quarantined " + noise, "head")


@pytest.mark.parametrize("env", [
    {"RED_B200_STREAM_HOT": "0"},                                        # the walk on the full table (chunk engine)
    {"RED_B200_STREAM_HOT": "1", "RED_B200_STREAM_HOT_STATES": "257"},   # two-byte codes, lanes leave the image
    {"RED_B200_STREAM_HOT": "1", "RED_B200_HOT_PROFILE": "0"},           # one-byte codes by breadth-first order
    {"RED_B200_HOT_ASCII": "-1"},                                        # 256-column rows
    {"RED_B200_HOT_ASCII": "1"},                                         # 128-column rows whatever the text holds
    {"RED_B200_STREAM_HOT_WARPS": "5"},
])
def test_stream_hot_kernel_variants(gpu_red, vectors, oracle_cls, monkeypatch, env):
    """stream_hot_kernel: code widths, row widths, the way out of the image and
    back, bytes >= 128 - all against the oracle."""
    from one_b200 import synth
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    blob = unz(next(c for c in vectors["cases"] if c["name"] == "union64")["program"])
    text, _, _ = synth.make_corpus(30000, seed=33)
    raw = bytearray(text.tobytes())
    check_stream(gpu_red, oracle_cls, blob, bytes(raw), ("plain", env))
    rng = np.random.RandomState(9)
    for at in rng.randint(0, len(raw), size=400):
        raw[at] = 0x80 + (at % 120)
    check_stream(gpu_red, oracle_cls, blob, bytes(raw), ("high bytes", env))


def test_stream_edge_sizes(gpu_red, vectors, oracle_cls):
    case = next(c for c in vectors["cases"] if c["name"] == "newyork")
    blob = unz(case["program"])
    rng = np.random.RandomState(3)
    for n in (1, 15, 16, 17, 255, 1024, 1025, 16384, 16385, 70001):
        body = bytes(rng.choice(list(b"abc NewYork."), size=n).astype(np.uint8))
        check_stream(gpu_red, oracle_cls, blob, body, n)
    import torch
    prog = gpu_red.Program(blob)
    assert tuple(prog.match_stream_device(torch.empty(0, dtype=torch.uint8, device="cuda"), "match", "last")) == (0, 0, 0)
    prog.close()


@pytest.mark.parametrize("name", ["single_auto", "loose_both", "n3_fmt2", "nullable", "initloop", "exact", "bytes"])
def test_stream_other_programs(gpu_red, vectors, oracle_cls, name):
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == name)
    blob = unz(case["program"])
    rng = np.random.RandomState(9)
    text, _, _ = synth.make_corpus(5000, seed=2)
    check_stream(gpu_red, oracle_cls, blob, text.tobytes(), name)
    for alphabet in (b"ab", b"abc0123456789", b"abc\x00\xff\x80", b"(ab|c)xyz"):
        body = bytes(rng.choice(list(alphabet), size=200000).astype(np.uint8))
        check_stream(gpu_red, oracle_cls, blob, body, (name, alphabet))


def test_stream_dfa_that_never_forgets(gpu_red, redlib, oracle_cls):
    """A parity counter: the guess from the lookback window is wrong half of
    the time and a wrong start gives a wrong end, so stitching has to iterate.
    Still exact."""
    p = redlib.Parser()
    p.add(b"((a|b)(a|b))*", 1, 0)
    blob = redlib.compileToSerialized(p)
    for n in (40000, 40001):
        check_stream(gpu_red, oracle_cls, blob, b"ab" * (n // 2) + b"a" * (n % 2), ("parity", n))


def test_stream_windows(gpu_red, redlib, vectors, oracle_cls, monkeypatch):
    """The stream is taken in windows of growing size (32 MB, x8); with a tiny
    first window a small text crosses many of them: the state, the last
    accept and the last departure from the initial state carry over, and a
    walk that is decided stops at its window."""
    from one_b200 import synth
    monkeypatch.setenv("RED_B200_STREAM_WINDOW", "8192")
    union = unz(next(c for c in vectors["cases"] if c["name"] == "union64")["program"])
    noise = b"\n".join(l for l in synth.make_lines(30000, seed=6) if b"CRON" in l or b"ntpd" in l)
    assert len(noise) > 400000
    check_stream(gpu_red, oracle_cls, union, noise, "w-noise")
    for at in (100, 8191, 8192, 40959, 40960, 300000, len(noise) - 1):
        check_stream(gpu_red, oracle_cls, union, noise[:at] + b" Quarantined " + noise[at:], ("w-union", at))
    # tight patterns: non-absorbing accepts, results that change, dead ends
    p = redlib.Parser()
    p.add(b"[a-c ]*NewYork[a-c ]*", 1, 0)
    p.add(b"[a-c ]*NewYork[a-c ]*\\.[a-c .]*", 2, 0)
    tight = redlib.compileToSerialized(p)
    rng = np.random.RandomState(8)
    for n in (9000, 50000, 200000):
        body = bytes(rng.choice(list(b"abc "), size=n).astype(np.uint8))
        for at in (10, 8190, n // 2, n - 8):
            for extra in (b"", b".", b"!"):
                k = min(at + 9000, n)
                raw = body[:at] + b"NewYork" + body[at:k] + extra + body[k:]
                check_stream(gpu_red, oracle_cls, tight, raw, ("w-tight", n, at, extra))
    parity = redlib.Parser()
    parity.add(b"((a|b)(a|b))*", 1, 0)
    check_stream(gpu_red, oracle_cls, redlib.compileToSerialized(parity), b"ab" * 30000 + b"a", "w-parity")


def test_stream_start_far_behind_the_match(gpu_red, redlib, oracle_cls, monkeypatch):
    """The start position is searched for backwards from the match, lazily.
    Worst case: the initial state is left at byte 0 and never again, the match
    is at the very end - the search has to go back through every piece and
    every window.  Also with exits scattered thinly, and with none at all."""
    p = redlib.Parser()
    p.add(b"a.*z", 1, 0)
    p.add(b"[b-y ]*q.*z", 2, 0)
    blob = redlib.compileToSerialized(p)
    rng = np.random.RandomState(17)
    body = bytes(rng.choice(list(b"bcdefgh ijklmnop"), size=900000).astype(np.uint8))
    cases = [b"a" + body + b"z", b"a" + body + b"z" + body[:5000],
             body[:300000] + b"q" + body[300000:] + b"z", body + b"z", b"a" + body]
    for small in (False, True):
        if small:
            monkeypatch.setenv("RED_B200_STREAM_WINDOW", "16384")
        for i, raw in enumerate(cases):
            check_stream(gpu_red, oracle_cls, blob, raw, ("far", small, i))


def test_large_dfa_uses_global_table(gpu_red, redlib, oracle_cls):
    """BASELINE configs[2]: a union too large for shared memory (L2-resident
    global table, 4-byte entries), first/last styles."""
    from one_b200 import synth
    p = redlib.Parser()
    rules = synth.rules_keywords(2000)
    for pat, res, flags, lang in rules:
        p.addAs(lang, pat, res, flags)
    blob = redlib.compileToSerialized(p)
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    info = prog.info
    assert info["table_mode"] == 2 and info["state_count"] > 10000, info
    rng = np.random.RandomState(4)
    lines = synth.make_lines(30000, seed=12)
    # plant keywords into a third of the lines
    for i in range(0, len(lines), 3):
        kw = rules[rng.randint(len(rules))][0]
        k = rng.randint(0, len(lines[i]))
        lines[i] = lines[i][:k] + kw + lines[i][k:]
    text, starts, lens = synth.pack_lines(lines)
    # (30 000 lines: the first call's profile re-orders the states of the image
    # the batch kernels use - the visited ones in front, for the shared-memory
    # mirror; every style sees that image)
    for verb in ("check", "match"):
        for sty in ("first", "last", "instant", "tangent", "full"):
            res, ms, me = prog.match_batch_host(text, starts, lens, verb, sty)
            eres, ems, eme = orc.batch(text, starts, lens, verb, sty)
            assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), (verb, sty)
            if sty == "instant":
                assert (eres > 0).sum() > 5000
    check_stream(gpu_red, oracle_cls, blob, text.tobytes()[:400000], "keywords2000")
    prog.close()


def dev_repeat(unit, total, tail=b""):
    """unit repeated up to `total` bytes (+ tail) on the GPU and as bytes"""
    import torch
    reps = total // len(unit)
    d = torch.frombuffer(bytearray(unit), dtype=torch.uint8).cuda().repeat(reps)
    if tail:
        d = torch.cat([d, torch.frombuffer(bytearray(tail), dtype=torch.uint8).cuda()])
    d = torch.cat([d, torch.zeros(64, dtype=torch.uint8, device="cuda")])[:reps * len(unit) + len(tail)]
    return d, unit * reps + tail


def test_stream_1GB_dfa_that_never_forgets(gpu_red, redlib, oracle_cls):
    """A parity counter over 1 GB: every chunk is walked from each of the DFA's
    few states, the chunks' state maps are composed on the device - a fixed,
    small number of launches, exact."""
    p = redlib.Parser()
    p.add(b"((a|b)(a|b))*", 1, 0)
    blob = redlib.compileToSerialized(p)
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    for tail in (b"", b"a"):
        dt, raw = dev_repeat(b"ab" * 4096, 1 << 30, tail)
        for sty in ("full", "last", "instant"):
            before = prog.launch_count
            got = tuple(prog.match_stream_device(dt, "match", sty))
            assert prog.launch_count - before <= 40, (sty, prog.launch_count - before)
            assert got == orc.run(raw, "match", sty, False), (tail, sty, got)
        del dt, raw
    prog.close()


def test_stream_1GB_accepting_run(gpu_red, redlib, oracle_cls):
    """Tangent and First keep going while the text accepts: a run of a gigabyte
    of digits is ended in parallel (every chunk behind the first match reports
    where the run would break), not by one thread."""
    p = redlib.Parser()
    p.add(b"x[0-9]+", 1, 2)        # loose start
    p.add(b"x[0-9]+y[0-9]*", 2, 2)
    blob = redlib.compileToSerialized(p)
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    import torch
    n = 1 << 30
    for head, tail in ((b"abc x", b"!rest"), (b"x", b""), (b"zz x", b"y77 ")):
        raw = head + b"7" * n + tail
        dt = torch.cat([torch.frombuffer(bytearray(head), dtype=torch.uint8).cuda(),
                        torch.full((n,), 55, dtype=torch.uint8, device="cuda"),
                        torch.frombuffer(bytearray(tail + bytes(64)), dtype=torch.uint8).cuda()])[:len(raw)]
        # the text starts at a 16-byte boundary, as the entry point asks
        assert dt.data_ptr() % 16 == 0
        for sty in ("tangent", "first", "last"):
            before = prog.launch_count
            got = tuple(prog.match_stream_device(dt, "match", sty))
            assert prog.launch_count - before <= 60, (sty, prog.launch_count - before)
            assert got == orc.run(raw, "match", sty, False), (head, tail, sty, got)
        del dt, raw
    prog.close()


def test_stream_large_dfa_that_forgets_nothing(gpu_red, redlib, oracle_cls):
    """A counter modulo 10 (more states than a chunk takes candidates, and no
    byte narrows them down): no chunk has candidates, and the window is resolved
    in order by the repair launch - slow, bounded, exact."""
    p = redlib.Parser()
    p.add(b"(" + b"(a|b)" * 10 + b")*", 1, 0)
    blob = redlib.compileToSerialized(p)
    rng = np.random.RandomState(12)
    for n in (100000, 100005):
        body = bytes(rng.choice(list(b"ab"), size=n).astype(np.uint8))
        check_stream(gpu_red, oracle_cls, blob, body, ("mod10", n))
