#!/usr/bin/env python
"""Small fixed workload over the paths around the batch matcher (line index,
long stream, compaction, sliding verbs) - run plainly for timings, and under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from one_b200 import red  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    lines = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    dev = torch.device("cuda", 0)
    pool = bench.host_pool(1234)
    text, starts, lens, n = bench.device_corpus(pool, lines, dev)
    gb = text.numel() / 1e9
    ms = timed(lambda: red.split_lines_device(text, capacity=n + 1))
    print("line index %.2f GB: %.3f ms %.0f GB/s" % (gb, ms, gb / ms * 1e3), flush=True)
    stream_gb = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    reps = max(1, int(stream_gb / gb))
    big = text.repeat(reps) if reps > 1 else text
    bgb = big.numel() / 1e9
    for name in ("union64", "single_auto"):
        prog = red.Program(bench.load_program_blob(name))
        for sty in ("last", "full", "instant"):
            ms = timed(lambda: prog.match_stream_device(big, "match", sty))
            print("stream %.2f GB %-11s %-8s: %.3f ms %.0f GB/s -> %s" % (
                bgb, name, sty, ms, bgb / ms * 1e3, tuple(prog.match_stream_device(big, "match", sty))), flush=True)
        prog.close()
    del big
    res = torch.empty(n, dtype=torch.int32, device=dev)
    ends0 = torch.empty(n, dtype=torch.int32, device=dev)
    nbytes = int(lens.sum().item())
    # every style on the union64 program (bench.py times Instant)
    prog = red.Program(bench.load_program_blob("union64"))
    for sty in ("instant", "first", "tangent", "last", "full"):
        ms = timed(lambda: prog.match_batch_device(text, starts, lens, "match", sty, out=(res, None, ends0)))
        print("union64 match/%-8s: %.3f ms %.0f GB/s" % (sty, ms, nbytes / ms / 1e6), flush=True)
    prog.close()
    # BASELINE configs[2]: 2000 keywords, table in global memory; hot rows mirrored in shared memory or not
    from one_b200 import synth as _synth
    p = red.Parser()
    for pat, r, f, lang in _synth.rules_keywords(2000):
        p.addAs(lang, pat, r, f)
    t0 = time.time()
    big_blob = red.compileToSerialized(p)
    print("compile 2000 keywords: %.1f s" % (time.time() - t0), flush=True)
    for hot in ("0", None):
        if hot is None:
            os.environ.pop("RED_B200_HOT_ROWS_BYTES", None)
        else:
            os.environ["RED_B200_HOT_ROWS_BYTES"] = hot
        prog = red.Program(big_blob)
        for sty in ("instant", "first", "last"):
            ms = timed(lambda: prog.match_batch_device(text, starts, lens, "match", sty, out=(res, None, ends0)))
            print("keywords2000 (%d states, global table, hot rows %s) match/%-8s: %.3f ms %.0f GB/s" % (
                prog.info["state_count"], "off" if hot == "0" else "on", sty, ms, nbytes / ms / 1e6), flush=True)
        prog.close()
    os.environ.pop("RED_B200_HOT_ROWS_BYTES", None)
    prog = red.Program(bench.load_program_blob("union64"))
    prog.match_batch_device(text, starts, lens, "match", "instant", out=(res, None, None))
    ms = timed(lambda: red.compact_matches_device(res))
    print("compaction of %d results: %.3f ms" % (n, ms), flush=True)
    prog.close()
    # sliding verbs: the same 64 phrases with tight ends (a loose-start program
    # makes search() quadratic by definition, in the reference as well)
    from one_b200 import synth
    p = red.Parser()
    for i, ph in enumerate(synth.PHRASES):
        p.addAs(synth.LANG_RAW, ph, i + 1, 0)
    prog = red.Program(red.compileToSerialized(p))
    m = min(n, 2_000_000)
    ends = torch.empty(m, dtype=torch.int32, device=dev)
    nb = int(lens[:m].sum().item())
    for verb in ("search", "scan"):
        ms = timed(lambda: prog.match_batch_device(text, starts[:m], lens[:m], verb, "last",
                                                   out=(res[:m], ends if verb == "search" else None, ends)), reps=2)
        print("%s/last (64 tight phrases) on %d lines: %.3f ms %.1f GB/s, %d hits" % (
            verb, m, ms, nb / ms / 1e6, int((res[:m] > 0).sum().item())), flush=True)
    # replace(): every phrase -> "<X>", all occurrences, rewritten lines packed with newlines
    ms = timed(lambda: prog.replace_batch_device(text, starts[:m], lens[:m], b"<X>", 1 << 40, "last", 10), reps=2)
    out, ostarts, olens, counts = prog.replace_batch_device(text, starts[:m], lens[:m], b"<X>", 1 << 40, "last", 10)
    print("replace/last (64 tight phrases) on %d lines: %.3f ms %.1f GB/s in, %d substitutions, %.1f MB out" % (
        m, ms, nb / ms / 1e6, int(counts.sum().item()), out.numel() / 1e6), flush=True)
    # StatefulMatcher: every line fed as two segments with the state carried
    half = (lens[:m] // 2).to(torch.int32)
    rest = lens[:m] - half
    st = torch.full((m,), -1, dtype=torch.int32, device=dev)   # RED_B200_STATE_FRESH
    r2 = torch.empty(m, dtype=torch.int32, device=dev)

    def feed():
        st.fill_(-1)
        prog.advance_batch_device(text, starts[:m], half, st, r2)
        prog.advance_batch_device(text, starts[:m] + half.to(torch.int64), rest, st, r2)
    ms = timed(feed, reps=2)
    print("advance (2 segments per line, state carried) on %d lines: %.3f ms %.1f GB/s" % (m, ms, nb / ms / 1e6), flush=True)
    prog.close()


def stream16():
    """BASELINE configs[3]: one 16 GB stream."""
    dev = torch.device("cuda", 0)
    pool = bench.host_pool(1234)
    text, starts, lens, n = bench.device_corpus(pool, 5_000_000, dev)
    reps = int(16e9 / text.numel()) + 1
    big = text.repeat(reps)
    del text
    gb = big.numel() / 1e9
    prog = red.Program(bench.load_program_blob("union64"))
    for sty in ("last", "full", "instant"):
        ms = timed(lambda: prog.match_stream_device(big, "match", sty), reps=2)
        print("stream %.2f GB union64 %-8s: %.3f ms %.0f GB/s -> %s" % (
            gb, sty, ms, gb / ms * 1e3, tuple(prog.match_stream_device(big, "match", sty))), flush=True)
    prog.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stream16":
        stream16()
        sys.exit(0)
    main()
