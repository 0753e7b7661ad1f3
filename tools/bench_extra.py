#!/usr/bin/env python
"""Side measurements for DESIGN.md (GPU box): the other BASELINE.json configs.
  - configs[0]: single syslog regex, 'contained', per-line batch
  - configs[2]: large union DFA (global table), first/last styles
  - configs[3]: one long stream (chunked, speculated, stitched)
  - line index and grep pipeline on a device-resident blob
Prints one line per measurement."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from one_b200 import red, synth  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    lines = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    stream_gb = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    dev = torch.device("cuda", 0)
    pool = bench.host_pool(1234)
    text, starts, lens, n = bench.device_corpus(pool, lines, dev)
    nbytes = int(pool[2].astype(np.uint64).sum()) * (n // bench.POOL_LINES)
    res = torch.empty(n, dtype=torch.int32, device=dev)
    ends = torch.empty(n, dtype=torch.int32, device=dev)
    print("corpus: %d lines, %.2f GB of line bytes" % (n, nbytes / 1e9), flush=True)

    # configs[0]
    prog = red.Program(bench.load_program_blob("single_auto"))
    for sty in ("instant", "last"):
        ms = timed(lambda: prog.match_batch_device(text, starts, lens, "match", sty, out=(res, None, ends)))
        print("config0 single regex (%d states, byte rows) %-8s: %.3f ms %.0f GB/s" % (
            prog.info["state_count"], sty, ms, nbytes / ms / 1e6), flush=True)
    prog.close()

    # configs[2]: 2000 keywords -> global table
    p = red.Parser()
    for pat, r, f, lang in synth.rules_keywords(2000):
        p.addAs(lang, pat, r, f)
    t0 = time.time()
    blob = red.compileToSerialized(p)
    prog = red.Program(blob)
    info = prog.info
    print("config2 compile 2000 keywords: %.1f s, %d states, %d classes, table %.1f MB, mode %d" % (
        time.time() - t0, info["state_count"], info["class_count"], info["table_bytes"] / 1e6, info["table_mode"]), flush=True)
    for sty in ("first", "last"):
        ms = timed(lambda: prog.match_batch_device(text, starts, lens, "match", sty, out=(res, None, ends)))
        print("config2 large DFA (global table) %-8s: %.3f ms %.0f GB/s" % (sty, ms, nbytes / ms / 1e6), flush=True)
    prog.close()

    # configs[3]: one long stream
    reps = max(1, int(stream_gb * 1e9 / text.numel()))
    big = text.repeat(reps) if reps > 1 else text
    for name in ("union64", "single_auto"):
        prog = red.Program(bench.load_program_blob(name))
        for sty in ("last", "full", "instant"):
            t0 = time.perf_counter()
            oc = prog.match_stream_device(big, "match", sty)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t0 = time.perf_counter()
            oc = prog.match_stream_device(big, "match", sty)
            torch.cuda.synchronize()
            dt = min(dt, time.perf_counter() - t0)
            print("config3 stream %.1f GB %-11s %-8s: %.1f ms %.0f GB/s  -> %s" % (
                big.numel() / 1e9, name, sty, dt * 1e3, big.numel() / dt / 1e9, tuple(oc)), flush=True)
        prog.close()
    del big

    # line index on the device-resident blob
    def split():
        red.split_lines_device(text, capacity=n + 1)
    ms = timed(split, reps=2)
    print("line index (split at \\n) of %.2f GB: %.3f ms %.0f GB/s" % (text.numel() / 1e9, ms, text.numel() / ms / 1e6), flush=True)

    # sliding verbs on a slice
    prog = red.Program(bench.load_program_blob("union64"))
    m = min(n, 2_000_000)
    sl_bytes = int(lens[:m].sum().item())
    for verb in ("search", "scan"):
        ms = timed(lambda: prog.match_batch_device(text, starts[:m], lens[:m], verb, "last",
                                                   out=(res[:m], ends[:m] if verb == "search" else None, ends[:m])), reps=1)
        print("%s/last on %d lines: %.3f ms %.1f GB/s" % (verb, m, ms, sl_bytes / ms / 1e6), flush=True)
    prog.close()


if __name__ == "__main__":
    main()
