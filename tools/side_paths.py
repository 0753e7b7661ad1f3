#!/usr/bin/env python
"""scan / search / replace on one GPU: the 64 phrases with tight ends over 2 M
corpus lines (a loose-start program makes search() quadratic by definition, in
the reference as well).  python tools/side_paths.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from one_b200 import red, synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    pool = bench.host_pool(1234)
    m = int(os.environ.get("SIDE_LINES", "2000000"))
    text, starts, lens, n = bench.device_corpus(pool, m, dev)
    p = red.Parser()
    for i, ph in enumerate(synth.PHRASES):
        p.addAs(synth.LANG_RAW, ph, i + 1, 0)
    prog = red.Program(red.compileToSerialized(p))
    res = torch.empty(n, dtype=torch.int32, device=dev)
    ends = torch.empty(n, dtype=torch.int32, device=dev)
    nb = int(lens.sum().item())
    for verb in ("search", "scan"):
        for sty in ("instant", "last"):
            ms = timed(lambda: prog.match_batch_device(text, starts, lens, verb, sty,
                                                       out=(res, ends if verb == "search" else None, ends)))
            print("%s/%-8s (64 tight phrases) on %d lines: %8.3f ms %7.1f GB/s, %d hits" % (
                verb, sty, n, ms, nb / ms / 1e6, int((res > 0).sum().item())), flush=True)
    for sty in ("instant", "last"):
        ms = timed(lambda: prog.replace_batch_device(text, starts, lens, b"<X>", 1 << 40, sty, 10))
        out, ostarts, olens, counts = prog.replace_batch_device(text, starts, lens, b"<X>", 1 << 40, sty, 10)
        print("replace/%-8s (64 tight phrases) on %d lines: %8.3f ms %7.1f GB/s in, %d substitutions, %.1f MB out" % (
            sty, n, ms, nb / ms / 1e6, int(counts.sum().item()), out.numel() / 1e6), flush=True)
    prog.close()


if __name__ == "__main__":
    main()
