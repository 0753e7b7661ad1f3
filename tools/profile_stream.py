#!/usr/bin/env python
"""The long-stream matcher on one GPU: union64 over the corpus text repeated to
--gb gigabytes, every style; and small DFAs that never forget.  --once: one call
(style last) only, for an ncu launch list."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from one_b200 import red  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--program", default="union64")
    ap.add_argument("--regex", default=None, help="a loose-start regex instead of --program")
    a = ap.parse_args()
    pool = bench.host_pool(1234)
    tp = torch.from_numpy(pool[0].copy()).cuda()
    reps = int(a.gb * (1 << 30)) // tp.numel() + 1
    text = tp.repeat(reps)
    if a.regex:
        rp = red.Parser()
        rp.add(a.regex.encode(), 1, 2)
        prog = red.compile(rp)
        a.program = "regex"
    else:
        prog = red.Program(bench.load_program_blob(a.program))
    styles = ["last"] if a.once else ["instant", "first", "tangent", "last", "full"]
    for sty in styles:
        o = prog.match_stream_device(text, "match", sty)
        torch.cuda.synchronize()
        if a.once:
            break
        l0 = prog.launch_count
        t0 = time.perf_counter()
        for _ in range(3):
            o = prog.match_stream_device(text, "match", sty)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        print(a.program + " %-8s %6.2f GB  %8.3f ms  %8.1f GB/s  %4.1f launches/call  outcome %s" % (
            sty, text.numel() / 1e9, ms, text.numel() / ms / 1e6, (prog.launch_count - l0) / 3, tuple(o)), flush=True)
    prog.close()
    if a.once:
        return
    p = red.Parser()
    p.add(b"((a|b)(a|b))*", 1, 0)
    prog = red.compile(p)
    ab = torch.frombuffer(bytearray(b"ab" * 8192), dtype=torch.uint8).cuda().repeat(int(a.gb * (1 << 30)) // 16384)
    for sty in ("full", "last"):
        o = prog.match_stream_device(ab, "match", sty)
        torch.cuda.synchronize()
        l0 = prog.launch_count
        t0 = time.perf_counter()
        o = prog.match_stream_device(ab, "match", sty)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        print("parity  %-8s %6.2f GB  %8.3f ms  %8.1f GB/s  %4.1f launches/call  outcome %s" % (
            sty, ab.numel() / 1e9, ms, ab.numel() / ms / 1e6, prog.launch_count - l0, tuple(o)), flush=True)
    prog.close()


if __name__ == "__main__":
    main()
