#!/usr/bin/env python
"""A/B timing of the check/match batch kernels on one GPU: one line per
configuration (environment knobs read at upload), HBM-resident corpus.

    python tools/hot_sweep.py [--lines N] [--style instant] [--program union64] CONFIG...

CONFIG is a comma-separated list of KNOB=VALUE (RED_B200_ prefix implied), e.g.
    HOT=0   HOT_WARPS=16,HOT_SLOTS=4   HOT_PROFILE=0
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from one_b200 import red  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=20_000_000)
    ap.add_argument("--style", default="instant")
    ap.add_argument("--program", default="union64")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("configs", nargs="*", default=[""])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    blob = bench.load_program_blob(a.program)
    pool = bench.host_pool(1234)
    text, starts, lens, n = bench.device_corpus(pool, a.lines, dev)
    line_bytes = int(pool[2].astype(np.uint64).sum()) * (n // bench.POOL_LINES)
    res = torch.empty(n, dtype=torch.int32, device=dev)
    ends = torch.empty(n, dtype=torch.int32, device=dev)
    ref = None
    for cfg in a.configs:
        keys = []
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ["RED_B200_" + k] = v
            keys.append("RED_B200_" + k)
        prog = red.Program(blob, device=0)
        for k in keys:
            del os.environ[k]
        for _ in range(3):
            prog.match_batch_device(text, starts, lens, "match", a.style, out=(res, None, ends))
        torch.cuda.synchronize()
        h0, l0 = prog.hot_launch_count, prog.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            prog.match_batch_device(text, starts, lens, "match", a.style, out=(res, None, ends))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        key = (int(res.sum().item()), int(ends.sum().item()))
        if ref is None:
            ref = key
        print("%-44s %8.3f ms %8.1f GB/s  launches/step %.1f (hot %.1f)  %s" % (
            cfg or "(default)", ms, line_bytes / ms / 1e6, (prog.launch_count - l0) / a.steps,
            (prog.hot_launch_count - h0) / a.steps, "ok" if key == ref else "DIFFERENT RESULTS"), flush=True)
        if "HOT_DEBUG=1" in cfg:
            import ctypes as C
            out = (C.c_uint64 * 16)()
            red.lib().red_b200_program_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
            red.lib().red_b200_program_debug_counters(prog._h, out)
            d = [int(x) for x in out]
            nb, nt = max(d[3], 1), max(d[8], 1)
            print("    producers: %.0f cycles per batch (index %.0f, slot wait %.0f, work %.0f); "
                  "consumers: %.0f cycles per tile (wait %.0f, walk %.0f, write-back %.0f); %d batches %d tiles" % (
                      d[0] / nb, d[1] / nb, d[2] / nb, (d[0] - d[1] - d[2]) / nb,
                      d[4] / nt, d[5] / nt, d[6] / nt, d[7] / nt, d[3], d[8]), flush=True)
        prog.close()


if __name__ == "__main__":
    main()
