#!/usr/bin/env python
"""Kernel timing sweep (GPU box): variants x styles on the union64 corpus.
Prints one line per configuration; used to pick launch defaults."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from one_b200 import red  # noqa: E402


def main():
    lines = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    progname = sys.argv[2] if len(sys.argv) > 2 else "union64"
    blob = bench.load_program_blob(progname)
    dev = torch.device("cuda", 0)
    pool = bench.host_pool(1234)
    text, starts, lens, n = bench.device_corpus(pool, lines, dev)
    nbytes = int(pool[2].astype(np.uint64).sum()) * (n // bench.POOL_LINES)
    res = torch.empty(n, dtype=torch.int32, device=dev)
    ends = torch.empty(n, dtype=torch.int32, device=dev)
    sts = torch.empty(n, dtype=torch.int32, device=dev)
    print("lines", n, "bytes %.2f GB" % (nbytes / 1e9), flush=True)
    envs = []
    for threads, ctas, simple in itertools.product([256, 512, 1024], [1, 2, 4], [0, 1]):
        if threads * ctas > 2048:
            continue
        envs.append({"RED_B200_THREADS": str(threads), "RED_B200_CTAS_PER_SM": str(ctas),
                     "RED_B200_SIMPLE_WALK": str(simple)})
    extra = os.environ.get("SWEEP_EXTRA")
    for env in envs:
        os.environ.update(env)
        try:
            prog = red.Program(blob)
        except Exception as ex:  # noqa: BLE001
            print(env, "upload failed", ex)
            continue
        for style, want_start in (("instant", False), ("last", False), ("full", False), ("last", True)):
            out = (res, sts if want_start else None, ends)
            try:
                for _ in range(2):
                    prog.match_batch_device(text, starts, lens, "match", style, out=out)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    prog.match_batch_device(text, starts, lens, "match", style, out=out)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 3
                print("thr=%s ctas=%s simple=%s style=%s start=%d : %.3f ms  %.1f GB/s" % (
                    env["RED_B200_THREADS"], env["RED_B200_CTAS_PER_SM"], env["RED_B200_SIMPLE_WALK"],
                    style, want_start, ms, nbytes / ms / 1e6), flush=True)
            except Exception as ex:  # noqa: BLE001
                print(env, style, "failed", ex, flush=True)
                torch.cuda.synchronize()
        prog.close()


if __name__ == "__main__":
    main()
