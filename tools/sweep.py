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
    cfgs = os.environ.get("SWEEP", "v0,v3,v1").split(",")
    for c in cfgs:
        env = {"RED_B200_THREADS": "1024", "RED_B200_CTAS_PER_SM": "1", "RED_B200_SIMPLE_WALK": "0",
               "RED_B200_VARIANT": "0", "RED_B200_LINES_PER_LANE": "0"}
        for kv in c.split("+"):
            if kv == "v0":
                env["RED_B200_VARIANT"] = "0"
            elif kv == "v1":
                env["RED_B200_VARIANT"] = "1"
            elif kv == "v2":
                env["RED_B200_VARIANT"] = "2"
            elif kv in ("v4",):
                env["RED_B200_VARIANT"] = kv[1:]
            elif kv == "v3":
                env["RED_B200_VARIANT"] = "3"
            elif kv.startswith("T"):
                env["RED_B200_LINES_PER_LANE"] = kv[1:]
            elif kv.startswith("thr"):
                env["RED_B200_THREADS"] = kv[3:]
            elif kv == "nobyte":
                env["RED_B200_NO_BYTE_ROWS"] = "1"
            elif kv == "simple":
                env["RED_B200_SIMPLE_WALK"] = "1"
        env["name"] = c
        envs.append(env)
    extra = os.environ.get("SWEEP_EXTRA")
    for env in envs:
        os.environ.pop("RED_B200_NO_BYTE_ROWS", None)
        os.environ.update({k: v for k, v in env.items() if k != 'name'})
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
                print("%-16s style=%s start=%d : %.3f ms  %.1f GB/s" % (
                    env["name"], style, want_start, ms, nbytes / ms / 1e6), flush=True)
            except Exception as ex:  # noqa: BLE001
                print(env, style, "failed", ex, flush=True)
                torch.cuda.synchronize()
        prog.close()


if __name__ == "__main__":
    main()
