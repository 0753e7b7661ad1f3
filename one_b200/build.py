"""Build libred_b200.so in-tree with nvcc for sm_100a.

    python -m one_b200.build [--force]

nvcc cross-compiles without a GPU.  The library links cudart statically and
has no dependency on torch or Python.
"""

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libred_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wextra",
          "-I", os.path.join(HERE, "..", "include")]


def sources():
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp"))
                  + glob.glob(os.path.join(CSRC, "compile", "*.cpp")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "compile", "*.h"))
                  + glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    return srcs, hdrs


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs, hdrs = sources()
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in srcs:
        obj = os.path.join(OUT_DIR, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
        objs.append(obj)
        if force or stale(obj, [src] + hdrs):
            cmd = [NVCC] + ARCH + COMMON + ["-c", src, "-o", obj]
            if src.endswith(".cu") and verbose:
                cmd += ["-Xptxas", "-v"]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            sys.stderr.write("nvcc failed on %s\n" % src)
            failed = True
    if failed:
        raise RuntimeError("build failed")
    if force or procs or stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
