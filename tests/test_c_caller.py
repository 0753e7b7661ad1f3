"""The C ABI from plain C: tests/c_caller/caller.c is compiled with gcc against
include/red_b200.h, linked with libred_b200.so and run on the GPU box - upload,
the batch call, single texts and replace() checked against the oracle's
answers.  (tests/test_integration_binding.py, the C++ binding of
INTEGRATION.md, needs the reference's headers and does not run there.)"""

import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, unz

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c_caller", "caller.c")


def build_caller(redlib, out_dir):
    lib = redlib.lib()._name
    exe = os.path.join(out_dir, "caller")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), SRC,
                    "-o", exe, "-L", os.path.dirname(lib), "-lred_b200",
                    "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return exe


def test_c_caller_compiles_and_links(redlib, tmp_path):
    """CPU: the header is valid C99 and the library exports what the caller uses."""
    exe = build_caller(redlib, str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_caller_runs(gpu_red, vectors, oracle_cls, tmp_path):
    from one_b200 import synth
    exe = build_caller(gpu_red, str(tmp_path))
    blob = unz(next(c for c in vectors["cases"] if c["name"] == "union64")["program"])
    orc = oracle_cls(blob)
    lines = synth.make_lines(5000, seed=77)
    lines[0] = b"sshd[1]: Failed password for root and again Failed password for root"
    text, starts, lens = synth.pack_lines(lines)
    (tmp_path / "prog.reda").write_bytes(blob)
    (tmp_path / "lines.txt").write_bytes(text.tobytes())
    res, ms, me = orc.batch(text, starts, lens, "match", "last")
    rows = ["%d %d %d" % t for t in zip(res.tolist(), ms.tolist(), me.tolist())]
    for ln in lines[:16]:
        rows.append("%d %d %d" % tuple(orc.run(ln, "match", "instant")))
    count, out = orc.replace(lines[0], b"<X>", (1 << 64) - 1, "last")
    rows.append("%d %d" % (len(out), count))
    (tmp_path / "expected.txt").write_text("\n".join(rows) + "\n")
    r = subprocess.run([exe, str(tmp_path / "prog.reda"), str(tmp_path / "lines.txt"), str(tmp_path / "expected.txt")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK"), r.stdout
