"""INTEGRATION.md's C++ binding compiles against the reference's own headers,
links against the product library, and - on a host without a B200 - fails the
way the document says: RED_B200_ENODEVICE surfaces as a RedExcept, there is no
CPU fallback.  Needs /root/reference (headers) and oracle/_ref (the reference
compiled as a library); skipped where they are absent (the GPU box)."""

import os
import re
import subprocess

import pytest

from conftest import ROOT

REF = "/root/reference/quol/red"


def test_binding_compiles_links_and_reports_no_device(redlib, tmp_path):
    from oracle import loader
    if not os.path.isdir(os.path.join(REF, "include")) or not loader.have_ref():
        pytest.skip("reference headers / oracle/_ref not present")
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```cpp\n(.*?)```", doc, re.S).group(1)
    assert "class GpuMatcher" in code
    (tmp_path / "GpuMatcher.h").write_text(code)
    (tmp_path / "main.cpp").write_text(r'''
#include <cstdio>
#include <cstring>
#include "Compile.h"
#include "Parser.h"
#include "GpuMatcher.h"
using namespace zezax::red;
int main() {
  Parser p;
  p.add("ab*c", 1, 0);
  Executable rex = compile(p);
  Outcome cpu = match(rex, "xxabbc", styLast);   // the reference's own matcher still works
  std::printf("cpu %d\n", (int)cpu.result_);
  try {
    GpuMatcher gm(rex);
    Outcome oc = gm.match("abbc", 4, styLast);
    std::printf("gpu %d %zu %zu\n", (int)oc.result_, (size_t)oc.start_, (size_t)oc.end_);
    std::vector<Result> res;
    std::vector<uint32_t> ends;
    gm.matchLines("abc\nxyz\n", {0, 4}, {3, 3}, styInstant, res, ends);
    std::printf("lines %d %d\n", (int)res[0], (int)res[1]);
    return 0;
  }
  catch (const RedExcept &ex) {
    std::printf("except %s\n", ex.what());
    return 3;
  }
}
''')
    so_dir = os.path.dirname(redlib.lib()._name)
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    exe = str(tmp_path / "binding")
    gxx = "/opt/gcc/bin/g++" if os.path.exists("/opt/gcc/bin/g++") else "g++"
    cmd = [gxx, "-std=c++20", "-O1", "-Wall", "-iquote", os.path.join(REF, "include"),
           "-I", os.path.join(REF, "include"), "-I", os.path.join(ROOT, "include"), "-I", str(tmp_path),
           str(tmp_path / "main.cpp"), "-o", exe,
           "-L", ref_dir, "-lred_ref", "-L", so_dir, "-lred_b200",
           "-Wl,-rpath," + ref_dir, "-Wl,-rpath," + so_dir]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "cpu 0" in run.stdout or "cpu 1" in run.stdout, run.stdout + run.stderr
    if redlib.device_count() < 1:
        assert run.returncode == 3, run.stdout + run.stderr
        assert "CPU fallback" in run.stdout or "device" in run.stdout, run.stdout
    else:
        assert run.returncode == 0 and "gpu 1 0 4" in run.stdout and "lines 1 0" in run.stdout, run.stdout
