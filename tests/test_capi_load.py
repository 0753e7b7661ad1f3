"""The C-ABI library: builds with nvcc without a GPU, loads, exports every
symbol include/red_b200.h declares, validates headers on the host, and fails
loudly (no CPU fallback) when no sm_100 device is present.  CPU only."""

import ctypes
import os
import re

import pytest

from conftest import ROOT, unz


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "red_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(red_b200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(redlib):
    lib = ctypes.CDLL(redlib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name


def test_abi_version(redlib):
    assert redlib.lib().red_b200_abi_version() == 1


def test_check_header_host_side(redlib, vectors):
    prog = unz(vectors["cases"][0]["program"])
    L = redlib.lib()

    def chk(b):
        msg = L.red_b200_check_header(ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b))
        return msg.decode() if msg else None

    assert chk(prog) is None
    assert chk(prog[:64]) == "Serialized DFA: header too short"
    assert chk(b"XEDA" + prog[4:]) == "Serialized DFA: bad magic number"
    bad = bytearray(prog); bad[-1] ^= 0x40
    assert chk(bytes(bad)) == "serialized DFA: checksum mismatch"
    bad = bytearray(prog); bad[6] = 9
    assert chk(bytes(bad)) == "Serialized DFA: unrecognized version"


def test_no_silent_cpu_fallback(redlib, vectors):
    """Without a usable GPU every compute entry point must raise."""
    if redlib.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is for CPU boxes")
    prog = unz(vectors["cases"][0]["program"])
    with pytest.raises(redlib.RedExceptNoDevice):
        redlib.Program(prog)


def test_bad_program_is_api_error(redlib):
    with pytest.raises(redlib.RedExceptApi):
        redlib.Program(b"not a dfa")
    with pytest.raises(redlib.RedExceptApi):
        redlib.Program(b"")


def test_product_does_not_touch_oracle():
    """one_b200/ must not import, link or load anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "one_b200")):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in txt.lower().replace("oracle/_ref", "").replace("oracle/", "ORACLEDIR") \
                    or "import oracle" not in txt, fn
                assert "from oracle" not in txt and "import oracle" not in txt, fn
                assert "libred_oracle" not in txt and "libred_ref" not in txt, fn
