import base64
import json
import os
import sys
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def unb64(s):
    return base64.b64decode(s)


def unz(s):
    return zlib.decompress(base64.b64decode(s))


def rules_from_json(rules):
    return [(unb64(r[0]), r[1], r[2], r[3]) for r in rules]


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "reference_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def vectors():
    with open(os.path.join(GOLDEN, "ref_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_cls():
    """red_oracle.c (built on demand with gcc).  Test infrastructure only."""
    from oracle import loader
    loader.build(ref=False)
    return loader.Oracle


@pytest.fixture(scope="session")
def ref_cls():
    """The compiled reference, when oracle/_ref/libred_ref.so is present."""
    from oracle import loader
    if not loader.have_ref():
        pytest.skip("oracle/_ref/libred_ref.so not built (needs /root/reference)")
    return loader.RefLib


@pytest.fixture(scope="session")
def redlib():
    """The product library (built on demand with nvcc; no GPU needed to build)."""
    from one_b200 import build, red
    build.build()
    return red


@pytest.fixture(scope="session")
def gpu_red(redlib):
    if redlib.device_count() < 1:
        pytest.fail("-m gpu tests need an sm_100 GPU and found none")
    return redlib


@pytest.fixture(scope="session")
def extra_vectors():
    """replace() / StatefulMatcher outputs of the reference (tools/gen_golden.py)."""
    with open(os.path.join(GOLDEN, "ref_extra.json")) as f:
        return json.load(f)
