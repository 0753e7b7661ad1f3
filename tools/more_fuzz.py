"""More seeds for tests/test_gpu_fuzz.py (GPU box): python tools/more_fuzz.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from one_b200 import build, red  # noqa: E402
from oracle import loader  # noqa: E402
import test_gpu_fuzz as f  # noqa: E402

build.build()
loader.build(ref=False)
for seed in (404, 505, 606, 707, 808, 909, 1234, 4321):
    f.test_fuzz_random_programs(red, red, loader.Oracle, seed)
    print("seed", seed, "ok", flush=True)
