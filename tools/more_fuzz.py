"""More seeds for tests/test_gpu_fuzz.py (GPU box): python tools/more_fuzz.py"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from one_b200 import build, red
from oracle import loader
import test_gpu_fuzz as f
build.build(); loader.build(ref=False)
for seed in (404, 505, 606, 707, 808, 909, 1234, 4321):
    f.test_fuzz_random_programs(red, red, loader.Oracle, seed)
    print("seed", seed, "ok", flush=True)
