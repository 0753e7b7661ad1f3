"""north_star's multi-GPU path on hardware, when the box has two GPUs: one
corpus cut by shard.byte_balanced_ranges, one process per GPU (torchrun, NCCL),
every rank matches its range through the C ABI, the matched line numbers are
gathered and checked against the CPU matcher - bench.py's own code path, on a
small corpus.  (The N > 1 host logic alone runs on CPU with gloo in
test_shard_gloo.py.)"""

import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_one_corpus(gpu_red):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--steps", "2", "--warmup", "3", "--lines", "600000", "--e2e-lines", "400000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["n_gpus"] == 2
    chk = out["gathered_check"]
    assert chk["first_copy_identical"] and chk["in_line_order"]
    assert chk["matched_lines_expected"] == chk["matched_lines_gathered"] > 0
    assert out["strong_scaling"]["lines_total"] == 600000
    assert out["config"]["lines_total"] == 1200000
    assert out["hot_kernel_launches"] == out["gpu_launches"] == 2
