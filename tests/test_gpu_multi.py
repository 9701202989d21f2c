"""Two GPUs, two processes (torchrun + NCCL): the target-sharded run reproduces the single-GPU results bit for bit.
bench.py asserts it once per run before timing (`multi_gpu_check`); this test drives that path on a small gas disk.
Skipped on boxes with a single GPU (the driver's 1/2/4/8-GPU scaling run exercises the same assertion at full size)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_slices_equal_single_gpu_bitwise():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", "disk400k", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--no-fp64"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and "bit for bit" in (line["multi_gpu_check"] or "")
    assert line["e2e"]["d2h_bytes_per_step"] == 400000 * (4 + 8 * 9)            # every result column of every particle comes back once
