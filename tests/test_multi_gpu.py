"""Launches tests/mgpu/run_mgpu.py with torchrun on 2 GPUs (NCCL halo exchange, remap, 3D3V and 2D2V
simulations against the single-GPU results and the golden file).  Skipped on boxes with one GPU; the
outputs of the 2/4/8-GPU runs made with `gpurun --gpus N` are kept under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu", "run_mgpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
