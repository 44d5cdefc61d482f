"""Multi-GPU data-parallel parity on real devices (needs >= 2 GPUs; the world-size-2 gloo test in
tests/test_host_cpu.py covers the host logic on CPU): scripts/dp_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_gradients_match_single_gpu_and_replicas_stay_identical(cuda_device):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "DP_CHECK_OK" in r.stdout
