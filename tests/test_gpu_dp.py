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


def test_engine_on_a_device_that_is_not_the_current_one(cuda_device):
    """Engine(device='cuda:1') while the current device is 0: every launch must go to device 1's stream and
    the results must equal those of the same engine on device 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import numpy as np
    from fcn8s_tensorflow_b200.engine import Engine
    from fcn8s_tensorflow_b200.fcn8s import synthetic_weights
    torch.cuda.set_device(0)
    w = synthetic_weights(5, 2, decoder_std_scale=10.0)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(1, 64, 96, 3), dtype=np.uint8)
    lab = np.eye(5, dtype=np.uint8)[rng.integers(0, 5, size=(1, 64, 96))]
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        e = Engine(5, precision="fp32", device=dev)
        e.load_weights(w)
        logits = e.forward(torch.from_numpy(img).to(dev)).cpu()
        e.train_step(torch.from_numpy(img).to(dev), torch.from_numpy(lab).to(dev), 1e-4, keep_prob=1.0)
        assert torch.cuda.current_device() == 0
        torch.cuda.synchronize(dev)
        outs.append((logits, e.loss_value((1, 64, 96))))
    assert torch.equal(outs[0][0], outs[1][0])          # per-tile arithmetic does not depend on the device
    # the loss is a sum of per-CTA partial sums added with atomics: their order is not fixed (dynamic tile scheduling)
    assert abs(outs[0][1] - outs[1][1]) <= 1e-6 * abs(outs[0][1])
