"""Diagnostic (GPU): per-layer activation agreement between the engine's reduced-precision modes and the
quantisation-aware oracle (oracle.forward(storage=...))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import fcn8s_oracle as oracle  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402

C, N, H, W = 5, 2, 64, 96
weights = oracle.init_weights(C, seed=2, decoder_std_scale=10.0)
images, labels = oracle.synthetic_batch(N, H, W, C, seed=0)
dev = torch.device("cuda", 0)
for precision in sys.argv[1:] or ["bf16"]:
    logits, inter = oracle.forward(weights, images, dtype=torch.float64, return_intermediates=True, storage=precision)
    e = Engine(C, precision=precision, device=dev)
    e.load_weights(weights)
    x = torch.from_numpy(images).to(dev)
    y = torch.from_numpy(labels.view(np.uint8)).to(dev)
    e.loss_and_backward(x, y, keep_prob=1.0)
    torch.cuda.synchronize()
    A = e._arena(N, H, W)
    print("==== %s" % precision)
    for name, ref in inter.items():
        if name not in A:
            continue
        got = A[name].double().cpu()
        if got.shape != ref.shape:
            got = got[..., :ref.shape[-1]]
        d = (got - ref).abs()
        mism = ((got > 0) != (ref > 0)).double().mean().item()
        print("%-8s max-rel %.3e  mean-rel %.3e  sign-mismatch %.2e  frac(|d|>0) %.3f" % (
            name, d.max().item() / ref.abs().max().item(), d.mean().item() / ref.abs().mean().item(), mism,
            (d > 0).double().mean().item()))
    got = A["logits"].double().cpu()
    print("logits   max-rel %.3e" % ((got - logits).abs().max().item() / logits.abs().max().item()))
