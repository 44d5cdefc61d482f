"""Diagnostic (GPU): per-layer error of the fp32-equivalent modes against the fp64 oracle at the benchmark's image size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import fcn8s_oracle as oracle  # noqa: E402
from fcn8s_tensorflow_b200 import ops  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402

C, N, H, W = 20, 1, 512, 1024
weights = oracle.init_weights(C, seed=2, decoder_std_scale=10.0)
images, labels = oracle.synthetic_batch(N, H, W, C, seed=7)
dev = torch.device("cuda", 0)
with torch.no_grad():
    logits, inter = oracle.forward(weights, images, dtype=torch.float64, return_intermediates=True)
for precision in sys.argv[1:] or ["fp32"]:
    e = Engine(C, precision=precision, device=dev)
    e.load_weights(weights)
    x = torch.from_numpy(images).to(dev)
    e.forward(x, train=True)   # train=True: the pre-pool activations are stored too
    torch.cuda.synchronize()
    A = e._arena(N, H, W)
    print("==== %s" % precision)
    for name, ref in inter.items():
        if name not in A:
            continue
        got = A[name]
        got = ops.from_pair(got) if e.pair and got.dtype == torch.bfloat16 else got
        got = got.double().cpu()[..., :ref.shape[-1]]
        d = got - ref
        print("%-8s max-rel %.3e  rms-rel %.3e  mean signed (got-ref)/rms(ref) %+.3e" % (
            name, d.abs().max().item() / ref.abs().max().item(), d.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item(),
            d.mean().item() / ref.pow(2).mean().sqrt().item()))
    got = A["logits"].double().cpu()
    d = got - logits
    print("logits   max-rel %.3e  rms-rel %.3e" % (d.abs().max().item() / logits.abs().max().item(),
                                                   d.pow(2).mean().sqrt().item() / logits.pow(2).mean().sqrt().item()))
