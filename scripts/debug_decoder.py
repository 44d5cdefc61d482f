"""Diagnostic (GPU): per-tensor error of the decoder chain (score heads, upscore2 / upscore_pool4 with their skip adds,
logits) against the fp64 oracle at a small size, with the location of the largest error."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import fcn8s_oracle as oracle  # noqa: E402
from fcn8s_tensorflow_b200 import ops  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402

C, N, H, W = 5, 2, 64, 96
weights = oracle.init_weights(C, seed=2, decoder_std_scale=10.0)
images, labels = oracle.synthetic_batch(N, H, W, C, seed=0)
dev = torch.device("cuda", 0)
with torch.no_grad():
    logits, inter = oracle.forward(weights, images, dtype=torch.float64, return_intermediates=True)
for precision in sys.argv[1:] or ["fp32"]:
    e = Engine(C, precision=precision, device=dev)
    e.load_weights(weights)
    x = torch.from_numpy(images).to(dev)
    got_logits = e.forward(x, train=True).double().cpu()
    torch.cuda.synchronize()
    A = e._arena(N, H, W)
    print("==== %s" % precision)
    names = {"pool3": "pool3", "pool4": "pool4", "fc7": "fc7", "s3": "s_h3", "s4": "s_h4", "s7": "s_h7", "f4": "f4",
             "f3": "f3"}
    for rname, aname in names.items():
        ref = inter[rname]
        t = A[aname]
        if aname in ("f4", "f3") or aname.startswith("s_"):
            got = ops.from_pair(t).double().cpu()[..., :ref.shape[-1]]
        else:
            got = (ops.from_pair(t) if e.pair else t).double().cpu()
        d = (got - ref).abs()
        i = np.unravel_index(int(d.argmax()), d.shape)
        print("%-6s max-rel %.3e at %s (got %.6g ref %.6g)  rms-rel %.3e" % (
            rname, d.max().item() / ref.abs().max().item(), tuple(int(v) for v in i), got[i].item(), ref[i].item(),
            d.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item()))
    d = (got_logits - logits).abs()
    i = np.unravel_index(int(d.argmax()), d.shape)
    print("logits max-rel %.3e at %s  rms-rel %.3e" % (d.max().item() / logits.abs().max().item(),
                                                        tuple(int(v) for v in i),
                                                        d.pow(2).mean().sqrt().item() / logits.pow(2).mean().sqrt().item()))
    # error per output row / column of the logits (border effects show up here)
    print("  max err by y:", ["%.0e" % v for v in d.amax(dim=(0, 2, 3)).tolist()][:40])
    print("  max err by x:", ["%.0e" % v for v in d.amax(dim=(0, 1, 3)).tolist()][:40])
