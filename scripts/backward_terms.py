"""Does the backward pass of the fp32-equivalent mode need all three bf16 products per algorithmic product?

For Engine(precision="fp32", backward_terms=3 | 2 | 1): gradient error (rel-L2 per variable and over the flat buffer)
against the fp64 CPU oracle on ONE 512x1024x20 image with keep_prob 1 (identical masks), and the c2 step time
(4 images, keep_prob 0.5, CUDA graph replay).  The forward pass (logits, loss) is the 3-product path in every row.
Writes a markdown table to the path given as argv[1] (default gpurun_out/backward_terms.md).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402
from oracle import fcn8s_oracle as oracle  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/backward_terms.md"
    dev = torch.device("cuda", 0)
    C, H, W = 20, 512, 1024
    w = oracle.init_weights(C, seed=2, decoder_std_scale=10.0)
    images, labels = oracle.synthetic_batch(4, H, W, C, seed=9)
    torch.set_num_threads(os.cpu_count() or 1)
    loss, logits, grads = oracle.loss_and_grads(w, images[:1], labels[:1], dtype=torch.float64)
    flat_ref = torch.cat([g.reshape(-1) for g in grads.values()])
    rows = []
    for bt in (3, 2, 1):
        e = Engine(C, precision="fp32", device=dev, backward_terms=bt)
        e.load_weights(w)
        x1 = torch.from_numpy(images[:1]).to(dev)
        y1 = torch.from_numpy(labels[:1].view(np.uint8)).to(dev)
        e.loss_and_backward(x1, y1, keep_prob=1.0)
        torch.cuda.synchronize()
        got = e.grad_dict()
        errs = {k: float((got[k].double() - v).norm() / v.norm()) for k, v in grads.items()}
        flat = torch.cat([got[k].double().reshape(-1) for k in grads])
        flat_err = float((flat - flat_ref).norm() / flat_ref.norm())
        cos = float(torch.dot(flat, flat_ref) / (flat.norm() * flat_ref.norm()))
        logit_err = float((e._arena(1, H, W)["logits"].double().cpu() - logits).abs().max() / logits.abs().max())
        worst = max(errs.items(), key=lambda t: t[1])
        x = torch.from_numpy(images).to(dev)
        y = torch.from_numpy(labels.view(np.uint8)).to(dev)
        for _ in range(5):
            e.train_step(x, y, 1e-4, keep_prob=0.5)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            e.train_step(x, y, 1e-4, keep_prob=0.5)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        rows.append((bt, logit_err, flat_err, 1.0 - cos, worst[0], worst[1], ms, 4e3 / ms))
        print(rows[-1], flush=True)
        del e
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        f.write("Backward products per algorithmic product in the fp32-equivalent mode (forward always 3): gradient error "
                "vs the fp64 CPU oracle on one 512x1024x20 image (keep_prob 1), and the c2 step (4 images, keep_prob 0.5).\n\n")
        f.write("| backward_terms | logits max-rel | flat gradient rel-L2 | 1 - cos(flat gradient) | worst variable | its "
                "rel-L2 | ms / step | img/s |\n|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %d | %.2e | %.2e | %.2e | %s | %.2e | %.2f | %.1f |\n" % r)
    print(open(out).read())


if __name__ == "__main__":
    main()
