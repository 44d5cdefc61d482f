"""Diagnostic (GPU): per-layer error of the fp32-equivalent mode against the fp64 oracle at the benchmark's image size,
for one of the weight / input distributions of tests/test_gpu_engine.py (default: the He-normal one).

    python scripts/debug_fullsize.py [he_normal|reference_init|sparse_image|positive_weights] [promo_kb ...]
(promo_kb: fcn8_debug_set(9, .) -- 0 = library default, -1 = unpromoted accumulation, P > 0 = chunk length)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import fcn8s_oracle as oracle  # noqa: E402
from fcn8s_tensorflow_b200 import _capi, ops  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402
import test_gpu_engine as T  # noqa: E402

C, N, H, W = 20, 1, 512, 1024
case = sys.argv[1] if len(sys.argv) > 1 else "he_normal"
consts = [int(a) for a in sys.argv[2:]] or [0]
if case == "he_normal":
    weights, images = oracle.init_weights(C, seed=2, decoder_std_scale=10.0), oracle.synthetic_batch(N, H, W, C, seed=7)[0]
else:
    weights, images = T._distribution_case(case, C, H, W)
dev = torch.device("cuda", 0)
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    logits, inter = oracle.forward(weights, images, dtype=torch.float64, return_intermediates=True)
lib = _capi.load()
for c in consts:
    lib.fcn8_debug_set(9, c)
    e = Engine(C, precision="fp32", device=dev)
    e.load_weights(weights)
    x = torch.from_numpy(images).to(dev)
    got_logits = e.forward(x, train=True).double().cpu()   # train=True: the pre-pool activations are stored too
    torch.cuda.synchronize()
    A = e._arena(N, H, W)
    print("==== %s, promo_kb %s" % (case, "library default" if c == 0 else ("off" if c < 0 else str(c))))
    names = dict((k, k) for k in inter)
    names.update(s3="s_h3", s4="s_h4", s7="s_h7")
    for name, ref in inter.items():
        t = A.get(names[name])
        if t is None:
            continue
        got = ops.from_pair(t).double().cpu()[..., :ref.shape[-1]]
        d = got - ref
        rms = ref.pow(2).mean().sqrt().item()
        print("%-8s max-rel %.3e  rms-rel %.3e  mean signed (got-ref)/rms(ref) %+.3e   mean(ref)/rms(ref) %+.2f" % (
            name, d.abs().max().item() / ref.abs().max().item(), d.pow(2).mean().sqrt().item() / rms,
            d.mean().item() / rms, ref.mean().item() / rms))
    d = got_logits - logits
    print("logits   max-rel %.3e  rms-rel %.3e" % (d.abs().max().item() / logits.abs().max().item(),
                                                   d.pow(2).mean().sqrt().item() / logits.pow(2).mean().sqrt().item()))
    del e
    torch.cuda.empty_cache()
