"""Calibration of the round-toward-zero compensation constant (ConvGemmArgs::rz_c, kRzBiasPerMma) at the benchmark
geometry: full-size (512x1024x20) logits error against the fp64 CPU oracle for several constants (fcn8_debug_set(7, c *
1e10)) on the four weight / input distributions of tests/test_gpu_engine.py.  Prints max-rel and the signed mean error
relative to rms(ref) (negative = the accumulators still shrink, positive = over-compensated)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fcn8s_tensorflow_b200 import _capi  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402
from oracle import fcn8s_oracle as oracle  # noqa: E402
import test_gpu_engine as T  # noqa: E402


def main():
    lib = _capi.load()
    dev = torch.device("cuda", 0)
    C, H, W = 20, 512, 1024
    torch.set_num_threads(os.cpu_count() or 1)
    cases = {"he_normal": (oracle.init_weights(C, seed=2, decoder_std_scale=10.0),
                           oracle.synthetic_batch(1, H, W, C, seed=7)[0])}
    for name in ("reference_init", "sparse_image", "positive_weights"):
        cases[name] = T._distribution_case(name, C, H, W)
    consts = [float(a) for a in sys.argv[1:]] or [0.0, 1.5e-8, 2.1e-8, 3.0e-8, 4.5e-8]
    print("| distribution | " + " | ".join("c = %.1e" % c for c in consts) + " |")
    print("|---|" + "---|" * len(consts))
    for name, (w, img) in cases.items():
        with torch.no_grad():
            ref = oracle.forward(w, img, dtype=torch.float64)
        rms = ref.pow(2).mean().sqrt().item()
        row = []
        for c in consts:
            if c == 0.0:
                lib.fcn8_debug_set(0, 1)
            else:
                lib.fcn8_debug_set(0, 0)
                lib.fcn8_debug_set(7, int(round(c * 1e10)))
            e = Engine(C, precision="fp32", device=dev)
            e.load_weights(w)
            got = e.forward(torch.from_numpy(img).to(dev)).double().cpu()
            d = got - ref
            row.append("%.2e (%+.1e)" % (d.abs().max().item() / ref.abs().max().item(), d.mean().item() / rms))
            del e
            torch.cuda.empty_cache()
        print("| %s | %s |" % (name, " | ".join(row)), flush=True)
    lib.fcn8_debug_set(0, 0)
    lib.fcn8_debug_set(7, 0)


if __name__ == "__main__":
    main()
