"""Promoted accumulation (ConvGemmArgs::promo_kb) at the benchmark geometry: full-size (512x1024x20) logits error
against the fp64 CPU oracle for several chunk lengths (fcn8_debug_set(9, P): -1 = off, P > 0 = k-blocks per chunk) on
the four weight / input distributions of tests/test_gpu_engine.py, and the forward time of each setting.  Prints
max-rel and the signed mean error relative to rms(ref) (negative = the accumulators shrink)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
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
    chunks = [int(a) for a in sys.argv[1:]] or [-1, 24, 12, 6]
    label = lambda p: "unpromoted" if p < 0 else "%d k-blocks" % p   # noqa: E731
    print("| distribution | " + " | ".join(label(p) for p in chunks) + " |")
    print("|---|" + "---|" * len(chunks))
    times = {}
    for name, (w, img) in cases.items():
        with torch.no_grad():
            ref = oracle.forward(w, img, dtype=torch.float64)
        rms = ref.pow(2).mean().sqrt().item()
        row = []
        for p in chunks:
            lib.fcn8_debug_set(9, p)
            e = Engine(C, precision="fp32", device=dev)
            e.load_weights(w)
            x = torch.from_numpy(img).to(dev)
            got = e.forward(x).double().cpu()
            d = got - ref
            row.append("%.2e (%+.1e)" % (d.abs().max().item() / ref.abs().max().item(), d.mean().item() / rms))
            if name == "he_normal":
                x4 = x.expand(4, -1, -1, -1).contiguous()
                for _ in range(3):
                    e.forward(x4)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(10):
                    e.forward(x4)
                t1.record()
                torch.cuda.synchronize()
                times[p] = t0.elapsed_time(t1) / 10
            del e
            torch.cuda.empty_cache()
        print("| %s | %s |" % (name, " | ".join(row)), flush=True)
    print("| forward of 4 images, ms | " + " | ".join("%.3f" % times[p] for p in chunks) + " |")
    lib.fcn8_debug_set(9, 0)


if __name__ == "__main__":
    main()
