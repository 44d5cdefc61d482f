"""Measurement: how much of the implicit-GEMM kernels' time is the epilogue's global traffic?

Runs the c2 train step eagerly with per-launch CUDA events (ops.KernelTimer) three times per precision mode:
normal, with the conv epilogues' output stores skipped (fcn8_debug_set(3, 1)), and with their mask / residual loads
skipped as well (3).  Results of the last two are numerically meaningless; only the kernel times are read.
    python scripts/epilogue_cost.py [bf16 fp32]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fcn8s_tensorflow_b200 import _capi as capi  # noqa: E402
from fcn8s_tensorflow_b200 import ops  # noqa: E402
from fcn8s_tensorflow_b200.fcn8s import FCN8s, synthetic_weights  # noqa: E402


def main():
    os.environ["FCN8_GRAPHS"] = "0"
    lib = capi.load()
    dev = torch.device("cuda", 0)
    weights = synthetic_weights(bench.C, 2)
    images, labels = bench.synthetic_feed(bench.PER_GPU_BATCH, 1000)
    x = torch.from_numpy(images).to(dev)
    y = torch.from_numpy(labels.view("uint8")).to(dev)
    for precision in (sys.argv[1:] or ["bf16", "fp32"]):
        model = FCN8s(weights=weights, precision=precision, device=dev)
        eng = model.engine
        for _ in range(3):
            eng.train_step(x, y, 1e-4, keep_prob=0.5)
        for dbg, name in ((0, "normal"), (1, "no output stores"), (3, "no stores, no mask/residual loads")):
            lib.fcn8_debug_set(3, dbg)
            eng.train_step(x, y, 1e-4, keep_prob=0.5)
            torch.cuda.synchronize()
            timer = ops.KernelTimer()
            ops.TIMER = timer
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                eng.train_step(x, y, 1e-4, keep_prob=0.5)
            e1.record()
            torch.cuda.synchronize()
            ops.TIMER = None
            s = timer.summary()
            parts = ["%s %.3f ms (%.0f TF/s)" % (k, v["ms"] / 3, v["flops"] / v["ms"] / 1e9) for k, v in sorted(s.items())]
            print("%-5s %-34s step %.3f ms | %s" % (precision, name, e0.elapsed_time(e1) / 3, " | ".join(parts)), flush=True)
        lib.fcn8_debug_set(3, 0)
        model.close()


if __name__ == "__main__":
    main()
