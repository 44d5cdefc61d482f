"""Measurement: where do the MMA warps of the generic implicit-GEMM kernels wait?

One eager c2 train step with fcn8_debug_buffer installed: every fcn8_conv_gemm / fcn8_wgrad_gemm launch gets a slot
in which each CTA's MMA warp records the cycles it spent in its tile loop, waiting for operand stages (full
barriers = the TMA feed is late) and waiting for a free accumulator (= the epilogue is late); the producer records
its wait for free stages (= the MMAs are the slow side).  conv_halo_kernel reports its activation-patch waits under "operands" (k-block = one patch) and its weight-stage waits
separately; wgrad_halo_kernel leaves its slot empty.
    python scripts/wait_profile.py [bf16|fp32]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FCN8_GRAPHS"] = "0"
import bench  # noqa: E402
from fcn8s_tensorflow_b200 import _capi as capi  # noqa: E402
from fcn8s_tensorflow_b200 import ops  # noqa: E402
from fcn8s_tensorflow_b200.fcn8s import FCN8s, synthetic_weights  # noqa: E402


def main():
    lib = capi.load()
    dev = torch.device("cuda", 0)
    cfg = bench.CONFIGS["c2"]
    weights = synthetic_weights(cfg["C"], 2)
    images, labels = bench.synthetic_feed(cfg, cfg["per_gpu"], 1000)
    x = torch.from_numpy(images).to(dev)
    y = torch.from_numpy(labels.view("uint8")).to(dev)
    slots = 160
    for precision in (sys.argv[1:] or ["bf16"]):
        model = FCN8s(weights=weights, precision=precision, device=dev)
        eng = model.engine
        for _ in range(3):
            eng.train_step(x, y, 1e-4, keep_prob=0.5)
        torch.cuda.synchronize()
        buf = torch.zeros((slots, 148, 8), dtype=torch.int64, device=dev)
        timer = ops.KernelTimer()
        ops.TIMER = timer
        lib.fcn8_debug_buffer(buf.data_ptr(), slots)
        eng.train_step(x, y, 1e-4, keep_prob=0.5)
        torch.cuda.synchronize()
        lib.fcn8_debug_buffer(None, 0)
        ops.TIMER = None
        recs = [(tag, fl, e0.elapsed_time(e1)) for tag, fl, e0, e1 in timer.records if tag in ("conv_gemm", "wgrad_gemm")]
        b = buf.cpu().double()
        print("%s: %d launches; columns: ms, TFLOP/s, CTAs, kcycles in the MMA loop (mean), %% waiting for operands, "
              "%% waiting for an accumulator, producer %% waiting for a free stage, cycles per k-block"
              % (precision, len(recs)))
        for i, (tag, fl, ms) in enumerate(recs):
            s = b[i]
            live = s[:, 0] > 0
            if not bool(live.any()):
                print("%3d %-10s %8.3f ms %7.0f TF/s   (wgrad_halo kernel, not instrumented)" % (i, tag, ms, fl / ms / 1e9))
                continue
            tot = s[live, 0]
            print("%3d %-10s %8.3f ms %7.0f TF/s  %3d CTAs %8.1f kcyc  operands %5.1f%%  accumulator %5.1f%%  "
                  "producer-stage %5.1f%%  %6.0f cyc/k-block  (halo kernels: weight stages %5.1f%%)  epilogue warp idle %5.1f%%"
                  % (i, tag, ms, fl / ms / 1e9, int(live.sum()), tot.mean() / 1e3, 100 * (s[live, 1] / tot).mean(),
                     100 * (s[live, 2] / tot).mean(), 100 * (s[live, 4] / tot).mean(),
                     (tot / s[live, 3].clamp(min=1)).mean(), 100 * (s[live, 5] / tot).mean(),
                     100 * (s[live, 6] / s[live, 7].clamp(min=1)).mean()))
        model.close()


if __name__ == "__main__":
    main()
