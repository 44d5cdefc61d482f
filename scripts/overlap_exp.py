"""Side-stream overlap experiments on the c2 / c3 training step (one GPU): background Adam of the [decoder | fc7 | fc6]
prefix under the conv backward (Engine.adam_overlap), max-pool backward under the filter-gradient GEMM
(Engine.pool_overlap), each with static / dynamic tile scheduling.  For every variant: the state after the first step
against the baseline variant (everything but the atomically accumulated bias gradients must agree bit for bit), the
loss after five steps, and the CUDA-event time of graph-replayed steps.

    python scripts/overlap_exp.py [steps]

Needs profiles/r02_overlap_experiment.patch applied (the experiment was measured slower and is not in the product:
profiles/r02_overlap_experiment.md).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcn8s_tensorflow_b200 import _capi as capi  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402
from fcn8s_tensorflow_b200.fcn8s import synthetic_weights  # noqa: E402

VARIANTS = [("base", 0, 0, 0), ("adam", 1, 0, 0), ("adam+dyn", 1, 0, 1), ("adam+pool", 1, 1, 0), ("pool", 0, 1, 0),
            ("base again", 0, 0, 0)]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = capi.load()
    C, H, W = 20, 512, 1024
    weights = synthetic_weights(C)
    rng = np.random.default_rng(0)
    results = []
    for precision, per in (("fp32", 4), ("bf16", 4), ("bf16", 2)):
        x = torch.from_numpy(rng.integers(0, 256, size=(per, H, W, 3), dtype=np.uint8)).to(dev)
        y = torch.from_numpy(np.eye(C, dtype=np.uint8)[rng.integers(0, C, size=(per, H, W))]).to(dev)
        eng = Engine(C, precision=precision, device=dev)
        base = None
        for name, adam, pool, dyn in VARIANTS:
            eng.load_weights(weights)
            eng.adam_m.zero_()
            eng.adam_v.zero_()
            eng.global_step = 0
            eng.repack()
            eng._graphs.clear()
            eng._warm.clear()
            eng.adam_overlap = bool(adam)
            eng.pool_overlap = bool(pool)
            lib.fcn8_debug_set(10, dyn)
            lib.fcn8_debug_set(12, 5 if pool else 16)
            eng.train_step(x, y, 1e-4, keep_prob=0.5)
            torch.cuda.synchronize()
            snap = [t.clone() for t in (eng.params, eng.adam_m, eng.adam_v, eng.w_hi)]
            rec = dict(precision=precision, per_gpu=per, variant=name)
            if base is None:
                base = snap
            else:
                nb = eng.bias_block
                rec["step1_nonbias_bit_equal"] = all(bool(torch.equal(a[:nb], b[:nb])) for a, b in zip(snap, base))
                rec["step1_mismatching_elements"] = int(sum(int((a[:nb] != b[:nb]).sum().item())
                                                            for a, b in zip(snap, base)))
                rec["step1_bias_m_max_rel"] = float(((snap[1][nb:] - base[1][nb:]).abs().max() /
                                                     base[1][nb:].abs().max()).item())
            for _ in range(4):
                eng.train_step(x, y, 1e-4, keep_prob=0.5)
            torch.cuda.synchronize()
            rec["loss_after_5"] = eng.loss_value(x.shape)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(2):
                a.record()
                for _ in range(steps):
                    eng.train_step(x, y, 1e-4, keep_prob=0.5)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / steps)
            rec["ms_per_step"] = best
            rec["bg_chunks_claimed_sms"] = int((eng.adam_ctl[2:] > 0).sum().item())
            print(json.dumps(rec))
            sys.stdout.flush()
            results.append(rec)
        lib.fcn8_debug_set(10, 0)
        lib.fcn8_debug_set(12, 16)
        del eng, base, snap
        torch.cuda.empty_cache()
    return 0


if __name__ == "__main__":
    sys.exit(main())
