#!/usr/bin/env python
"""FCN-8s training throughput on B200 -- BASELINE.json's metric: train images/sec @1024x512, 20 classes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16|tf32] [--impl reference]

A step = one pass of the hot path over one synthetic Cityscapes-shaped batch: uint8 feed pre-processing, VGG-16
encoder, decoder, softmax-CE loss, full backward, (N>1: one NCCL all-reduce of the flat gradient buffer) and the fused
Adam update.  Workload = BASELINE configs[1]: 4 images of 512x1024x3 per GPU, 20 classes; weak scaling over N GPUs.
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference graph (oracle/) on the
host cores instead (TensorFlow 1.x cannot run here).
"""
import argparse
import re
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, C, PER_GPU_BATCH = 512, 1024, 20, 4
FWD_GFLOP_PER_IMAGE = 445.04
TRAIN_GFLOP_PER_IMAGE = 1333.31
METRIC = "FCN-8s train images/sec @1024x512 20-class"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm=d["hbm_gbs"],
                    source="MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback of B200_PROFILING.md")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_feed(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)
    ids = rng.integers(0, C, size=(n, H, W))
    labels = np.eye(C, dtype=bool)[ids]          # what convert_IDs_to_one_hot yields (bool one-hot)
    return images, labels


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_fn(weights_np, sample):
    """Returns (fn, images_per_step): fn() runs one train step of the CPU oracle on a bounded sample."""
    import torch
    from oracle import fcn8s_oracle as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    w = {k: torch.from_numpy(v.copy()) for k, v in weights_np.items()}
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v_ = {k: torch.zeros_like(v) for k, v in w.items()}
    h, wd = sample
    images, labels = synthetic_feed(1, 123)
    images, labels = images[:, :h, :wd], labels[:, :h, :wd]
    state = {"step": 0}

    def fn():
        loss, state["step"] = oracle.train_step(w, m, v_, state["step"], images, labels, 1e-4, keep_prob=1.0,
                                                dtype=torch.float32)
        return loss
    return fn, (h * wd) / float(H * W)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fcn8s_tensorflow_b200.fcn8s import synthetic_weights
    total = args.steps + args.warmup
    # ~12 s per full 512x1024 image on 8 cores; keep the whole run within a few minutes
    sample = (512, 1024) if total <= 12 else ((256, 512) if total <= 48 else (128, 256))
    fn, img_per_step = cpu_reference_step_fn(synthetic_weights(C, 2), sample)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = img_per_step * args.steps / dt
    desc = "1 image crop of %dx%d px per step (%.3f of a 512x1024 image), batch 1, torch-CPU fp32" % (
        sample[0], sample[1], img_per_step)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "FCN-8s train step 512x1024x3, 20 classes (BASELINE configs[1]); CPU restatement of "
                               "the reference TF1 graph (TensorFlow 1.x not installable here), bounded sample",
                   "sample": desc},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------- product arm
def time_engine(precision, weights, args, rank, local_rank, world, dev):
    """Device-resident timed loop + e2e loop for one precision mode. Returns a dict of numbers."""
    import torch
    import torch.distributed as tdist
    from fcn8s_tensorflow_b200 import _capi as capi
    from fcn8s_tensorflow_b200 import dist as fdist
    from fcn8s_tensorflow_b200 import ops
    from fcn8s_tensorflow_b200.fcn8s import FCN8s

    lib = capi.load()
    model = FCN8s(weights=weights, precision=precision, device=dev, data_parallel=world > 1)
    eng = model.engine
    images, labels = synthetic_feed(PER_GPU_BATCH, 1000 + rank)
    x = torch.from_numpy(images).to(dev)
    y = torch.from_numpy(labels.view("uint8")).to(dev)
    lr = 1e-4

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.train_step(x, y, lr, keep_prob=0.5)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = lib.fcn8_launch_count() + eng.graph_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.train_step(x, y, lr, keep_prob=0.5)
    e1.record()
    barrier()
    ms = fdist.max_over_ranks(e0.elapsed_time(e1), dev)
    launches = lib.fcn8_launch_count() + eng.graph_launches - l0
    clk = clocks.stop()
    loss = eng.loss_value(x.shape)
    # per-kernel CUDA-event timing of the tensor-core GEMMs: the same steps, run eagerly right after the timed region
    # (events cannot be recorded inside the replayed CUDA graph the timed region uses)
    timer = ops.KernelTimer()
    ops.TIMER = timer
    k_steps = max(1, min(args.steps, 5))
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(k_steps):
        eng.train_step(x, y, lr, keep_prob=0.5)
    e3.record()
    barrier()
    ops.TIMER = None
    ksum = timer.summary()
    ksum["_steps"] = k_steps
    ksum["_ms"] = e2.elapsed_time(e3)

    if args.profile:
        return dict(ms=ms, launches=launches, clocks=clk, kernels=ksum, loss=loss, e2e_ms=float("nan"), h2d=0, d2h=0,
                    mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    # end to end through the public class surface: FCN8s.train() pulling host numpy batches from a generator
    # (pinned staging + H2D of every batch on a side stream, loss D2H every step), exactly the user's call
    import contextlib
    import io

    def host_batches():
        while True:
            yield images, labels

    def run_train(n):
        with contextlib.redirect_stdout(io.StringIO()):      # tqdm progress goes to stdout like the reference's
            model.train(host_batches(), epochs=1, steps_per_epoch=n, learning_rate_schedule=lambda step: lr,
                        keep_prob=0.5, record_summaries=False)

    run_train(min(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    run_train(args.steps)
    e1.record()
    barrier()
    e2e_ms = fdist.max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)), dev)
    out = dict(ms=ms, launches=launches, clocks=clk, kernels=ksum, loss=loss, e2e_ms=e2e_ms,
               h2d=int(images.nbytes + labels.nbytes), d2h=int(eng.loss_buf.numel() * 4),
               mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    model.engine = None
    del model, eng
    torch.cuda.empty_cache()
    return out


def ncu_traffic(precision):
    """DRAM bytes per step of the dominant kernel family from the committed ncu metrics pass (profiles/): sum of
    dram__bytes_read.sum + dram__bytes_write.sum over the step's conv_gemm_kernel / conv_halo_kernel launches of the
    encoder (bf16 operand instantiations).  None when the profile is missing or was taken in another mode."""
    path = os.path.join(ROOT, "profiles", "r01_step_kernels_final.json")
    if not os.path.exists(path):
        return None, None
    d = json.load(open(path)).get(precision)
    if not d:
        return None, None
    # conv_gemm_kernel<BN, TF32 = 0, PAIR> are the bf16-operand instantiations (the tf32 ones serve the decoder)
    tot = sum(v["dram_bytes"] for k, v in d.items()
              if re.match(r"conv_gemm_kernel<\d+, 0, [01]>", k) or k.startswith("conv_halo_kernel"))
    return tot, "profiles/r01_step_kernels_final.json"


def line_for(precision, r, args, world, peaks):
    n_img = world * PER_GPU_BATCH * args.steps
    value = n_img / (r["ms"] * 1e-3)
    k = r["kernels"].get("conv_gemm", dict(launches=0, flops=0.0, ms=1.0))
    kw = r["kernels"].get("wgrad_gemm", dict(launches=0, flops=0.0, ms=1.0))
    ksteps = r["kernels"].get("_steps", args.steps)
    step_ms = r["ms"] / args.steps
    achieved = k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["launches"] else 0.0
    dtype = {"fp32": "fp32-equivalent (bf16 hi/lo pair storage, error-compensated hi*hi+hi*lo+lo*hi products on the "
                     "bf16 tensor cores, fp32 accumulate; logits within 1e-4 of the fp64 CPU graph)",
             "tf32x3": "fp32 storage, 3xTF32 error-compensated products", "tf32": "tf32", "bf16": "bf16"}[precision]
    return {
        "value": value, "ms_per_step": r["ms"] / args.steps, "dtype": dtype,
        "gpu_launches": int(r["launches"]),
        "roofline": {
            "bound": "tensor", "kernel": "conv_gemm_kernel + conv_halo_kernel (tcgen05 implicit-GEMM fprop + dgrad of the "
                                         "encoder, all tile widths)",
            "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
            "traffic": ncu_traffic(precision)[0], "traffic_unit": "DRAM bytes per step over the same launches "
                                                                  "(ncu dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": ncu_traffic(precision)[1], "peak_source": peaks["source"],
            "launches_timed": k["launches"], "kernel_ms_per_step": k["ms"] / ksteps,
            "share_of_step": k["ms"] / ksteps / step_ms,
            "note": "achieved = algorithmic 2*M*N*K FLOPs of the launches / their CUDA-event time on the launching "
                    "stream, timed live in this run over %d eager steps right after the timed region (the timed "
                    "region replays a CUDA graph of the step, inside which events cannot be recorded)" % ksteps
                    + ("; the fp32-equivalent mode executes 3 bf16 MMAs per algorithmic product, so its own ceiling "
                       "is peak/3" if precision == "fp32" else ""),
            "wgrad_gemm": {"achieved": kw["flops"] / (kw["ms"] * 1e-3) / 1e12 if kw["launches"] else 0.0,
                           "kernel_ms_per_step": kw["ms"] / ksteps, "share_of_step": kw["ms"] / ksteps / step_ms},
            "whole_step_tflops_per_gpu": value / world * TRAIN_GFLOP_PER_IMAGE / 1e3,
        },
        "e2e": {"value": n_img / (r["e2e_ms"] * 1e-3), "unit": "images/s", "h2d_bytes_per_step": r["h2d"],
                "d2h_bytes_per_step": r["d2h"]},
        "clocks": r["clocks"], "final_loss": r["loss"], "peak_mem_gb": r["mem_gb"],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "tf32", "tf32x3"],
                    help="main line; the other of fp32/bf16 is reported under 'alt'")
    ap.add_argument("--no-alt", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="ncu-friendly run: no warm-up floor, no e2e loop, no alt, no CPU baseline (never a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.profile:
        args.no_alt = args.no_cpu_baseline = True
    else:
        args.warmup = max(args.warmup, 3)

    import torch
    from fcn8s_tensorflow_b200 import dist as fdist
    from fcn8s_tensorflow_b200.fcn8s import synthetic_weights
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the "
                         "CPU baseline)")
    rank, local_rank, world = fdist.init("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = measured_peaks()
    weights = synthetic_weights(C, 2)
    main_r = time_engine(args.precision, weights, args, rank, local_rank, world, dev)
    line = {"metric": METRIC, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic"}
    line.update(line_for(args.precision, main_r, args, world, peaks))
    line["config"] = {
        "workload": "BASELINE configs[1]: FCN-8s train step, %d images of 512x1024x3 per GPU, 20 classes, keep_prob "
                    "0.5, Adam lr 1e-4, synthetic uint8 images + bool one-hot labels, He-init encoder" % PER_GPU_BATCH,
        "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * world, "height": H, "width": W, "classes": C,
        "parallelism": "dp%d: one NCCL all-reduce of the 134.5M-element flat gradient buffer per step" % world
                       if world > 1 else "single GPU",
        "l2_flush": "not needed: every step streams > 2 GB of activations and 2 GB of optimizer state, far above "
                    "the 126 MB L2 (inputs larger than L2)",
        "train_gflop_per_image": TRAIN_GFLOP_PER_IMAGE,
    }
    if not args.no_alt and args.precision in ("fp32", "bf16"):
        alt_p = "bf16" if args.precision == "fp32" else "fp32"
        alt_r = time_engine(alt_p, weights, args, rank, local_rank, world, dev)
        line["alt"] = line_for(alt_p, alt_r, args, world, peaks)
        line["alt"]["note"] = "same workload and run, precision mode '%s' (BASELINE configs[2] trains in bf16)" % alt_p
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on the host cores: one untimed step, then whole steps for ~10 s
        fn, img_per_step = cpu_reference_step_fn(weights, (512, 1024))
        fn()
        n_cpu = 0
        t0 = time.perf_counter()
        while n_cpu < 8 and (n_cpu == 0 or time.perf_counter() - t0 < 10.0):
            fn()
            n_cpu += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_cpu * img_per_step / dt, "unit": "images/s", "cores": os.cpu_count(),
                                "kind": "port",
                                "sample": "%d train steps on 1 image of 512x1024 (batch 1) after one warm-up step, "
                                          "torch-CPU fp32 restatement of the reference graph (TF1 unavailable), %.1f s"
                                          % (n_cpu, dt)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
