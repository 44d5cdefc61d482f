"""Where does the data-parallel step spend the all-reduce's exposed time?  (run under torchrun, eager steps)
CUDA events on the compute stream: step start -> head chunk started -> backward done -> collectives joined -> Adam done,
with the all-reduce on and off, for one config of bench.py.
    torchrun --nproc-per-node N scripts/dp_timeline.py [c3|c2] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FCN8_GRAPHS"] = "0"
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402

import bench  # noqa: E402
from fcn8s_tensorflow_b200 import dist as fdist  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402
from fcn8s_tensorflow_b200.fcn8s import synthetic_weights  # noqa: E402


def main():
    cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
    precision = sys.argv[2] if len(sys.argv) > 2 else cfg["precision"]
    rank, local_rank, world = fdist.init("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    e = Engine(cfg["C"], precision=precision, device=dev)
    e.use_graphs = False
    e.load_weights(synthetic_weights(cfg["C"], 2))
    fdist.attach(e)
    fdist.broadcast_parameters(e)
    images, labels = bench.synthetic_feed(cfg, cfg["per_gpu"], 1000 + rank)
    x = torch.from_numpy(images).to(dev)
    y = torch.from_numpy(labels.view("uint8")).to(dev)
    marks = {}

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(dev))
        marks.setdefault(name, []).append(ev)

    orig_start, orig_radam = e._start_reduce, e._reduce_and_adam
    state = {"first": True}

    def start(lo, hi):
        if state["first"]:
            mark("head_started")
            state["first"] = False
        orig_start(lo, hi)

    def radam(lr_t):
        mark("backward_done")
        if e.allreduce is not None:
            e._start_reduce(e._reduced_upto, e.n_flat)
            e.allreduce.finish()
            e._reduced_upto = 0
        mark("joined")
        saved, e.allreduce = e.allreduce, None
        try:
            orig_radam(lr_t)
        finally:
            e.allreduce = saved
        mark("adam_done")

    e._start_reduce, e._reduce_and_adam = start, radam
    saved_ar = e.allreduce
    for label, ar in (("all-reduce ON ", saved_ar), ("all-reduce OFF", None)):
        e.allreduce = ar
        for i in range(13):
            if i == 3:
                marks.clear()
            state["first"] = True
            mark("start")
            if ar is None:
                pass
            e.train_step(x, y, 1e-4, keep_prob=cfg["keep_prob"])
        torch.cuda.synchronize()
        tdist.barrier()
        n = len(marks["start"])

        def seg(a, b):
            if a not in marks or b not in marks:
                return float("nan")
            return sum(p.elapsed_time(q) for p, q in zip(marks[a], marks[b])) / n
        if rank == 0:
            print("%s  fwd+bwd %.3f ms (to head start %.3f, head start -> backward done %.3f) | join %.3f | adam+repack %.3f | "
                  "step %.3f" % (label, seg("start", "backward_done"), seg("start", "head_started"),
                                 seg("head_started", "backward_done"), seg("backward_done", "joined"),
                                 seg("joined", "adam_done"), seg("start", "adam_done")))
        marks.clear()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
