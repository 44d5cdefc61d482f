"""Data-parallel parity (run under torchrun, one rank per GPU): the averaged shard gradients after the engine's
two-chunk all-reduce equal the single-GPU gradients of the concatenated batch, and the replicas stay bit-identical
after Adam steps.  Used by tests/test_gpu_dp.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402

from fcn8s_tensorflow_b200 import dist as fdist  # noqa: E402
from fcn8s_tensorflow_b200.engine import Engine  # noqa: E402
from fcn8s_tensorflow_b200.fcn8s import synthetic_weights  # noqa: E402


def main():
    rank, local_rank, world = fdist.init("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    C, per, H, W = 5, 2, 64, 96
    N = per * world
    rng = np.random.default_rng(0)
    images = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
    labels = np.eye(C, dtype=np.uint8)[rng.integers(0, C, size=(N, H, W))]
    weights = synthetic_weights(C, 2, decoder_std_scale=10.0)
    ok = True
    # (fp32: 6e-8 without ReLU / max-pool flips between the sharded and the single-GPU run, ~1e-3 with a handful)
    for precision, tol in (("fp32", 5e-3), ("bf16", 2e-2)):
        for overlap in (True, False):
            e = Engine(C, precision=precision, device=dev)
            e.load_weights(weights)
            fdist.attach(e)
            e.allreduce.overlap = overlap
            fdist.broadcast_parameters(e)
            xi, yi = fdist.shard_batch(images, labels, rank, world)
            x = torch.from_numpy(np.ascontiguousarray(xi)).to(dev)
            y = torch.from_numpy(np.ascontiguousarray(yi)).to(dev)
            e.loss_and_backward(x, y, keep_prob=1.0)
            e._start_reduce(e._reduced_upto, e.n_flat)     # the tail chunk, in the engine's wire format
            e.allreduce.finish()
            e._reduced_upto = 0
            torch.cuda.synchronize()
            # bf16 wire format: the summed gradient lives in the bf16 wire buffer (that is what Adam reads)
            avg = ((e.g16.float() if e.g16 is not None else e.grads) / world).clone()
            ref = Engine(C, precision=precision, device=dev)
            ref.load_weights(weights)
            ref.loss_and_backward(torch.from_numpy(images).to(dev), torch.from_numpy(labels).to(dev), keep_prob=1.0)
            torch.cuda.synchronize()
            err = float((avg - ref.grads).norm() / ref.grads.norm())
            worst = 0.0
            for name in e.layout:
                a, b = e.view(name, avg), ref.view(name, ref.grads)
                worst = max(worst, float((a - b).norm() / (b.norm() + 1e-30)))
            print("rank %d %s overlap=%d: flat grad rel-l2 %.3e, worst tensor %.3e (tol %.0e)" %
                  (rank, precision, overlap, err, worst, tol))
            ok &= err <= tol and worst <= 10 * tol
            # two DP train steps: replicas must stay bit-identical
            for _ in range(2):
                e.train_step(x, y, 1e-4, keep_prob=0.5)
            torch.cuda.synchronize()
            mine = e.params.clone()
            other = mine.clone()
            tdist.broadcast(other, src=0)
            same = bool(torch.equal(mine, other))
            print("rank %d %s overlap=%d: replicas identical after 2 steps: %s" % (rank, precision, overlap, same))
            ok &= same
            del e, ref
            torch.cuda.empty_cache()
    flag = torch.tensor([1 if ok else 0], device=dev)
    tdist.all_reduce(flag, op=tdist.ReduceOp.MIN)
    tdist.barrier()
    tdist.destroy_process_group()
    if rank == 0:
        print("DP_CHECK_OK" if int(flag.item()) else "DP_CHECK_FAILED")
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
