"""Data parallelism for the FCN-8s step: one process per GPU, images sharded across ranks, ONE all-reduce of the flat
gradient buffer per step (NCCL over NVLink 5 / NVSwitch), 1/world folded into the fused Adam kernel.

The reference has no distributed code at all; the loss is a mean over N*H*W
(fcn8s_tensorflow.py:253), so equal per-rank batches + gradient averaging reproduce the single-device batch exactly up
to summation order.  The host logic is backend-agnostic (`gloo` on CPU tensors in tests/test_dist_cpu.py).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend="nccl"):
    """Initialise the default process group from the environment (MASTER_ADDR / MASTER_PORT set by torchrun)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_bounds(n, rank, world):
    """Rank r takes items [r*n/world, (r+1)*n/world) of each generator batch; n must divide evenly so that the
    per-rank means average to the global mean."""
    if n % world:
        raise ValueError("batch of %d images does not split evenly across %d ranks" % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def shard_batch(images, labels, rank, world):
    lo, hi = shard_bounds(len(images), rank, world)
    return images[lo:hi], (None if labels is None else labels[lo:hi])


class GradientAllReduce:
    """Callable installed as `Engine.allreduce`: sums the flat gradient buffer over all ranks in one collective.
    The 1/world average is applied by the consumer (Adam's grad_scale), not here."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self, flat_grad):
        if self.world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return flat_grad


def attach(engine, group=None):
    """Make `engine` data-parallel over the default (or given) process group."""
    ar = GradientAllReduce(group)
    engine.world = ar.world
    engine.allreduce = ar if ar.world > 1 else None
    return engine


def broadcast_parameters(engine, src=0, group=None):
    """Replicas must start identical (the reference has a single copy): broadcast params + Adam state from `src`."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in (engine.params, engine.adam_m, engine.adam_v):
            dist.broadcast(t, src=src, group=group)
        engine._packed_dirty = True


def max_over_ranks(value, device):
    """Max of a python float over ranks (timings are reported as the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
