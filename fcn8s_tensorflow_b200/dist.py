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
    """Installed as `Engine.allreduce`: sums the flat gradient buffer over all ranks.  The 1/world average is applied
    by the consumer (Adam's grad_scale), not here.

    The buffer is laid out in backward-completion order (engine.flat_layout), so the engine calls `start(head)` on the
    [decoder | fc7 | fc6] prefix (89 % of the bytes) as soon as the fc6 filter gradient has been enqueued -- that
    collective then runs on NCCL's stream underneath the conv5 -> conv1 backward -- and `start(tail)` + `finish()` just
    before Adam.  Logically still one all-reduce of one buffer per step, issued as two chunks.  `overlap=False` (or
    FCN8_DP_OVERLAP=0) issues it as a single blocking collective after the backward pass."""

    def __init__(self, group=None, overlap=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = (os.environ.get("FCN8_DP_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.pending = []

    def start(self, chunk):
        """Enqueue the all-reduce of a contiguous slice of the flat gradient (async; ordered after the work already
        enqueued on the current stream)."""
        if self.world > 1 and chunk.numel():
            self.pending.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Make the current stream wait for every started chunk."""
        for w in self.pending:
            w.wait()
        self.pending = []

    def __call__(self, flat_grad):
        self.start(flat_grad)
        self.finish()
        return flat_grad


def _registered_zeros(n, dtype, device, group):
    """`n` zeros allocated with ncclMemAlloc and registered with the communicator (torch's NCCL memory pool), so that
    the all-reduce can use NVLink-SHARP (NVLS) zero-copy user buffers: the reduction then happens in the NVSwitch and
    the collective needs a handful of CTAs instead of competing with the backward GEMMs for SMs.  Returns
    (tensor, pool) -- (None, None) when the backend cannot do it (gloo, old torch) or FCN8_NCCL_REGISTER=0."""
    if os.environ.get("FCN8_NCCL_REGISTER", "1") == "0" or not str(device).startswith("cuda"):
        return None, None
    try:
        pg = group if group is not None else dist.group.WORLD
        backend = pg._get_backend(torch.device(device))
        pool = torch.cuda.MemPool(backend.mem_allocator)
        with torch.cuda.use_mem_pool(pool, device=torch.device(device)):
            t = torch.zeros(n, dtype=dtype, device=device)
        backend.register_mem_pool(pool)
        return t, pool
    except Exception as e:  # noqa: BLE001 -- registration is an optimisation, never a requirement
        if os.environ.get("FCN8_DEBUG_DP"):
            print("fcn8 dist: NCCL buffer registration unavailable (%s: %s)" % (type(e).__name__, e))
        return None, None


def attach(engine, group=None):
    """Make `engine` data-parallel over the default (or given) process group."""
    ar = GradientAllReduce(group)
    engine.world = ar.world
    engine.rank = dist.get_rank(group) if dist.is_initialized() else 0
    engine.allreduce = ar if ar.world > 1 else None
    engine.nccl_registered = False
    if engine.allreduce is None:
        return engine
    lib = getattr(engine, "lib", None)
    if lib is not None and os.environ.get("FCN8_DYNAMIC_TILES", "1") != "0":
        # the collective's CTAs take SMs away from the persistent GEMM kernels of the backward pass: let the resident
        # CTAs take over the tiles of those that cannot start (cluster launch control, ConvGemmArgs::dyn)
        lib.fcn8_debug_set(10, 1)
    if getattr(engine, "grad_comm", "fp32") == "bf16":
        # wire copy of the flat gradient: the collective moves 269 MB instead of 538 MB per step
        t, pool = _registered_zeros(engine.n_flat, torch.bfloat16, engine.device, group)
        engine.g16 = t if t is not None else torch.zeros(engine.n_flat, dtype=torch.bfloat16, device=engine.device)
    else:
        # fp32 wire format: the flat gradient buffer itself is what NCCL reduces in place
        g = getattr(engine, "grads", None)
        t, pool = _registered_zeros(g.numel(), g.dtype, g.device, group) if g is not None else (None, None)
        if t is not None:
            engine.grads = t
    engine._nccl_pool = pool          # keeps the registered segment alive for the life of the engine
    engine.nccl_registered = pool is not None
    return engine


def broadcast_parameters(engine, src=0, group=None):
    """Replicas must start identical (the reference has a single copy): broadcast params + Adam state from `src`."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in (engine.params, engine.adam_m, engine.adam_v):
            dist.broadcast(t, src=src, group=group)
        step = torch.tensor([engine.global_step], dtype=torch.int64, device=engine.params.device)
        dist.broadcast(step, src=src, group=group)
        engine.global_step = int(step.item())
        engine._packed_dirty = True
        engine._shadow_dirty = True


def all_reduce_metrics(loss_pair, conf, group=None):
    """Evaluation under data parallelism: sum the per-rank (loss sum, batch count) pair and the confusion matrix so
    that every rank reports the metrics of the whole evaluation set (and takes the same save-best decisions)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(loss_pair, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(conf, op=dist.ReduceOp.SUM, group=group)
    return loss_pair, conf


def max_over_ranks(value, device):
    """Max of a python float over ranks (timings are reported as the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
