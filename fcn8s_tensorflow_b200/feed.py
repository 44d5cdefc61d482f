"""Prefetching feed for the `FCN8s` train / evaluate loops.

The reference pulls `next(generator)` synchronously inside the step loop and lets `sess.run` copy the numpy feeds to
the device every step (fcn8s_tensorflow.py:551-572, 683-689), so the GPU idles during PNG decoding and the H2D copy.
Here a background thread pulls exactly `count` batches from the generator (never more, so a generator shared between
training and evaluation is consumed in the reference's order), stages each in pinned host memory and issues the H2D
copy on a side CUDA stream; the step loop only makes the compute stream wait on the copy's event.  Slots are recycled
once the step that consumed them has been enqueued (event on the compute stream).

Labels: the generators yield bool one-hot [n,H,W,C] (helpers/ground_truth_conversion_utils.py:84-88), which the
reference's feed widens to int32 on the host (4C bytes per pixel over PCIe).  For C >= 8 the feed thread packs the batch
to one class id per pixel (`fcn8_pack_labels`, host code that also verifies that every pixel IS one-hot), ships C times
fewer bytes and restores the one-hot tensor on the device (`fcn8_expand_labels`, on the copy stream); a batch with a
soft / multi-hot / empty label row is shipped unchanged.  FCN8_FEED_PACK_LABELS=0 switches the packing off.
"""
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops

_COPY_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="fcn8-feed-copy")


def _parallel_copy(dst, src):
    """numpy memcpy of a batch into pinned memory, split over the batch axis across a few threads (numpy releases the
    GIL; under torchrun OMP_NUM_THREADS=1 makes a single torch copy_ of the 42 MB one-hot label batch the bottleneck of
    the whole feed at 8 ranks per host)."""
    n = src.shape[0]
    if n < 2 or src.nbytes < (4 << 20):
        np.copyto(dst, src)
        return
    parts = min(4, n)
    bounds = [n * i // parts for i in range(parts + 1)]
    futs = [_COPY_POOL.submit(np.copyto, dst[bounds[i]:bounds[i + 1]], src[bounds[i]:bounds[i + 1]])
            for i in range(parts)]
    for f in futs:
        f.result()


class _Slot:
    def __init__(self):
        self.pin = {}
        self.dev = {}
        self.copied = torch.cuda.Event()
        self.consumed = None  # event recorded on the compute stream after the consuming step was enqueued


class Feeder:
    def __init__(self, model, generator, count, depth=3):
        self.model = model
        self.generator = generator
        self.count = int(count)
        self.device = model.engine.device
        # pinned staging slots and the copy stream live on the model: allocating page-locked memory costs
        # milliseconds, and a Feeder is created per epoch / per evaluation
        cache = model.__dict__.setdefault("_feed_cache", {})
        if "stream" not in cache:
            cache["stream"] = torch.cuda.Stream(device=self.device)
            cache["slots"] = [_Slot() for _ in range(depth)]
        self.copy_stream = cache["stream"]
        self.pack_labels = os.environ.get("FCN8_FEED_PACK_LABELS", "1") != "0"
        self.h2d_bytes = 0          # bytes of the last staged batch that crossed PCIe
        prev = cache.get("feeder")
        if prev is not None and prev.thread.is_alive():
            # an earlier Feeder's thread is still staging (its loop raised, or close() timed out): it may hold the
            # cached slots, so this one gets fresh ones instead of sharing pinned / device buffers with it
            prev.stop.set()
            cache["slots"] = [_Slot() for _ in range(depth)]
        cache["feeder"] = self
        self.stop = threading.Event()
        self.free = queue.Queue()
        for slot in cache["slots"]:
            self.free.put(slot)
        self.ready = queue.Queue()
        self.current = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _device_buffer(self, slot, key, shape, dtype):
        t = slot.dev.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            if t is not None:       # last used on the compute stream: the allocator must not recycle it before that
                t.record_stream(torch.cuda.current_stream(self.device))
                if slot.consumed is not None:
                    slot.consumed.synchronize()
            t = slot.dev[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def _pinned_buffer(self, slot, key, shape, dtype):
        t = slot.pin.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = slot.pin[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return t

    def _stage(self, slot, key, a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        dtype = torch.from_numpy(a[:0]).dtype
        pin = self._pinned_buffer(slot, key, a.shape, dtype)
        dev = self._device_buffer(slot, key, a.shape, dtype)
        _parallel_copy(pin.numpy(), a)
        dev.copy_(pin, non_blocking=True)
        self.h2d_bytes += a.nbytes
        return dev

    def _stage_labels(self, slot, a):
        """One-hot labels [n,H,W,C]: class ids over PCIe when the batch is exactly one-hot (module docstring)."""
        a = np.ascontiguousarray(a)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        if not (self.pack_labels and a.dtype == np.uint8 and a.ndim == 4 and 8 <= a.shape[-1] <= 255):
            return self._stage(slot, "labels", a)
        ids_pin = self._pinned_buffer(slot, "label_ids", a.shape[:-1], torch.uint8)
        if not ops.pack_labels(a, ids_pin.numpy(), threads=4):
            return self._stage(slot, "labels", a)
        ids_dev = self._device_buffer(slot, "label_ids", a.shape[:-1], torch.uint8)
        dev = self._device_buffer(slot, "labels", a.shape, torch.uint8)
        ids_dev.copy_(ids_pin, non_blocking=True)
        ops.expand_labels(ids_dev, dev)      # on the copy stream, like the copies
        self.h2d_bytes += ids_pin.numel()
        return dev

    def _run(self):
        try:
            torch.cuda.set_device(self.device)
            for _ in range(self.count):
                if self.stop.is_set():
                    return
                images, labels = next(self.generator)
                images = self.model._check_images(images)
                labels = self.model._check_labels(labels)
                slot = self.free.get()
                if self.stop.is_set():
                    return
                slot.copied.synchronize()            # the pinned buffers are free once their last H2D has finished
                with torch.cuda.stream(self.copy_stream):
                    if slot.consumed is not None:     # the device buffers are free once their last step was enqueued
                        self.copy_stream.wait_event(slot.consumed)
                    self.h2d_bytes = 0
                    x = self._stage(slot, "images", images)
                    y = self._stage_labels(slot, labels)
                    self.model.__dict__["feed_h2d_bytes_per_batch"] = self.h2d_bytes
                    slot.copied.record(self.copy_stream)
                self.ready.put((slot, x, y))
        except BaseException as e:  # noqa: BLE001 -- re-raised in the consumer
            self.ready.put(e)

    def get(self):
        """Next (images, labels) device pair; the current stream waits for its copy."""
        item = self.ready.get()
        if isinstance(item, BaseException):
            raise item
        slot, x, y = item
        torch.cuda.current_stream(self.device).wait_event(slot.copied)
        self.current = slot
        return x, y

    def release(self):
        """Call after the step that uses the current batch has been enqueued."""
        slot = self.current
        if slot.consumed is None:
            slot.consumed = torch.cuda.Event()
        slot.consumed.record(torch.cuda.current_stream(self.device))
        self.current = None
        self.free.put(slot)

    def close(self):
        """Stop the prefetch thread (it exits before pulling another batch) and wait for it."""
        self.stop.set()
        self.free.put(_Slot())          # wakes a thread blocked on an empty free queue
        self.thread.join(timeout=30)
