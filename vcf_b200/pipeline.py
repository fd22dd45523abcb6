"""Chunked, fixed-depth pipeline between host file IO and the GPU path.

The reference's intra-only driver (/root/reference/src/III.py:64-144) handles one frame at a time:
read a PNG, transform, entropy-code, write -- strictly in sequence.  Here a sequence is cut into
chunks of a few frames that travel through a ring of ``depth`` slots; each slot owns pinned host
buffers, device buffers and a CUDA stream, so that at any moment

    chunk i+1   is being read / entropy-decoded by the host threads into its pinned buffer,
    chunk i     is being copied in, transformed and copied out on its own CUDA stream,
    chunk i-1   is being entropy-coded / written by the host threads from its pinned buffer.

The GPU never waits for a file and a host thread never waits for the GPU longer than one chunk.
Memory is bounded by ``depth * chunk`` frames however long the sequence is.  The transform itself
is whatever ``gpu_fn`` launches on the slot's stream (``Codec.encode`` / ``Codec.decode`` on the
slot's device tensors); nothing in here knows about DCTs.
"""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor, wait
from typing import Callable, Optional, Sequence

import numpy as np


class _Slot:
    def __init__(self):
        self.stream = None
        self.event = None
        self.pin_in = None
        self.pin_out = None
        self.dev_in = None
        self.dev_out = None
        self.pending = []            # futures still using this slot's buffers


class ChunkPipeline:
    """``run`` drives ``n_items`` items (frames) through read -> GPU -> finish.

    read_fn(i) -> numpy array                      runs on a host thread; every item must have the same shape/dtype
    gpu_fn(dev_in, n) -> device tensor             called on the slot's stream with the first ``n`` items of the chunk;
                                                   returns the chunk's result (first dimension >= n)
    finish_fn(i, result_i, on_device) -> any       runs on a host thread once the chunk's result is available:
                                                   ``result_i`` is a numpy view of the pinned output (on_device False) or
                                                   the device tensor of item i (on_device True, for GPU entropy stages)
    keep_on_device: hand device tensors to finish_fn instead of copying the result to the host.
    """

    def __init__(self, device: int = 0, depth: int = 3, chunk: int = 8, io_threads: int = 8, keep_on_device: bool = False):
        if depth < 2 or chunk < 1:
            raise ValueError("depth must be >= 2 and chunk >= 1")
        self.device, self.depth, self.chunk = int(device), int(depth), int(chunk)
        self.io_threads = max(1, int(io_threads))
        self.keep_on_device = keep_on_device
        self.timeline = []           # (chunk, phase, t_start, t_end) for tests / tuning

    def run(self, n_items: int, read_fn: Callable, gpu_fn: Callable, finish_fn: Callable, first: int = 0) -> list:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ChunkPipeline needs a CUDA device: vcf_b200 has no CPU fallback")
        if n_items <= 0:
            return []
        dev = torch.device("cuda", self.device)
        nchunks = (n_items + self.chunk - 1) // self.chunk
        slots = [_Slot() for _ in range(min(self.depth, nchunks))]
        results: list = [None] * n_items
        t_origin = time.perf_counter()

        def now():
            return time.perf_counter() - t_origin

        with torch.cuda.device(dev), ThreadPoolExecutor(self.io_threads) as pool:
            for sl in slots:
                sl.stream = torch.cuda.Stream()
                sl.event = torch.cuda.Event()

            def bounds(c):
                lo = c * self.chunk
                return lo, min(n_items, lo + self.chunk)

            def read_into(sl, j, i):
                arr = proto if (i == 0 and proto is not None) else np.ascontiguousarray(read_fn(first + i))
                if sl.pin_in is None or tuple(sl.pin_in.shape[1:]) != arr.shape or sl.pin_in.numpy().dtype != arr.dtype:
                    raise ValueError("all items of a sequence must have the same shape and dtype")
                np.copyto(sl.pin_in.numpy()[j], arr)

            def alloc_in(sl, arr):
                t = torch.from_numpy(np.empty(0, arr.dtype)).dtype
                sl.pin_in = torch.empty((self.chunk,) + arr.shape, dtype=t, pin_memory=True)
                sl.dev_in = torch.empty((self.chunk,) + arr.shape, dtype=t, device=dev)

            def finish(sl, c, j, i):
                sl.event.synchronize()
                if self.keep_on_device:
                    return finish_fn(first + i, sl.dev_out[j], True)
                return finish_fn(first + i, sl.pin_out.numpy()[j], False)

            reads = {}
            proto = np.ascontiguousarray(read_fn(first))             # the first item fixes shape and dtype
            for c in range(nchunks + 1):
                if c < nchunks:
                    sl = slots[c % len(slots)]
                    if sl.pending:                         # the chunk that last used this slot has been written
                        wait(sl.pending)
                        for f in sl.pending:
                            f.result()
                        sl.pending = []
                    lo, hi = bounds(c)
                    if sl.pin_in is None:
                        alloc_in(sl, proto)
                    t0 = now()
                    reads[c] = (t0, [pool.submit(read_into, sl, i - lo, i) for i in range(lo, hi)])
                cc = c - 1
                if cc >= 0:
                    sl = slots[cc % len(slots)]
                    lo, hi = bounds(cc)
                    t0, futs = reads.pop(cc)
                    for f in futs:
                        f.result()
                    self.timeline.append((cc, "read", t0, now()))
                    tg = now()
                    with torch.cuda.stream(sl.stream):
                        sl.dev_in[: hi - lo].copy_(sl.pin_in[: hi - lo], non_blocking=True)
                        out = gpu_fn(sl.dev_in[: hi - lo], hi - lo)
                        sl.dev_out = out
                        if not self.keep_on_device:
                            if sl.pin_out is None or tuple(sl.pin_out.shape[1:]) != tuple(out.shape[1:]) or sl.pin_out.dtype != out.dtype:
                                sl.pin_out = torch.empty((self.chunk,) + tuple(out.shape[1:]), dtype=out.dtype, pin_memory=True)
                            sl.pin_out[: hi - lo].copy_(out[: hi - lo], non_blocking=True)
                        sl.event.record(sl.stream)
                    self.timeline.append((cc, "gpu_issue", tg, now()))
                    tw = now()
                    futs = [pool.submit(finish, sl, cc, i - lo, i) for i in range(lo, hi)]
                    sl.pending = futs
                    for i, f in zip(range(lo, hi), futs):
                        results[i] = f
                    self.timeline.append((cc, "finish_submit", tw, now()))
            for i in range(n_items):
                results[i] = results[i].result()
            torch.cuda.synchronize(dev)
        return results
