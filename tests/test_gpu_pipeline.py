"""vcf_b200/pipeline.py: the chunked ring between host IO and the GPU path --
results identical to one big batch, bounded buffers, and host work of neighbouring chunks overlapping."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vcf_oracle as O


def test_pipeline_equals_one_batch_and_overlaps_host_work():
    import torch
    from vcf_b200 import Codec
    from vcf_b200.pipeline import ChunkPipeline
    H, W, n = 72, 256, 22
    frames = [O.synthetic_frame(H, W, 600 + i, "natural" if i % 3 else "noise") for i in range(n)]
    codec = Codec(block_size=8, q=16)
    ref = codec.encode(np.stack(frames))
    active = {"read": 0, "finish": 0, "max_both": 0}
    lock = threading.Lock()

    def read(i):
        with lock:
            active["read"] += 1
            active["max_both"] = max(active["max_both"], min(active["read"], 1) + min(active["finish"], 1))
        time.sleep(0.02)                       # stands for PNG decoding
        with lock:
            active["read"] -= 1
        return frames[i]

    def finish(i, k, on_device):
        assert not on_device
        with lock:
            active["finish"] += 1
            active["max_both"] = max(active["max_both"], min(active["read"], 1) + min(active["finish"], 1))
        time.sleep(0.02)                       # stands for zlib + file write
        with lock:
            active["finish"] -= 1
        return np.array(k)

    pipe = ChunkPipeline(device=0, depth=3, chunk=4, io_threads=4)
    pipe.run(n, read, lambda x, m: codec.encode(x), finish)      # first use: pinned allocations, streams, thread pool
    t0 = time.perf_counter()
    out = pipe.run(n, read, lambda x, m: codec.encode(x), finish)
    wall = time.perf_counter() - t0
    assert len(out) == n
    for i in range(n):
        assert np.array_equal(out[i], ref[i]), i
    # reads of chunk c+1 ran while chunk c was being finished
    assert active["max_both"] == 2
    serial = n * 0.04 / 4                      # reads + finishes, 4 host threads, strictly one phase after the other
    # (a loose bound: the overlap itself is asserted above, wall clock on a shared box is not a measurement)
    assert wall < serial * 0.9 + 1.0, (wall, serial)
    # device-resident hand-off (GPU entropy stages) and a sequence shorter than one chunk
    pipe2 = ChunkPipeline(device=0, depth=2, chunk=8, io_threads=2, keep_on_device=True)
    out2 = pipe2.run(3, lambda i: frames[i], lambda x, m: codec.encode(x), lambda i, k, dev: (dev, k.cpu().numpy()), first=0)
    assert all(d for d, _ in out2) and all(np.array_equal(k, ref[i]) for i, (_, k) in enumerate(out2))
    with pytest.raises(ValueError):
        pipe.run(2, lambda i: frames[0] if i == 0 else frames[0][:8], lambda x, m: codec.encode(x), finish)
