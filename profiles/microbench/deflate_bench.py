"""Row F4: throughput and compressed size of vcfb_deflate_dev on index planes of the transform path,
next to zlib (level 6, one host core) on the same bytes.

    python profiles/microbench/deflate_bench.py [frames] > gpurun_out/deflate_bench.json
"""
import json
import os
import sys
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vcf_oracle as O          # synthetic frames only (input generator)  # noqa: E402
from vcf_b200 import Codec, _lib            # noqa: E402

H, W = 2160, 3840


def time_call(kb, reps=10):
    L = _lib.lib()
    n = kb.numel()
    dst = torch.empty(L.vcfb_deflate_bound(n), dtype=torch.uint8, device="cuda")
    ws = torch.empty(L.vcfb_deflate_workspace(n), dtype=torch.uint8, device="cuda")
    nb = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def go():
        _lib.check(L.vcfb_deflate_dev(kb.data_ptr(), n, dst.data_ptr(), dst.numel(), nb.data_ptr(), ws.data_ptr(), ws.numel(), st))
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, int(nb.item()), dst


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    imgs = np.stack([O.synthetic_frame(H, W, 2 + i, "natural") for i in range(min(frames, 4))])
    x = torch.from_numpy(imgs).cuda()
    x = x.repeat((frames + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:frames].contiguous()
    rows = []
    for q in (4, 8, 16, 32, 64):
        k = Codec(8, q).encode(x)
        for nf in (1, frames):
            kb = k[:nf].reshape(-1)
            ms, nbytes, dst = time_call(kb)
            row = {"q": q, "frames": nf, "input_MB": kb.numel() / 1e6, "ms": ms, "input_GB_s": kb.numel() / 1e6 / ms,
                   "mpixel_s": nf * H * W / 1e3 / ms, "bits_per_pixel": 8.0 * nbytes / (nf * H * W)}
            if nf == 1:
                host = kb.cpu().numpy().tobytes()
                assert zlib.decompress(dst[:nbytes].cpu().numpy().tobytes(), -15) == host
                t0 = time.perf_counter()
                z = zlib.compress(host, 6)
                dt = time.perf_counter() - t0
                c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_RLE)
                rle = len(c.compress(host) + c.flush())
                row.update({"zlib6_bits_per_pixel": 8.0 * len(z) / (H * W), "zlib_rle_bits_per_pixel": 8.0 * rle / (H * W),
                            "zlib6_mpixel_s_one_core": H * W / 1e6 / dt, "size_vs_zlib6": nbytes / len(z)})
            rows.append(row)
            print(json.dumps(row), flush=True)
    noise = torch.randint(0, 256, (H * W * 3,), dtype=torch.uint8, device="cuda")
    ms, nbytes, _ = time_call(noise)
    print(json.dumps({"input": "uniform noise (stored blocks)", "ms": ms, "input_GB_s": noise.numel() / 1e6 / ms, "out_over_in": nbytes / noise.numel()}))


if __name__ == "__main__":
    main()
