"""Encode / decode time per block size on 8 resident 3840x2160 frames (q = 32): which kernel serves each request
and how fast.  usage: python profiles/b_sweep_bench.py [natural|noise]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import vcf_oracle as O
from vcf_b200 import Codec, _lib


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
n, H, W = 8, 2160, 3840
x = torch.from_numpy(np.stack([O.synthetic_frame(H, W, 50 + i, kind) for i in range(n)])).cuda()
spin = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(200):
    spin.add_(1)
out = {"content": kind, "frames": n}
for B in (4, 8, 16, 32):
    enc, d64, d32 = Codec(block_size=B, q=32), Codec(block_size=B, q=32, fp64=True), Codec(block_size=B, q=32)
    idx = enc.encode(x)
    r = {"encode_ms": round(timed(lambda: enc.encode(x, out=idx)), 4), "encode_kernel": _lib.last_kernel()}
    y = d64.decode(idx, (H, W))
    r["decode_f64_ms"] = round(timed(lambda: d64.decode(idx, (H, W), out=y)), 4)
    r["decode_f64_kernel"] = _lib.last_kernel()
    r["decode_f32_ms"] = round(timed(lambda: d32.decode(idx, (H, W), out=y)), 4)
    r["decode_f32_kernel"] = _lib.last_kernel()
    px = n * H * W / 1e6
    r["gpixel_s"] = {k[:-3]: round(px / r[k], 1) for k in ("encode_ms", "decode_f64_ms", "decode_f32_ms")}
    out[f"B{B}"] = r
print(json.dumps(out))
