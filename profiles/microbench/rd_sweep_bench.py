"""BASELINE config 3: the RD sweep B in {4,8,16,32} x 8 q values on 4K frames kept on the device.
   python profiles/microbench/rd_sweep_bench.py [frames]
Prints ms per (B, q) point (encode + decode with statistics, CUDA events) and the total."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
from vcf_b200 import _lib
from vcf_b200.rd import rd_point

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = bench.make_frames(torch, n, torch.device("cuda", 0), 77)
QS = (4, 8, 12, 16, 24, 32, 48, 64)
res = {}
for B in (4, 8, 16, 32):
    rd_point(x, B, 32)                      # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pts = [rd_point(x, B, q) for q in QS]
    e1.record()
    torch.cuda.synchronize()
    res[B] = {"ms_per_point": e0.elapsed_time(e1) / len(QS), "last_kernel": _lib.last_kernel(),
              "psnr_q32": [p["psnr"] for p in pts if p["q"] == 32][0]}
tot = sum(v["ms_per_point"] * len(QS) for v in res.values())
print(json.dumps({"workload": f"{n} frames 3840x2160, 32 RD points (encode + float64 decode + statistics incl. histogram)",
                  "per_block_size": res, "total_ms": tot, "gpixel_s": 32 * n * 3840 * 2160 / 1e9 / (tot / 1e3)}))
