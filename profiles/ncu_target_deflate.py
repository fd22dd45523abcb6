"""ncu target for the deflate kernels (row F4): usage: python profiles/ncu_target_deflate.py [frames] [q] [rows|runs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vcf_b200 import Codec
from vcf_b200.entropy import deflate_raw_dev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
q = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mode = sys.argv[3] if len(sys.argv) > 3 else "rows"
x = bench.make_frames(torch, n, 2160, 3840, torch.device("cuda", 0), 99, "natural")
idx = Codec(block_size=8, q=q).encode(x)
geom = (idx.shape[2] * 3, 3) if mode == "rows" else (0, 1)
for _ in range(3):
    dst, nb = deflate_raw_dev(idx.reshape(-1), geom)
torch.cuda.synchronize()
print("ok", int(nb.item()))
