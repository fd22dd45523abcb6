"""ncu target for the float64 decoders: usage: python profiles/ncu_target_dec.py [frames] [noise|natural] [q]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vcf_b200 import Codec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else "noise"
q = int(sys.argv[3]) if len(sys.argv) > 3 else 8
H, W = 2160, 3840
if kind == "noise":
    x = torch.randint(0, 256, (n, H, W, 3), device="cuda", dtype=torch.uint8)
else:
    x = bench.make_frames(torch, n, torch.device("cuda", 0), 1234)
idx = Codec(block_size=8, q=q).encode(x)
dec = Codec(block_size=8, q=q, fp64=True)
for _ in range(3):
    y = dec.decode(idx, (H, W))
torch.cuda.synchronize()
print("ok", int(y.sum().item()) & 0xffff)
