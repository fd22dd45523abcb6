"""Fixed workload for ncu captures of the tensor-core tier: fast-mode encode + decode of 4K frames at the
bench's batch size.  usage: python profiles/ncu_target_tc.py [frames] [q] [content noise|natural]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcf_b200 import Codec, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
q = int(sys.argv[2]) if len(sys.argv) > 2 else 8
content = sys.argv[3] if len(sys.argv) > 3 else "noise"
H, W = 2160, 3840
if content == "noise":
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randint(0, 256, (n, H, W, 3), generator=g, device="cuda", dtype=torch.uint8)
else:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    x = bench.make_frames(torch, n, H, W, torch.device("cuda"), 1234, "natural")
enc = Codec(block_size=8, q=q, fast=True)
dec = Codec(block_size=8, q=q)
idx = torch.empty_like(x)
y = torch.empty_like(x)
for _ in range(3):
    enc.encode(x, out=idx)
    ke = _lib.last_kernel()
    dec.decode(idx, (H, W), out=y)
    kd = _lib.last_kernel()
torch.cuda.synchronize()
print("ok", ke, kd, int(y[0, 0, 0, 0].item()))
