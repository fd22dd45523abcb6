"""Small fixed workload for ncu captures: encode + decode of a few 4K frames.
usage: python profiles/ncu_target.py [frames] [decode_mode fp32|fp64] [contract 0|1] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcf_b200 import Codec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "fp64"
contract = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
B = int(sys.argv[4]) if len(sys.argv) > 4 else 8
H, W = 2160, 3840
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randint(0, 256, (n, H, W, 3), generator=g, device="cuda", dtype=torch.uint8)
enc = Codec(block_size=B, q=32, contract=contract)
dec = Codec(block_size=B, q=32, fp64=(mode == "fp64"))
for _ in range(3):
    idx = enc.encode(x)
    y = dec.decode(idx, (H, W))
torch.cuda.synchronize()
print("ok", int(y.sum().item()) & 0xffff)
