"""ncu target: BASELINE config-5 style frame (7680x4320, YCrCb + B=16) through the general kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcf_b200 import Codec
H, W = 4320, 7680
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randint(0, 256, (2, H, W, 3), generator=g, device="cuda", dtype=torch.uint8)
enc = Codec(block_size=16, q=32, color="YCrCb", hist=False); dec = Codec(block_size=16, q=32, color="YCrCb", fp64=True)
for _ in range(2):
    idx, st = enc.encode(x, stats=True)
    y = dec.decode(idx, (H, W), original=x, stats=True)
torch.cuda.synchronize(); print("ok")
