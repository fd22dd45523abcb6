import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcf_b200 import Codec
H, W = 2160, 3840
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randint(0, 256, (4, H, W, 3), generator=g, device="cuda", dtype=torch.uint8)
enc = Codec(block_size=32, q=32, hist=False); dec = Codec(block_size=32, q=32, fp64=True)
for _ in range(2):
    idx, st = enc.encode(x, stats=True)
    y = dec.decode(idx, (H, W), original=x, stats=True)
torch.cuda.synchronize(); print("ok")
