"""Fixed workload for ncu captures of the fused rate/distortion sweep: BASELINE configs[2], one 4K frame,
8 steps, the four block sizes once each.  usage: python profiles/ncu_target_rd.py [natural|noise]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vcf_oracle as O
from vcf_b200.rd import rd_stats_fused

kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
x = torch.from_numpy(O.synthetic_frame(2160, 3840, 2, kind)).cuda()
for B in (4, 8, 16, 32):
    s = rd_stats_fused(x, B, (4, 8, 12, 16, 24, 32, 48, 64))
torch.cuda.synchronize()
print("ok", int(s[0, 0].item()))
