import sys; sys.path.insert(0, "/root/repo")
import torch, numpy as np
from oracle import vcf_oracle as O
from vcf_b200.rd import rd_sweep, rd_point
img = O.synthetic_frame(1080, 1920, 3, "natural"); x = torch.from_numpy(img).cuda()
for p in rd_sweep(x, (8,), (4, 8, 16, 32, 64)): print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.items()})
for q in (4, 64):
    k = O.encode_array(img, 8, q); y = O.decode_array(k, img.shape, 8, q)
    print(q, "oracle rmse", float(O.rmse(img, y)), "sse", O.sse_int(img, y))
