import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from vcf_b200 import Codec
H, W = 2160, 3840
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.randint(0, 256, (ne, H, W, 3), dtype=torch.uint8).pin_memory()
idx = torch.empty((ne, H, W, 3), dtype=torch.uint8, pin_memory=True)
y = torch.empty((ne, H, W, 3), dtype=torch.uint8, pin_memory=True)
xn, idxn, yn = x.numpy(), idx.numpy(), y.numpy()
enc = Codec(block_size=8, q=32); dec = Codec(block_size=8, q=32, fp64=True)
def step():
    enc.encode(xn, out=idxn); dec.decode(idxn, (H, W), out=yn)
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(6): step()
dt = (time.perf_counter() - t0) / 6
print(os.environ.get("VCFB_CHUNK_MB"), ne, f"{dt*1e3:.2f} ms/step  {ne*H*W/dt/1e6:.0f} Mpx/s  {2*ne*H*W*3/dt/1e9:.1f} GB/s each way")
t0 = time.perf_counter()
for _ in range(6): enc.encode(xn, out=idxn)
dt = (time.perf_counter() - t0) / 6
print("  encode only", f"{dt*1e3:.2f} ms  {ne*H*W*3/dt/1e9:.1f} GB/s each way")
