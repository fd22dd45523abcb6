import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from vcf_b200 import Codec, _lib
H, W = 2160, 3840
what = sys.argv[1] if len(sys.argv) > 1 else "enc"
ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 64]
g = torch.Generator(device="cuda"); g.manual_seed(1)
for n in ns:
    x = torch.randint(0, 256, (n, H, W, 3), generator=g, device="cuda", dtype=torch.uint8)
    idx = torch.empty_like(x); y = torch.empty_like(x)
    enc = Codec(block_size=8, q=32); dec64 = Codec(block_size=8, q=32, fp64=True); dec32 = Codec(block_size=8, q=32, contract=True)
    enc.encode(x, out=idx)
    def run(f, reps=10):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    out = []
    if "enc" in what:
        ms = run(lambda: enc.encode(x, out=idx)); out.append(f"enc {ms:.3f} ms {n*H*W/ms/1e6:.1f} Gpx/s ({6*n*H*W/ms/1e6/6478.9*100:.1f}%)")
    if "d64" in what:
        ms = run(lambda: dec64.decode(idx, (H, W), out=y)); out.append(f"dec64 {ms:.3f} ms {n*H*W/ms/1e6:.1f} Gpx/s ({6*n*H*W/ms/1e6/6478.9*100:.1f}%)")
    if "d32" in what:
        ms = run(lambda: dec32.decode(idx, (H, W), out=y)); out.append(f"dec32 {ms:.3f} ms {n*H*W/ms/1e6:.1f} Gpx/s ({6*n*H*W/ms/1e6/6478.9*100:.1f}%)")
    print(os.environ.get("VCFB_ENC_CFG", "-"), n, _lib.last_kernel(), " | ".join(out), flush=True)
