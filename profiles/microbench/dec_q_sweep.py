"""Per-q timing of the float64 decoders of the B=8 fast path (development aid).
   python profiles/microbench/dec_q_sweep.py [frames]
Prints ms per launch over `frames` 4K frames for the two-tier kernel configurations and the
single-tier exact kernel, natural-like and noise content."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
from vcf_b200 import Codec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
H, W = 2160, 3840
nat = bench.make_frames(torch, n, dev, 1234)
noise = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device=dev)
y = torch.empty_like(nat)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cfgs = [("probed", None), ("2t", "8x1"), ("exact+dcskip", "9x1"), ("exact", "9x2")]
print(f"{n} frames {W}x{H}; ms per launch")
for name, x in (("natural", nat), ("noise", noise)):
    for q in (4, 8, 12, 16, 24, 32, 64):
        idx = Codec(block_size=8, q=q).encode(x)
        dec = Codec(block_size=8, q=q, fp64=True)
        row = []
        outs = []
        for label, cfg in cfgs:
            if cfg:
                os.environ["VCFB_DEC_CFG"] = cfg
            else:
                os.environ.pop("VCFB_DEC_CFG", None)
            row.append(timed(lambda: dec.decode(idx, (H, W), out=y)))
            outs.append(y.clone() if label in ("probed", "exact") else None)
        same = bool(torch.equal(outs[0], outs[-1]))
        ref = outs[-1]
        dec32 = Codec(block_size=8, q=q)
        for label, cfg in (("f32 12w", None), ("f32 8w", "8x1"), ("f32 16w", "4x4"), ("f32 pocketfft", "9x1")):
            if cfg:
                os.environ["VCFB_DEC32_CFG"] = cfg
            else:
                os.environ.pop("VCFB_DEC32_CFG", None)
            row.append(timed(lambda: dec32.decode(idx, (H, W), out=y)))
            if cfg is None:
                dmax = int((y.to(torch.int16) - ref.to(torch.int16)).abs().max().item())
                nbad = float((y != ref).float().mean().item())
        os.environ.pop("VCFB_DEC32_CFG", None)
        cfgs_all = cfgs + [("f32 12w", 0), ("f32 8w", 0), ("f32 16w", 0), ("f32 pocketfft", 0)]
        print(f"{name:8s} q={q:3d}  " + "  ".join(f"{l}: {v:6.3f}" for (l, _), v in zip(cfgs_all, row)) + f"  identical={same}  f32: max|d|={dmax} differing={nbad:.2e}")
os.environ.pop("VCFB_DEC_CFG", None)
