"""Small workload touching every kernel family, for compute-sanitizer (memcheck) runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vcf_b200 import Codec, ColorCodec, _lib
rng = np.random.default_rng(0)
seen = set()
def rt(shape, B, q, off=0, **kw):
    n, H, W = shape
    buf = torch.zeros(n * H * W * 3 + 64, dtype=torch.uint8, device="cuda")
    x = buf[off:off + n * H * W * 3].view(n, H, W, 3)
    x.copy_(torch.from_numpy(rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)))
    enc = Codec(block_size=B, q=q, **kw); dec = Codec(block_size=B, q=q, **{k: v for k, v in kw.items() if k != "contract"}, fp64=True)
    idx, st = enc.encode(x, stats=True); seen.add(_lib.last_kernel())
    y, sd = dec.decode(idx, (H, W), original=x, stats=True); seen.add(_lib.last_kernel())
    y32 = Codec(block_size=B, q=q, **kw).decode(idx, (H, W)); seen.add(_lib.last_kernel())
    yf = dec.decode(idx, (H, W), return_float=True)
    return int(st.sum().item()) + int(sd.sum().item()) + int(y32.sum().item())
acc = 0
for shape in ((2, 16, 128), (3, 24, 384), (1, 40, 1024)):            # fast path (+ packed, f64 half-tile, stats kernels)
    for q in (8, 12):
        acc += rt(shape, 8, q); acc += rt(shape, 8, q, contract=True)
for B in (4, 8, 16, 32):                                             # general kernels: odd shapes, unaligned pointers
    for shape, off in (((1, 37, 53), 1), ((2, 64, 96), 3), ((1, 5, 3), 0)):
        acc += rt(shape, B, 8, off=off)
        acc += rt(shape, B, 5, off=off, disable_subbands=True)
    acc += rt((1, 48, 80), B, 4, perceptual=True)
    acc += rt((1, 48, 80), B, 32, color="YCrCb")
for color in ("YCoCg", "YCrCb"):
    cc = ColorCodec(color, 7)
    img = torch.from_numpy(rng.integers(0, 256, size=(33, 71, 3), dtype=np.uint8)).cuda()
    k = cc.encode(img); acc += int(cc.decode(k).sum().item()); seen.add(_lib.last_kernel())
x = rng.integers(0, 256, size=(2, 16, 128, 3), dtype=np.uint8)        # host API (chunked streams)
idx = Codec(8, 16).encode(x); acc += int(Codec(8, 16, fp64=True).decode(idx, (16, 128)).sum())
torch.cuda.synchronize()
print("ok", acc & 0xffff, sorted(seen))
