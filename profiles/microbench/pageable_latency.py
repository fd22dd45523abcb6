"""Single-frame latency of the numpy (pageable) host API, the plugin's calling pattern."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from vcf_b200 import Codec
H, W = 2160, 3840
rng = np.random.default_rng(0)
x = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
enc = Codec(8, 32); dec = Codec(8, 32, fp64=True)
idx = enc.encode(x); y = dec.decode(idx, (H, W))
for name, f in (("encode", lambda: enc.encode(x)), ("decode", lambda: dec.decode(idx, (H, W)))):
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(10): f()
    print(name, f"{(time.perf_counter() - t0) * 100:.2f} ms per 4K frame (pageable numpy in/out)", os.environ.get("VCFB_STAGE", ""))
