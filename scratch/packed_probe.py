import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 1:
    import torch
    from vcf_b200 import Codec
    from oracle import vcf_oracle as O
    H, W, n = 1080, 1920, 4
    frames = np.stack([O.synthetic_frame(H, W, 77 + i, "noise") for i in range(n)])
    got = Codec(block_size=8, q=1).encode(torch.from_numpy(frames).cuda()).cpu().numpy()
    np.save(sys.argv[1], got)
else:
    env = dict(os.environ)
    subprocess.check_call([sys.executable, __file__, "/tmp/p_packed.npy"], env=env)
    env["VCFB_ENC_SCALAR"] = "1"
    subprocess.check_call([sys.executable, __file__, "/tmp/p_scalar.npy"], env=env)
    a, b = np.load("/tmp/p_packed.npy"), np.load("/tmp/p_scalar.npy")
    bad = np.argwhere(a != b)
    print("mismatches", len(bad), "of", a.size)
    H, W = 1080, 1920
    ny, nx = H // 8, W // 8
    from collections import Counter
    cnt = Counter()
    for f, y, x, c in bad[:2000]:
        j, i = y // ny, x // nx
        cnt[(int(j), int(i), int(c))] += 1
    print(sorted(cnt.items(), key=lambda t: -t[1])[:20])
    for f, y, x, c in bad[:10]:
        print(f, y, x, c, a[f, y, x, c], b[f, y, x, c], "block", y % ny, x % nx, "lane-ish", (x % nx) % 16)
