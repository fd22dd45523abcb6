import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from oracle import vcf_oracle as O
    from vcf_b200 import Codec, _lib
    H, W, n, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    frames = np.stack([O.synthetic_frame(H, W, 5 + i, "natural") for i in range(n)])
    idx = np.stack([O.encode_array(f, 8, 32) for f in frames])
    ref = np.stack([O.decode_array(k, (H, W, 3), 8, 32) for k in idx])
    got = Codec(block_size=8, q=32, fp64=(mode == "f64")).decode(torch.from_numpy(idx).cuda(), (H, W))
    torch.cuda.synchronize()
    d = np.abs(got.cpu().numpy().astype(int) - ref.astype(int))
    print(H, W, n, mode, _lib.last_kernel(), "maxdiff", d.max(), "nbad", int((d > 0).sum()))
else:
    for (H, W, n) in [(8, 128, 1), (8, 128, 3), (24, 384, 1), (16, 512, 1), (8, 768, 1), (13, 256, 1), (1080, 1920, 1)]:
        for mode in ("f32", "f64"):
            r = subprocess.run([sys.executable, __file__, str(H), str(W), str(n), mode], capture_output=True, text=True)
            print((r.stdout.strip() or "FAIL: " + r.stderr.strip().splitlines()[-1][:200]))
