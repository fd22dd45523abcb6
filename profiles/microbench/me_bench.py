"""Row F3 measurement: full-search block matching (bs=16, sr=8) on 1920x1080 gray frame pairs.
   python profiles/microbench/me_bench.py [pairs]
Prints one JSON line: GPU Mpixel/s (CUDA events, inputs resident), the CPU oracle (vectorised
numpy restatement) and the reference's own loop form on one core for a cropped frame."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import me_oracle as M
from vcf_b200 import _lib
from vcf_b200.motion import block_matching

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, W, bs, sr = 1080, 1920, 16, 8
g = torch.Generator(device="cuda"); g.manual_seed(3)
ref = torch.randint(0, 256, (n, H, W), generator=g, device="cuda", dtype=torch.uint8)
cur = torch.roll(ref, shifts=(3, -5), dims=(1, 2)).contiguous()
cur += torch.randint(0, 3, cur.shape, generator=g, device="cuda", dtype=torch.uint8)
for _ in range(3):
    mv = block_matching(ref, cur, bs, sr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    mv = block_matching(ref, cur, bs, sr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
r0, c0 = ref[0].cpu().numpy(), cur[0].cpu().numpy()
t0 = time.perf_counter()
want = M.block_matching_full(r0, c0, bs, sr)
cpu_s = time.perf_counter() - t0
ok = bool(np.array_equal(mv[0].cpu().numpy(), want))
cand = (2 * sr + 1) ** 2
print(json.dumps({"workload": f"{n} pairs {W}x{H} gray, bs={bs}, sr={sr}, full search", "gpu_ms": ms,
                  "gpu_mpixel_s": n * H * W / 1e6 / (ms / 1e3), "gpu_gsad_s": n * (H // bs) * (W // bs) * cand * bs * bs / 1e9 / (ms / 1e3),
                  "cpu_oracle_mpixel_s_1core_vectorised": H * W / 1e6 / cpu_s, "matches_oracle": ok,
                  "kernel": _lib.last_kernel()}))
