"""Per-block-size timing of the fused rate/distortion sweep (vcfb_rd_sweep_dev) on BASELINE configs[2]:
one 3840x2160 frame, 8 steps.  CUDA events on torch's current stream (the one the library launches on)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import vcf_oracle as O
from vcf_b200.rd import rd_stats_fused, rd_sweep

QS = (4, 8, 12, 16, 24, 32, 48, 64)


def timed(fn, reps=40):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
    x = torch.from_numpy(O.synthetic_frame(2160, 3840, 2, kind)).cuda()
    spin = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    for _ in range(300):            # bring the clocks up before the first timed loop
        spin.add_(1)
    torch.cuda.synchronize()
    out = {"content": kind, "frame": "3840x2160", "steps": list(QS)}
    for B in (4, 8, 16, 32):
        for hist in (True, False):
            out[f"B{B}_hist{int(hist)}_ms"] = round(timed(lambda: rd_stats_fused(x, B, QS, hist=hist)), 4)
    out["fused_total_ms"] = round(timed(lambda: [rd_stats_fused(x, B, QS) for B in (4, 8, 16, 32)]), 4)
    out["per_point_total_ms"] = round(timed(lambda: rd_sweep(x, fused=False), reps=5), 4)
    out["gpixel_s_fused"] = round(32 * 2160 * 3840 / out["fused_total_ms"] / 1e6, 1)
    print(json.dumps(out))


main()
