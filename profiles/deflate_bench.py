"""GPU deflate of the index planes (row F4): time and size per quantisation step, one frame and 16 frames per
call, with the row geometry (vcfb_deflate_rows_dev: previous sample + samples above as match candidates) and
without (runs only), against zlib level 6 / Z_RLE on one host core.  Run on the B200:
    python profiles/deflate_bench.py > gpurun_out/deflate_bench.jsonl"""
import json, os, sys, time, zlib
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vcf_b200 import Codec, _lib
from vcf_b200.entropy import deflate_raw_dev
from bench import make_frames

L = _lib.lib()
H, W = 2160, 3840
dev = torch.device("cuda", 0)


def timed(kb, geom, reps=10):
    n = kb.numel()
    dst = torch.empty(L.vcfb_deflate_bound(n), dtype=torch.uint8, device=dev)
    ws = torch.empty(L.vcfb_deflate_workspace(n), dtype=torch.uint8, device=dev)
    nb = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _lib.check(L.vcfb_deflate_rows_dev(kb.data_ptr(), n, geom[0], geom[1], dst.data_ptr(), dst.numel(),
                                                      nb.data_ptr(), ws.data_ptr(), ws.numel(), st))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, int(nb.item()), dst


KINDS = tuple(os.environ.get("KINDS", "natural,noise").split(","))
for kind in KINDS:
    x = make_frames(torch, 16, H, W, dev, 99, kind)
    for q in (4, 8, 16, 32, 64):
        k = Codec(8, q).encode(x)
        Hp, Wp = k.shape[1], k.shape[2]
        geom = (3 * Wp, 3)
        host = k[0].cpu().numpy().tobytes()
        t0 = time.perf_counter(); z6 = len(zlib.compress(host, 6)); tz = time.perf_counter() - t0
        c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_RLE)
        zr = len(c.compress(host) + c.flush())
        for frames in (1, 16):
            kb = k[:frames].reshape(-1)
            row = {"content": kind, "q": q, "frames": frames, "input_MB": kb.numel() / 1e6}
            for name, g in (("rows", geom), ("runs", (0, 1))):
                ms, nbytes, dst = timed(kb, g)
                row[name] = {"ms": ms, "input_GB_s": kb.numel() / 1e6 / ms, "mpixel_s": frames * H * W / 1e3 / ms,
                             "bits_per_pixel": 8.0 * nbytes / (frames * H * W)}
                if frames == 1:
                    assert zlib.decompress(dst[:nbytes].cpu().numpy().tobytes(), -15) == host
                    row[name]["size_vs_zlib6"] = nbytes / z6
            if frames == 1:
                row.update({"zlib6_bits_per_pixel": 8.0 * z6 / (H * W), "zlib_rle_bits_per_pixel": 8.0 * zr / (H * W),
                            "zlib6_mpixel_s_one_core": H * W / 1e6 / tz})
            print(json.dumps(row), flush=True)
