"""Fused rate/distortion sweep (vcfb_rd_sweep_dev) against the per-point path
(encode + float64 decode with statistics) and against the oracle's restatement of the reference's
in-process loop (src/2D-DCT.py:533-579).  Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vcf_oracle as O

QS = (4, 8, 12, 16, 24, 32, 48, 64)


@pytest.fixture(scope="module")
def t():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _per_point(frames_t, B, q, color="YCoCg"):
    from vcf_b200 import Codec
    enc = Codec(block_size=B, q=q, color=color)
    dec = Codec(block_size=B, q=q, color=color, fp64=True)
    idx, se = enc.encode(frames_t, stats=True)
    _, sd = dec.decode(idx, frames_t.shape[-3:-1], original=frames_t, stats=True)
    return (se + sd).cpu().numpy()


@pytest.mark.parametrize("B", [2, 4, 8, 16, 32, 64, 128])
def test_fused_sweep_equals_per_point_path(B, t):
    """Every entry of the statistics vector (SSE per channel, signed difference, sample and index counts,
    non-zero count, sum |k|, the 3 x 256 histogram), every step, incl. steps whose indices wrap (q = 1)
    and non-integral steps; padded shapes; batches."""
    from vcf_b200 import _lib
    from vcf_b200.rd import rd_stats_fused
    shapes = [(128, 256), (67, 91), (130, 200), (1, 1)] if B <= 32 else [(128, 256), (67, 91)]
    qs = (1, 2.5) + QS
    for si, (H, W) in enumerate(shapes):
        for kind in ("natural", "noise"):
            x = t.from_numpy(np.stack([O.synthetic_frame(H, W, 700 + 10 * si + i, kind) for i in range(2)])).cuda()
            got = rd_stats_fused(x, B, qs).cpu().numpy()
            assert _lib.last_kernel() == ("rd_sweep" if 4 <= B <= 32 else "rd_sweep_anyb")
            for i, q in enumerate(qs):
                want = _per_point(x, B, q)
                bad = np.nonzero(got[i] != want)[0]
                assert bad.size == 0, (B, q, H, W, kind, bad[:8], got[i][bad[:8]], want[bad[:8]])
    x = t.from_numpy(O.synthetic_frame(128, 256, 777, "natural")).cuda()
    got = rd_stats_fused(x, B, (8, 32), color="YCrCb").cpu().numpy()
    for i, q in enumerate((8, 32)):
        assert np.array_equal(got[i], _per_point(x, B, q, "YCrCb")), (B, q)


@pytest.mark.parametrize("B", [2, 4, 8, 16, 32, 64, 128])
def test_nowrap_is_the_references_in_process_loop(B, t):
    """VCFB_F_NOWRAP: the dequantiser gets the quantiser's own indices (src/2D-DCT.py:565-568).  A bright
    frame at a small step makes the DC index of large blocks leave [-128, 127]."""
    from vcf_b200.codec import stats_dict
    from vcf_b200.rd import rd_stats_fused
    rng = np.random.default_rng(B)
    img = np.clip(O.synthetic_frame(128, 256, 800 + B, "natural").astype(np.int16) + 90, 0, 255).astype(np.uint8)
    img[:64] = np.clip(img[:64].astype(np.int16) - 200, 0, 255).astype(np.uint8)
    qs = (1, 3, 8, 32)
    wrapped_differs = False
    for offset in (0, 128):        # 0 = the loop as the reference runs it (VCFB_F_NO_OFFSET), 128 = as it reads
        got = rd_stats_fused(t.from_numpy(img).cuda(), B, qs, nowrap=True, no_offset=offset == 0).cpu().numpy()
        for i, q in enumerate(qs):
            k_u8, y, rm = O.optimize_block_size_point(img, B, q, offset=offset)
            st = stats_dict(got[i])
            d = img.astype(np.int64) - y.astype(np.int64)
            assert int(st["sse"].sum()) == int((d * d).sum()), (B, q, offset)
            assert st["sumdiff"] == int(d.sum()) and st["nsamples"] == img.size
            hist = np.stack([np.bincount(k_u8[..., c].ravel(), minlength=256) for c in range(3)])
            assert np.array_equal(st["hist"], hist), (B, q, offset)
            n = st["nsamples"]
            if offset:      # the distortion as the loop reads (:574): image still shifted by 128
                se = float(st["sse"].sum()) - 256.0 * st["sumdiff"] + 16384.0 * n
                nz, sabs, _ = O.index_stats(k_u8)
                assert st["nonzero"] == nz and st["sumabs"] == sabs
            else:           # as it runs: plain RMSE
                se = float(st["sse"].sum())
                k8 = k_u8.astype(np.int8).astype(np.int64)
                assert st["nonzero"] == int((k8 != 0).sum()) and st["sumabs"] == int(np.abs(k8).sum())
            assert abs(np.sqrt(se / n) - float(rm)) <= 2e-6 * float(rm) + 1e-6
            if offset:
                y_wrapped = O.decode_array(k_u8, img.shape, B, q)
                wrapped_differs |= not np.array_equal(y_wrapped, y)
            # the array handed to the entropy coder: the encoder with the same flag
            from vcf_b200 import Codec
            k_gpu = Codec(block_size=B, q=q, no_offset=offset == 0).encode(t.from_numpy(img).cuda()).cpu().numpy()
            assert np.array_equal(k_gpu, k_u8), (B, q, offset)
    assert wrapped_differs                  # the test does exercise the difference


def test_config3_sweep_full_size_fused(t):
    """BASELINE configs[2] on the 4K frame: B in {4,8,16,32} x 8 steps, fused == per-point."""
    import time
    from vcf_b200.rd import rd_sweep
    x = t.from_numpy(O.synthetic_frame(2160, 3840, 2, "natural")).cuda()
    for fused in (True, False):     # warm-up (module load, allocator)
        rd_sweep(x, fused=fused, qs=(8,))
    res = {}
    for fused in (True, False):
        t.cuda.synchronize()
        t0 = time.perf_counter()
        res[fused] = rd_sweep(x, fused=fused)
        t.cuda.synchronize()
        res[fused, "s"] = time.perf_counter() - t0
    assert len(res[True]) == 32
    for a, b in zip(res[True], res[False]):
        assert a["B"] == b["B"] and a["q"] == b["q"] and a["sse"] == b["sse"] and a["nonzero"] == b["nonzero"]
        assert a["bpp_entropy"] == b["bpp_entropy"] and a["rmse"] == b["rmse"]
    print(f"\nconfig 3 sweep: fused {res[True, 's'] * 1e3:.2f} ms, per point {res[False, 's'] * 1e3:.2f} ms")
    # no assertion on the wall clock of a shared box: `workloads.c3` of bench.py times both paths with CUDA events
