"""Tensor-core tier of the B=8 path (csrc/kernels_tc.cu): the fast mode of BASELINE.json's north star.

* encoder (VCFB_F_FAST / Codec(fast=True)): fewer than 1e-6 of the indices differ from the reference's
  float32 path (the oracle), every difference is one quantisation step, and NONE sits at the four rational
  coefficient positions (0,0) (0,4) (4,0) (4,4) -- those are recomputed with pocketfft's own rounding
  sequence because that is where exact ties live (1 822 exact DC ties per 4K noise frame at
  q = 8);
* decoder: covered by tests/test_gpu_parity.py::test_fast_mode_float32_decoder_tolerances and
  tests/test_gpu_variants.py (pixels within +-1 LSB, PSNR within 0.01 dB).
Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vcf_oracle as O


def _codec(**kw):
    from vcf_b200 import Codec
    return Codec(**kw)


def _mismatch_census(got, ref):
    """(count, max |difference|, count at the rational positions) for subband-ordered index arrays."""
    d = got != ref
    n = int(d.sum())
    if n == 0:
        return 0, 0, 0
    ny, nx = ref.shape[-3] // 8, ref.shape[-2] // 8
    ys, xs = np.nonzero(d.any(axis=-1).reshape(-1, ref.shape[-3], ref.shape[-2]).any(axis=0))
    rat = sum(1 for y, x in zip(ys, xs) if (y // ny) in (0, 4) and (x // nx) in (0, 4))
    mx = int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max())
    return n, mx, rat


def test_fast_encoder_full_size_rate_and_rational_positions():
    """3840x2160, uniform noise (the tie-dense, dense-index case) and natural-like content, every q of the
    bench cycle: aggregate mismatch rate < 1e-6, differences of one step only, none at a rational position."""
    import torch
    from vcf_b200 import _lib
    H, W = 2160, 3840
    total = bad = 0
    for kind, seed in (("noise", 2), ("natural", 2)):
        img = O.synthetic_frame(H, W, seed, kind)
        x = torch.from_numpy(img).cuda()
        for q in (8, 16, 32, 64):
            ref = O.encode_array(img, 8, q)
            got = _codec(block_size=8, q=q, fast=True).encode(x).cpu().numpy()
            assert _lib.last_kernel() == "enc8_tc"
            n, mx, rat = _mismatch_census(got, ref)
            assert mx <= 1 and rat == 0, (kind, q, n, mx, rat)
            assert n <= 1.2e-6 * ref.size, (kind, q, n)          # per frame (one frame has 24.9 M indices)
            if kind == "natural":
                assert n <= 2, (q, n)                             # sparse indices: boundaries are rarely met at all
            total += ref.size
            bad += n
    assert bad < 1e-6 * total, (bad, total)


def test_fast_encoder_shapes_padding_batches():
    """Partial tiles (nx not a multiple of 128), vertical padding (rows above the frame come from the TMA zero
    fill), batches: a handful of indices per million may differ, never at a rational position; frames of a batch
    are independent of their neighbours."""
    import torch
    from vcf_b200 import _lib
    tot = bad = 0
    for si, (H, W) in enumerate(((64, 256), (8, 128), (120, 640), (67, 640), (270, 1920), (1080, 1920), (37, 2048))):
        frames = np.stack([O.synthetic_frame(H, W, 400 + 7 * si + i, "noise" if i % 2 else "natural") for i in range(3)])
        x = torch.from_numpy(frames).cuda()
        for q in (8, 32, 128):
            ref = np.stack([O.encode_array(f, 8, q) for f in frames])
            got = _codec(block_size=8, q=q, fast=True).encode(x)
            assert _lib.last_kernel() == "enc8_tc", (H, W, q)
            g = got.cpu().numpy()
            n, mx, rat = _mismatch_census(g, ref)
            assert mx <= 1 and rat == 0, (H, W, q, n, mx, rat)
            one = _codec(block_size=8, q=q, fast=True).encode(x[1:2]).cpu().numpy()
            assert np.array_equal(one[0], g[1])
            tot += ref.size
            bad += n
    assert bad <= 4e-6 * tot + 2, (bad, tot)


def test_fast_flag_falls_back_to_the_exact_encoder():
    """Requests the tensor-core encoder does not cover are served bit-exactly: q that is not a power of two or
    below 8, other block sizes, -p, -x, the float64 mode, unaligned widths."""
    import torch
    from vcf_b200 import _lib
    img = O.synthetic_frame(64, 256, 5, "noise")
    x = torch.from_numpy(img).cuda()
    for kw in (dict(q=12), dict(q=4), dict(q=1), dict(q=32, block_size=16), dict(q=32, perceptual=True),
               dict(q=32, disable_subbands=True), dict(q=32, fp64=True)):
        B = kw.pop("block_size", 8)
        q = kw.pop("q")
        dt = np.float64 if kw.get("fp64") else np.float32
        okw = {k: v for k, v in kw.items() if k != "fp64"}
        ref = O.encode_array(img, B, q, dtype=dt, **okw)
        got = _codec(block_size=B, q=q, fast=True, **kw).encode(x).cpu().numpy()
        assert _lib.last_kernel() != "enc8_tc", (B, q, kw)
        assert np.array_equal(got, ref), (B, q, kw)
    img2 = O.synthetic_frame(64, 200, 6, "natural")           # W % 16 != 0
    got = _codec(block_size=8, q=32, fast=True).encode(torch.from_numpy(img2).cuda()).cpu().numpy()
    assert np.array_equal(got, O.encode_array(img2, 8, 32))


def test_fast_encoder_statistics_and_host_api():
    """Statistics behind the tensor-core encoder come from the streaming pass over its indices; the numpy
    (host pointer) entry point takes the same flag."""
    import torch
    from vcf_b200 import _lib
    from vcf_b200.codec import stats_dict
    img = O.synthetic_frame(120, 640, 9, "natural")
    x = torch.from_numpy(img).cuda()
    c = _codec(block_size=8, q=16, fast=True)
    idx, st = c.encode(x, stats=True)
    k = idx.cpu().numpy().astype(np.int16) - 128
    s = stats_dict(st.cpu().numpy())
    assert s["nonzero"] == int((k != 0).sum()) and s["sumabs"] == int(np.abs(k).sum()) and s["nindices"] == k.size
    host = c.encode(img)
    assert np.array_equal(host, idx.cpu().numpy())
    # without the histogram the sums come out of the encoder's epilogue (no second pass): frames whose last tile of a
    # block row is partial (1152 / 8 = 144 blocks = 128 + 16), vertical padding, a batch
    for H, W, n in ((120, 1024, 1), (100, 1152, 3), (64, 2048, 2)):
        frames = np.stack([O.synthetic_frame(H, W, 90 + i, "noise" if i % 2 else "natural") for i in range(n)])
        launches0 = _lib.launch_count()
        idx2, st2 = _codec(block_size=8, q=8, fast=True, hist=False).encode(torch.from_numpy(frames).cuda(), stats=True)
        assert _lib.last_kernel() == "enc8_tc" and _lib.launch_count() - launches0 <= 2      # encoder + the index count
        k2 = idx2.cpu().numpy().astype(np.int16) - 128
        s2 = stats_dict(st2.cpu().numpy())
        assert s2["nonzero"] == int((k2 != 0).sum()) and s2["sumabs"] == int(np.abs(k2).sum()) and s2["nindices"] == k2.size, (H, W)
    # round trip in fast mode stays within +-1 LSB of the reference's round trip
    ref = O.decode_array(O.encode_array(img, 8, 16), img.shape, 8, 16)
    dec = c.decode(idx, img.shape[:2]).cpu().numpy()
    assert _lib.last_kernel() == "dec8_tc"
    assert np.abs(dec.astype(np.int16) - ref.astype(np.int16)).max() <= 1 + 16      # an index step moves a pixel by <= q / 8 ... loosely bounded
    assert abs(O.psnr(img, dec) - O.psnr(img, ref)) < 0.01
