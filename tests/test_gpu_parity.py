"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the
golden vectors recorded from the unmodified reference.  Needs a B200."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vcf_oracle as O
from _util import golden_filter as _gfilter, golden_kw as _gkw, parse_flags as _parse

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _codec(**kw):
    from vcf_b200 import Codec
    return Codec(**kw)


# ---------------------------------------------------------------------------
# golden vectors of the reference (bit-exact, float32 encode / float64 decode)
# ---------------------------------------------------------------------------
FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "ref_flow_*.npz")) if "sa_" not in f)


@pytest.mark.parametrize("fn", FILES, ids=[os.path.basename(f)[9:-4] for f in FILES])
def test_reference_golden_bit_exact(fn, torch_cuda):
    g = np.load(fn)
    kw = _gkw(g)
    c32 = _codec(block_size=kw["B"], q=kw["q"], perceptual=kw["perceptual"],
                 disable_subbands=kw["disable_subbands"])
    idx = c32.encode(g["img"])
    assert idx.dtype == np.uint8 and idx.shape == g["idx"].shape
    assert np.array_equal(idx, g["idx"])               # reference's own float32 path
    c64 = _codec(block_size=kw["B"], q=kw["q"], perceptual=kw["perceptual"],
                 disable_subbands=kw["disable_subbands"], fp64=True)
    if _gfilter(g) == "gaussian_blur":
        # -f: the un-clipped float64 image goes through the chain's filter (src/2D-DCT.py:461, src/gaussian_blur.py)
        import cv2
        _, yf = c64.decode(g["idx"], g["img"].shape, return_float=True)
        assert yf.dtype == np.float64
        dec = np.clip(cv2.GaussianBlur(yf, (5, 5), 0), 0, 255).astype(np.uint8)
        assert np.array_equal(dec, g["decoded"])
        return
    dec = c64.decode(g["idx"], g["img"].shape)
    assert np.array_equal(dec, g["decoded"])           # reference's float64 decode chain
    # same through the torch / device-pointer entry points
    t = torch_cuda
    idx_t = c32.encode(t.from_numpy(g["img"]).cuda())
    assert np.array_equal(idx_t.cpu().numpy(), g["idx"])
    dec_t = c64.decode(t.from_numpy(g["idx"]).cuda(), g["img"].shape)
    assert np.array_equal(dec_t.cpu().numpy(), g["decoded"])


# ---------------------------------------------------------------------------
# seeded inputs vs the oracle
# ---------------------------------------------------------------------------
SHAPES = [(64, 96), (67, 91), (8, 8), (5, 3), (1, 1), (300, 520), (33, 1000)]


@pytest.mark.parametrize("B", [4, 8, 16, 32])
@pytest.mark.parametrize("q", [1, 5, 8, 12, 32, 64])
def test_encode_exact_fp32_fp64(B, q, torch_cuda):
    for si, (H, W) in enumerate(SHAPES):
        for kind in ("noise", "natural"):
            img = O.synthetic_frame(H, W, 100 + si, kind)
            for fp64, dt in ((False, np.float32), (True, np.float64)):
                ref = O.encode_array(img, B, q, dtype=dt)
                got = _codec(block_size=B, q=q, fp64=fp64).encode(img)
                assert got.shape == ref.shape
                nbad = int((got != ref).sum())
                assert nbad == 0, (B, q, H, W, kind, fp64, nbad)


@pytest.mark.parametrize("B", [4, 8, 16, 32])
@pytest.mark.parametrize("q", [1, 8, 12, 32])
def test_decode_fp64_bit_exact_and_fp32_within_1lsb(B, q, torch_cuda):
    for si, (H, W) in enumerate(SHAPES):
        img = O.synthetic_frame(H, W, 200 + si, "natural")
        idx = O.encode_array(img, B, q)
        ref = O.decode_array(idx, img.shape, B, q)
        got64 = _codec(block_size=B, q=q, fp64=True).decode(idx, img.shape)
        assert np.array_equal(got64, ref), (B, q, H, W)
        got32 = _codec(block_size=B, q=q).decode(idx, img.shape)
        d = np.abs(got32.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1, (B, q, H, W, int(d.max()))
        if H * W >= 4096:
            # The float64 decoder above is the parity-grade one (0 LSB, 0 dB).  The
            # float32 decoder is an opt-in fast mode: reconstructed samples are often
            # exact integers (every pixel of a DC-only block is q*k/B) and the
            # truncation of src/2D-DCT.py:466 then follows the last-ulp error of the
            # float64 chain, which float32 cannot reproduce.  It
            # stays within +-1 LSB; its PSNR is only bounded, not matched to 0.01 dB.
            assert abs(O.psnr(img, got32) - O.psnr(img, ref)) < 0.1, (B, q, H, W)


@pytest.mark.parametrize("B", [2, 64, 128])
def test_block_sizes_of_the_L_search_without_a_codelet(B, torch_cuda):
    """B = 2, 64, 128 -- the rest of the set `optimize_block_size` tries (src/2D-DCT.py:538) -- run as
    interpreted pocketfft programs (csrc/kernels_anyb.cu): bit-exact like every other size, every flag."""
    from vcf_b200 import _lib
    shapes = [(128, 256), (67, 91), (5, 3), (130, 200)]
    for si, (H, W) in enumerate(shapes):
        for kind in ("noise", "natural"):
            img = O.synthetic_frame(H, W, 900 + si, kind)
            for q in (1, 12, 32):
                for fp64, dt in ((False, np.float32), (True, np.float64)):
                    ref = O.encode_array(img, B, q, dtype=dt)
                    got = _codec(block_size=B, q=q, fp64=fp64).encode(img)
                    assert _lib.lib().vcfb_last_kernel() == b"encode_anyb"
                    assert got.shape == ref.shape and np.array_equal(got, ref), (B, q, H, W, kind, fp64)
                idx = O.encode_array(img, B, q)
                ref = O.decode_array(idx, img.shape, B, q)
                got = _codec(block_size=B, q=q, fp64=True).decode(idx, img.shape)
                assert _lib.lib().vcfb_last_kernel() == b"decode_anyb"
                assert np.array_equal(got, ref), (B, q, H, W, kind)
    img = O.synthetic_frame(128, 256, 950, "natural")
    for kw in (dict(disable_subbands=True), dict(perceptual=True), dict(color="YCrCb")):
        ref = O.encode_array(img, B, 8, **kw)
        got = _codec(block_size=B, q=8, **kw).encode(img)
        assert np.array_equal(got, ref), (B, kw)
        refd = O.decode_array(ref, img.shape, B, 8, **kw)
        gotd = _codec(block_size=B, q=8, fp64=True, **kw).decode(ref, img.shape)
        assert np.array_equal(gotd, refd), (B, kw)
    # a batch, statistics on both sides, float output
    t = torch_cuda
    batch = np.stack([O.synthetic_frame(128, 128, 960 + i, "natural") for i in range(3)])
    enc, dec = _codec(block_size=B, q=16), _codec(block_size=B, q=16, fp64=True)
    idx, se = enc.encode(t.from_numpy(batch).cuda(), stats=True)
    ref = np.stack([O.encode_array(f, B, 16) for f in batch])
    assert np.array_equal(idx.cpu().numpy(), ref)
    out, yf, sd = dec.decode(idx, (128, 128), original=t.from_numpy(batch).cuda(), stats=True, return_float=True)
    refd = np.stack([O.decode_array(k, (128, 128, 3), B, 16) for k in ref])
    assert np.array_equal(out.cpu().numpy(), refd)
    from vcf_b200.codec import stats_dict
    st = stats_dict((se + sd).cpu().numpy())
    d = batch.astype(np.int64) - refd.astype(np.int64)
    assert int(st["sse"].sum()) == int((d * d).sum()) and st["nsamples"] == batch.size
    assert st["nonzero"] == int((ref != 128).sum()) and st["nindices"] == ref.size
    assert np.array_equal(np.clip(yf.cpu().numpy(), 0, 255).astype(np.uint8), refd)


def test_random_indices_decode_fp64(torch_cuda):
    """Arbitrary index arrays (not produced by an encoder), incl. the int16 wrap
    of q*k and saturating clips."""
    rng = np.random.default_rng(7)
    for B in (4, 8, 16, 32):
        for q in (3, 32, 300):
            H, W = 64, 96
            idx = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
            ref = O.decode_array(idx, (H, W, 3), B, q)
            got = _codec(block_size=B, q=q, fp64=True).decode(idx, (H, W))
            assert np.array_equal(got, ref), (B, q)


@pytest.mark.parametrize("B", [4, 8, 16, 32])
def test_flags_nosub_perceptual(B, torch_cuda):
    img = O.synthetic_frame(72, 104, 300 + B, "natural")
    for kw in (dict(disable_subbands=True), dict(perceptual=True), dict(perceptual=True, disable_subbands=True)):
        for q in (4, 32):
            for fp64, dt in ((False, np.float32), (True, np.float64)):
                ref = O.encode_array(img, B, q, dtype=dt, **kw)
                got = _codec(block_size=B, q=q, fp64=fp64, **kw).encode(img)
                assert np.array_equal(got, ref), (B, q, kw, fp64)
            idx = O.encode_array(img, B, q, **kw)
            ref = O.decode_array(idx, img.shape, B, q, **kw)
            got = _codec(block_size=B, q=q, fp64=True, **kw).decode(idx, img.shape)
            assert np.array_equal(got, ref), (B, q, kw)


@pytest.mark.parametrize("B", [8, 16])
def test_ycrcb_extension(B, torch_cuda):
    img = O.synthetic_frame(80, 112, 400 + B, "natural")
    for q in (8, 32):
        for fp64, dt in ((False, np.float32), (True, np.float64)):
            ref = O.encode_array(img, B, q, dtype=dt, color="YCrCb")
            got = _codec(block_size=B, q=q, fp64=fp64, color="YCrCb").encode(img)
            assert np.array_equal(got, ref), (B, q, fp64)
        idx = O.encode_array(img, B, q, color="YCrCb")
        ref = O.decode_array(idx, img.shape, B, q, color="YCrCb")
        got = _codec(block_size=B, q=q, fp64=True, color="YCrCb").decode(idx, img.shape)
        assert np.array_equal(got, ref), (B, q)


def test_contract_mode_mismatch_rate(torch_cuda):
    """VCFB_F_CONTRACT (fused multiply-adds allowed): < 1e-6 of the indices may
    differ from the reference's float32 path, none at the rational positions
    {0,B/2}^2, and only by one step."""
    H, W = 1080, 1920
    for kind in ("noise", "natural"):
        img = O.synthetic_frame(H, W, 500, kind)
        for B, q in ((8, 8), (8, 32), (16, 8)):
            ref = O.encode_array(img, B, q)
            got = _codec(block_size=B, q=q, contract=True).encode(img)
            bad = np.argwhere(got != ref)
            assert len(bad) <= max(1, int(1e-6 * ref.size)) + 2, (kind, B, q, len(bad))
            ny, nx = H // B if H % B == 0 else (H + B - 1) // B, W // B
            for (y, x, c) in bad:
                j, i = y // ny, x // nx
                assert not (j in (0, B // 2) and i in (0, B // 2))
                assert abs(int(got[y, x, c]) - int(ref[y, x, c])) in (1, 255)


def test_float_output_and_stats(torch_cuda):
    img = O.synthetic_frame(120, 200, 600, "natural")
    B, q = 8, 16
    idx = O.encode_array(img, B, q)
    c64 = _codec(block_size=B, q=q, fp64=True)
    rgb, yf, st = c64.decode(idx, img.shape, original=img, stats=True, return_float=True)
    ref_f = O.decode_array(idx, img.shape, B, q, return_float=True)
    assert yf.dtype == np.float64 and np.array_equal(yf, ref_f)
    assert int(st["sse"].sum()) == O.sse_int(img, rgb)
    for c in range(3):
        assert int(st["sse"][c]) == O.sse_int(img[..., c], rgb[..., c])
    assert st["nsamples"] == img.size
    assert st["sumdiff"] == int((img.astype(np.int64) - rgb.astype(np.int64)).sum())
    assert abs(st["rmse"] - float(O.rmse(img, rgb))) < 1e-4
    got, se = _codec(block_size=B, q=q).encode(img, stats=True)
    nz, sabs, hist = O.index_stats(got)
    assert se["nonzero"] == nz and se["sumabs"] == sabs and se["nindices"] == got.size
    assert np.array_equal(se["hist"], hist)


def test_batch_equals_single_frames_and_unaligned_pointers(torch_cuda):
    t = torch_cuda
    n, H, W = 5, 70, 150
    frames = np.stack([O.synthetic_frame(H, W, 700 + i, "natural") for i in range(n)])
    c = _codec(block_size=8, q=8)
    ref = np.stack([O.encode_array(f, 8, 8) for f in frames])
    assert np.array_equal(c.encode(frames), ref)
    # device pointers at every byte alignment
    for off in (0, 1, 2, 3, 5, 8, 13):
        buf = t.zeros(frames.size + 64, dtype=t.uint8, device="cuda")
        view = buf[off:off + frames.size].view(n, H, W, 3)
        view.copy_(t.from_numpy(frames))
        got = c.encode(view)
        assert np.array_equal(got.cpu().numpy(), ref), off
        ibuf = t.zeros(ref.size + 64, dtype=t.uint8, device="cuda")
        iview = ibuf[off:off + ref.size].view(*ref.shape)
        iview.copy_(t.from_numpy(ref))
        dec = _codec(block_size=8, q=8, fp64=True).decode(iview, (H, W))
        refd = np.stack([O.decode_array(r, (H, W, 3), 8, 8) for r in ref])
        assert np.array_equal(dec.cpu().numpy(), refd), off


def test_full_size_4k_properties(torch_cuda):
    """BASELINE config 2 size (3840x2160): exact parity on one frame per q, plus
    size-independent properties on the batch."""
    t = torch_cuda
    H, W = 2160, 3840
    img = O.synthetic_frame(H, W, 2, "natural")
    for q in (8, 16, 32, 64):
        ref = O.encode_array(img, 8, q)
        c = _codec(block_size=8, q=q)
        got = c.encode(t.from_numpy(img).cuda())
        assert np.array_equal(got.cpu().numpy(), ref), q
        dec = _codec(block_size=8, q=q, fp64=True).decode(got, (H, W))
        refd = O.decode_array(ref, img.shape, 8, q)
        assert np.array_equal(dec.cpu().numpy(), refd), q
        assert abs(O.psnr(img, dec.cpu().numpy()) - O.psnr(img, refd)) == 0.0
        dec32 = c.decode(got, (H, W)).cpu().numpy()      # opt-in fast mode: +-1 LSB only
        assert np.abs(dec32.astype(np.int16) - refd.astype(np.int16)).max() <= 1
        assert abs(O.psnr(img, dec32) - O.psnr(img, refd)) < 0.1
    # B=32 pads 2160 -> 2176 (8 zero rows top and bottom)
    ref = O.encode_array(img, 32, 32)
    got = _codec(block_size=32, q=32).encode(t.from_numpy(img).cuda())
    assert got.shape == (2176, 3840, 3) and np.array_equal(got.cpu().numpy(), ref)
    # re-encoding a decoded frame with the same q is stable after one round (idempotence)
    c = _codec(block_size=8, q=32)
    d1 = c.decode(c.encode(t.from_numpy(img).cuda()), (H, W))
    d2 = c.decode(c.encode(d1), (H, W))
    d3 = c.decode(c.encode(d2), (H, W))
    assert (d3 != d2).float().mean().item() <= (d2 != d1).float().mean().item() + 1e-9


def test_errors_are_loud(torch_cuda):
    from vcf_b200 import VcfbError, Codec
    img = np.zeros((16, 16, 3), np.uint8)
    with pytest.raises(ValueError):
        Codec(block_size=7)
    with pytest.raises(ValueError):
        Codec(q=0)
    with pytest.raises(ValueError):
        Codec(block_size=8).decode(np.zeros((8, 8, 3), np.uint8), (16, 16))
    from vcf_b200 import _lib
    L = _lib.lib()
    assert L.vcfb_encode_dev(None, 1, 16, 16, 8, 32.0, 0, 0, None, None, None, None) < 0
    assert b"NULL" in L.vcfb_last_error()


FAST_SHAPES = [(8, 128), (24, 384), (16, 512), (40, 1024), (13, 256), (1080, 1920), (61, 640)]


@pytest.mark.parametrize("q", [1, 8, 12, 32, 64, 100])
def test_fast_encode_path_exact(q, torch_cuda):
    """The TMA fast path (B=8, W%16==0, nx%16==0) must be taken and must be bit-exact,
    including partial tiles (W % 384 != 0), vertical padding (H % 8 != 0) and batches."""
    from vcf_b200 import _lib
    t = torch_cuda
    for si, (H, W) in enumerate(FAST_SHAPES):
        n = 3 if H * W < 200000 else 1
        frames = np.stack([O.synthetic_frame(H, W, 900 + 10 * si + i, "noise" if i % 2 else "natural") for i in range(n)])
        ref = np.stack([O.encode_array(f, 8, q) for f in frames])
        got = _codec(block_size=8, q=q).encode(t.from_numpy(frames).cuda())
        assert _lib.last_kernel() == "enc8_fast", (H, W, _lib.last_kernel())
        assert np.array_equal(got.cpu().numpy(), ref), (H, W, q, int((got.cpu().numpy() != ref).sum()))
    # not eligible -> general kernel, still exact
    img = O.synthetic_frame(64, 96, 5, "natural")
    got = _codec(block_size=8, q=q).encode(t.from_numpy(img).cuda())
    assert _lib.last_kernel() == "encode_general"
    assert np.array_equal(got.cpu().numpy(), O.encode_array(img, 8, q))


@pytest.mark.parametrize("color", ["YCoCg", "YCrCb"])
@pytest.mark.parametrize("q", [1, 8, 12, 32, 100])
def test_fast16_encode_path_exact(q, color, torch_cuda):
    """B = 16 fast path (kernels_b16.cu): bit-exact indices and statistics, both colour
    transforms, batches, vertical padding (TMA zero fill), widths that are multiples of 256."""
    from vcf_b200 import _lib
    from vcf_b200.codec import stats_dict
    t = torch_cuda
    for si, (H, W) in enumerate(((16, 256), (48, 512), (40, 768), (64, 1280), (272, 3840))):
        n = 3 if H * W < 100000 else 1
        frames = np.stack([O.synthetic_frame(H, W, 1500 + 10 * si + i, "noise" if i % 2 else "natural") for i in range(n)])
        ref = np.stack([O.encode_array(f, 16, q, color=color) for f in frames])
        enc = _codec(block_size=16, q=q, color=color, hist=False)
        got, st = enc.encode(t.from_numpy(frames).cuda(), stats=True)
        assert _lib.last_kernel() == "enc16_fast", (H, W, _lib.last_kernel())
        assert np.array_equal(got.cpu().numpy(), ref), (H, W, q, color, int((got.cpu().numpy() != ref).sum()))
        nz, sabs, _ = O.index_stats(ref)
        s = stats_dict(st.cpu().numpy())
        assert s["nonzero"] == nz and s["sumabs"] == sabs and s["nindices"] == ref.size
        got2 = enc.encode(t.from_numpy(frames).cuda())
        assert np.array_equal(got2.cpu().numpy(), ref)
        if si == 1:          # with the histogram: fast kernel + streaming pass over the indices
            got3, st3 = _codec(block_size=16, q=q, color=color, hist=True).encode(t.from_numpy(frames).cuda(), stats=True)
            assert _lib.last_kernel() == "enc16_fast"
            s3 = stats_dict(st3.cpu().numpy())
            assert np.array_equal(got3.cpu().numpy(), ref) and np.array_equal(s3["hist"], O.index_stats(ref)[2])
            assert s3["nonzero"] == nz and s3["sumabs"] == sabs and s3["nindices"] == ref.size
    # a width that is not a multiple of 256 takes the general kernel
    img = O.synthetic_frame(32, 320, 3, "natural")
    got = _codec(block_size=16, q=q, color=color).encode(t.from_numpy(img).cuda())
    assert _lib.last_kernel() == "encode_general"
    assert np.array_equal(got.cpu().numpy(), O.encode_array(img, 16, q, color=color))


@pytest.mark.parametrize("color", ["YCoCg", "YCrCb"])
@pytest.mark.parametrize("q", [1, 8, 12, 32, 255])
def test_fast16_decode_path_exact(q, color, torch_cuda):
    """B = 16 fast path, float64 decoder: bit-exact pixels and distortion statistics, both colour
    transforms, batches, arbitrary (non-encoder) index arrays."""
    from vcf_b200 import _lib
    from vcf_b200.codec import stats_dict
    t = torch_cuda
    rng = np.random.default_rng(q)
    for si, (H, W) in enumerate(((16, 256), (48, 512), (64, 768), (272, 3840))):
        n = 3 if H * W < 100000 else 1
        frames = np.stack([O.synthetic_frame(H, W, 1600 + 10 * si + i, "noise" if i % 2 else "natural") for i in range(n)])
        idx = np.stack([O.encode_array(f, 16, q, color=color) for f in frames])
        if si % 2:
            idx[-1] = rng.integers(0, 256, size=idx[-1].shape, dtype=np.uint8)
        ref = np.stack([O.decode_array(k, (H, W, 3), 16, q, color=color) for k in idx])
        dec = _codec(block_size=16, q=q, color=color, fp64=True)
        x = t.from_numpy(frames).cuda()
        got, st = dec.decode(t.from_numpy(idx).cuda(), (H, W), original=x, stats=True)
        assert _lib.last_kernel() == "dec16_fast", (H, W, _lib.last_kernel())
        assert np.array_equal(got.cpu().numpy(), ref), (H, W, q, color, int((got.cpu().numpy() != ref).sum()))
        s = stats_dict(st.cpu().numpy())
        for c in range(3):
            assert int(s["sse"][c]) == O.sse_int(frames[..., c], ref[..., c])
        assert s["nsamples"] == frames.size
        assert s["sumdiff"] == int((frames.astype(np.int64) - ref.astype(np.int64)).sum())
        got2 = dec.decode(t.from_numpy(idx).cuda(), (H, W))
        assert np.array_equal(got2.cpu().numpy(), ref)
    # index arrays whose non-zero extent changes from one group of 8 blocks to the next: the pruned
    # codelet (dct16_inv_low4), its column skip, and the switch back to the full one
    H, W = 64, 1024
    ny, nx = H // 16, W // 16
    kk = np.zeros((ny, nx, 16, 16, 3), dtype=np.int64)
    kk[:, :, 0, 0, :] = rng.integers(-40, 41, size=(ny, nx, 3))
    for by in range(ny):
        for g0 in range(0, nx, 8):
            nu, ni = int(rng.choice([1, 2, 4, 5, 16])), int(rng.choice([1, 2, 4, 5, 16]))
            kk[by, g0:g0 + 8, :nu, :ni, :] = rng.integers(-9, 10, size=(8, nu, ni, 3))
    idx = (kk.transpose(2, 0, 3, 1, 4).reshape(H, W, 3) + 128).astype(np.uint8)
    ref = O.decode_array(idx, (H, W, 3), 16, q, color=color)
    got = _codec(block_size=16, q=q, color=color, fp64=True).decode(t.from_numpy(idx).cuda(), (H, W))
    assert _lib.last_kernel() == "dec16_fast"
    assert np.array_equal(got.cpu().numpy(), ref), ("extent mix", q, color, int((got.cpu().numpy() != ref).sum()))
    # vertical padding takes the general kernel
    img = O.synthetic_frame(40, 256, 3, "natural")
    k = O.encode_array(img, 16, q, color=color)
    got = _codec(block_size=16, q=q, color=color, fp64=True).decode(t.from_numpy(k).cuda(), (40, 256))
    assert _lib.last_kernel() == "decode_general"
    assert np.array_equal(got.cpu().numpy(), O.decode_array(k, img.shape, 16, q, color=color))


@pytest.mark.parametrize("B", [32])
@pytest.mark.parametrize("color", ["YCoCg", "YCrCb"])
def test_tile_fast_path_b32(B, color, torch_cuda):
    """B = 32 fast path (csrc/kernels_tile.cu): float32 encoder and float64 decoder bit-exact, float32
    decoder within +-1 LSB, statistics, vertical padding (2160 rows -> 2176 at B = 32), batches, wrapped indices."""
    from vcf_b200 import _lib
    from vcf_b200.codec import stats_dict
    t = torch_cuda
    rng = np.random.default_rng(B)
    for si, (H, W) in enumerate(((64, 256), (40, 384), (100, 128), (96, 1280))):
        n = 2
        frames = np.stack([O.synthetic_frame(H, W, 5000 + 10 * si + i, "noise" if i % 2 else "natural") for i in range(n)])
        x = t.from_numpy(frames).cuda()
        for q in (1, 8, 12, 32):
            ref = np.stack([O.encode_array(f, B, q, color=color) for f in frames])
            got, se = _codec(block_size=B, q=q, color=color, hist=False).encode(x, stats=True)
            assert _lib.last_kernel() == f"enc{B}_tile", _lib.last_kernel()
            assert np.array_equal(got.cpu().numpy(), ref), (B, q, H, W, int((got.cpu().numpy() != ref).sum()))
            got_h, se_h = _codec(block_size=B, q=q, color=color, hist=True).encode(x, stats=True)
            assert np.array_equal(got_h.cpu().numpy(), ref)
            idx = ref.copy()
            if si % 2:
                idx[-1] = rng.integers(0, 256, size=idx[-1].shape, dtype=np.uint8)
            refd = np.stack([O.decode_array(k, (H, W, 3), B, q, color=color) for k in idx])
            dec = _codec(block_size=B, q=q, color=color, fp64=True)
            gotd, sd = dec.decode(t.from_numpy(idx).cuda(), (H, W), original=x, stats=True)
            assert _lib.last_kernel() == f"dec{B}_tile", _lib.last_kernel()
            assert np.array_equal(gotd.cpu().numpy(), refd), (B, q, H, W, int((gotd.cpu().numpy() != refd).sum()))
            assert np.array_equal(dec.decode(t.from_numpy(idx).cuda(), (H, W)).cpu().numpy(), refd)
            s = stats_dict((se_h + sd).cpu().numpy())
            nz, sabs, hist = O.index_stats(ref)
            assert s["nonzero"] == nz and s["sumabs"] == sabs and np.array_equal(s["hist"], hist) and s["nindices"] == ref.size
            s0 = stats_dict(se.cpu().numpy())
            assert s0["nonzero"] == nz and s0["sumabs"] == sabs and s0["nindices"] == ref.size
            for c in range(3):
                assert int(s["sse"][c]) == O.sse_int(frames[..., c], refd[..., c])
            assert s["nsamples"] == frames.size
            assert s["sumdiff"] == int((frames.astype(np.int64) - refd.astype(np.int64)).sum())
            got32 = _codec(block_size=B, q=q, color=color).decode(t.from_numpy(ref).cuda(), (H, W)).cpu().numpy()
            assert _lib.last_kernel() == f"dec{B}_tile_f32"
            ref32 = np.stack([O.decode_array(k, (H, W, 3), B, q, color=color) for k in ref])
            assert np.abs(got32.astype(np.int16) - ref32.astype(np.int16)).max() <= 1
    # numpy entry points (host layer) and a non-integral step
    img = O.synthetic_frame(64, 256, 5100, "natural")
    k = _codec(block_size=B, q=2.5, color=color).encode(img)
    assert np.array_equal(k, O.encode_array(img, B, 2.5, color=color))
    y = _codec(block_size=B, q=2.5, color=color, fp64=True).decode(k, img.shape)
    assert np.array_equal(y, O.decode_array(k, img.shape, B, 2.5, color=color))
    # widths outside the fast path keep the general kernel
    img = O.synthetic_frame(64, 96, 5101, "natural")
    _codec(block_size=B, q=8, color=color).encode(t.from_numpy(img).cuda())
    assert _lib.last_kernel() == "encode_general"


@pytest.mark.parametrize("color", ["YCoCg", "YCrCb"])
def test_fast16_float32_decoder_tolerances(color, torch_cuda):
    """B = 16 fast path, float32 decoder (the fast mode): within +-1 LSB of the reference's float64 chain and
    PSNR within 0.01 dB, on natural, noise and smooth content; blocks with DC indices only are bit-exact."""
    from vcf_b200 import _lib
    t = torch_cuda
    H, W = 256, 3840
    for kind, q in (("natural", 8), ("natural", 32), ("noise", 8), ("noise", 32), ("smooth", 32), ("natural", 200)):
        if kind == "smooth":
            yy, xx = np.mgrid[0:H, 0:W]
            img = np.stack([110 + 60 * np.sin(xx / 400.0) * np.cos(yy / 300.0), 90 + 0.02 * xx, 140 - 0.1 * yy], axis=-1)
            img = np.clip(img, 0, 255).astype(np.uint8)
        else:
            img = O.synthetic_frame(H, W, 4000 + q, kind)
        idx = O.encode_array(img, 16, q, color=color)
        ref = O.decode_array(idx, img.shape, 16, q, color=color)
        got = _codec(block_size=16, q=q, color=color).decode(t.from_numpy(idx).cuda(), (H, W)).cpu().numpy()
        assert _lib.last_kernel() == "dec16_f32"
        d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1, (kind, q, int(d.max()))
        assert abs(O.psnr(img, got) - O.psnr(img, ref)) < 0.01, (kind, q, O.psnr(img, got), O.psnr(img, ref))
        # DC-only blocks: bit-exact
        k = idx.astype(np.int16) - 128
        blk = k.reshape(16, H // 16, 16, W // 16, 3)                  # subband layout [j][y][i][x][c]
        ac = np.abs(blk).sum(axis=(0, 2, 4)) - np.abs(blk[0, :, 0]).sum(axis=-1)
        dc_only = np.repeat(np.repeat(ac == 0, 16, axis=0), 16, axis=1)
        assert np.array_equal(got[dc_only], ref[dc_only]), (kind, q)
        if kind == "smooth":
            assert dc_only.mean() > 0.05
    # statistics go with it
    from vcf_b200.codec import stats_dict
    img = O.synthetic_frame(64, 512, 4100, "natural")
    idx = O.encode_array(img, 16, 16, color=color)
    got, st = _codec(block_size=16, q=16, color=color).decode(t.from_numpy(idx).cuda(), (64, 512),
                                                              original=t.from_numpy(img).cuda(), stats=True)
    assert _lib.last_kernel() == "dec16_f32"
    s = stats_dict(st.cpu().numpy())
    assert int(s["sse"].sum()) == O.sse_int(img, got.cpu().numpy()) and s["nsamples"] == img.size


@pytest.mark.parametrize("q", [1, 8, 12, 32, 64, 255])
def test_fast_decode_path(q, torch_cuda):
    """TMA decode fast path: float64 mode bit-exact with the reference chain, float32
    mode within +-1 LSB; partial tiles, vertical padding (cropped rows), batches, and
    arbitrary (non-encoder) index arrays."""
    from vcf_b200 import _lib
    t = torch_cuda
    rng = np.random.default_rng(q)
    for si, (H, W) in enumerate(FAST_SHAPES):
        n = 3 if H * W < 200000 else 1
        frames = np.stack([O.synthetic_frame(H, W, 950 + 10 * si + i, "noise" if i % 2 else "natural") for i in range(n)])
        idx = np.stack([O.encode_array(f, 8, q) for f in frames])
        if si % 2:
            idx[-1] = rng.integers(0, 256, size=idx[-1].shape, dtype=np.uint8)
        ref = np.stack([O.decode_array(k, (H, W, 3), 8, q) for k in idx])
        got = _codec(block_size=8, q=q, fp64=True).decode(t.from_numpy(idx).cuda(), (H, W))
        want = "dec8_fast" if H % 8 == 0 else "decode_general"     # vertical padding -> general kernel
        assert _lib.last_kernel() == want, (H, W, _lib.last_kernel())
        assert np.array_equal(got.cpu().numpy(), ref), (H, W, q, int((got.cpu().numpy() != ref).sum()))
        for contract in (False, True):
            g32 = _codec(block_size=8, q=q, contract=contract).decode(t.from_numpy(idx).cuda(), (H, W))
            assert _lib.last_kernel() == ("dec8_tc" if want == "dec8_fast" else want)     # float32 = the tensor-core tier
            assert np.abs(g32.cpu().numpy().astype(np.int16) - ref.astype(np.int16)).max() <= 1
    # q too large for the fast path -> general kernel (int16 wrap semantics), still exact
    img = O.synthetic_frame(16, 128, 7, "natural")
    idx = O.encode_array(img, 8, 300)
    got = _codec(block_size=8, q=300, fp64=True).decode(t.from_numpy(idx).cuda(), (16, 128))
    assert _lib.last_kernel() == "decode_general"
    assert np.array_equal(got.cpu().numpy(), O.decode_array(idx, img.shape, 8, 300))


def _crafted_indices(rng, H, W, kind):
    """Index arrays (subband layout, B=8) that steer the two-tier decoder into each of its
    branches: DC-only half-tiles, DC-only blocks next to dense ones, sparse +-1 indices whose
    colour mix cancels (exact integers in R, G or B although Y/Co/Cg carry AC terms)."""
    ny, nx = H // 8, W // 8
    k = np.zeros((ny, nx, 8, 8, 3), dtype=np.int64)            # [block y, block x, u, i, c]
    k[:, :, 0, 0, :] = rng.integers(-40, 41, size=(ny, nx, 3))
    if kind == "dc":
        pass
    elif kind == "dc_extreme":
        k[:, :, 0, 0, :] = rng.choice([-128, -127, 0, 126, 127], size=(ny, nx, 3))
    elif kind == "mixed":
        dense = rng.random((ny, nx)) < 0.15
        k[dense] = rng.integers(-6, 7, size=(int(dense.sum()), 8, 8, 3))
        sparse = (rng.random((ny, nx)) < 0.5) & ~dense
        n = int(sparse.sum())
        ks = np.zeros((n, 8, 8, 3), dtype=np.int64)
        ks[:, 0, 0, :] = rng.integers(-40, 41, size=(n, 3))
        u, i = rng.integers(0, 3, size=n), rng.integers(0, 3, size=n)
        pat = rng.integers(0, 4, size=n)
        for j in range(n):
            v = int(rng.integers(1, 3))
            if pat[j] == 0:      # Y and Cg equal: R and B planes lose the term, G keeps it
                ks[j, u[j], i[j], 0] = v; ks[j, u[j], i[j], 2] = v
            elif pat[j] == 1:    # Co only: G = Y + Cg stays DC-only
                ks[j, u[j], i[j], 1] = v
            elif pat[j] == 2:    # Y = -Cg: G plane DC-only
                ks[j, u[j], i[j], 0] = v; ks[j, u[j], i[j], 2] = -v
            else:                # rational positions only
                ks[j, 4, 0, 0] = v; ks[j, 0, 4, 1] = -v; ks[j, 4, 4, 2] = v
        k[sparse] = ks
    elif kind in ("low2", "low4", "low24", "low42"):
        # non-zero indices only in the first 2 / 4 coefficient rows and columns: the pruned
        # codelets (dct8_inv_low2 / _low4) of the exact decoder, every row/column combination
        nu, ni = {"low2": (2, 2), "low4": (4, 4), "low24": (2, 4), "low42": (4, 2)}[kind]
        k[:, :, :nu, :ni, :] = rng.integers(-9, 10, size=(ny, nx, nu, ni, 3))
        k[:, : nx // 2, :, :, 1] = 0                             # a channel without AC in half of the tiles
        k[:, : nx // 2, 0, 0, 1] = rng.integers(-40, 41, size=(ny, nx // 2))
    elif kind == "lowmix":
        # the extent of the non-zero indices changes from one group of 8 blocks to the next, so
        # consecutive half-tiles of a warp take different codelet variants (stale intermediates of a
        # previous half-tile must never be read)
        for by in range(ny):
            for g0 in range(0, nx, 8):
                nu, ni = int(rng.choice([1, 2, 3, 4, 6, 8])), int(rng.choice([1, 2, 3, 4, 6, 8]))
                k[by, g0:g0 + 8, :nu, :ni, :] = rng.integers(-9, 10, size=(min(8, nx - g0), nu, ni, 3))
    elif kind == "rowwise":
        # long DC-only runs interrupted by single dense blocks: partially DC-only half-tiles
        hit = rng.random((ny, nx)) < 0.04
        k[hit] = rng.integers(-3, 4, size=(int(hit.sum()), 8, 8, 3))
    sub = k.transpose(2, 0, 3, 1, 4).reshape(H, W, 3)           # sub[j*ny+y, i*nx+x, c]
    return (sub + 128).astype(np.uint8)


@pytest.mark.parametrize("cfg", ["8x1", "9x1", "4x3"], ids=["two_tier", "exact_dcskip", "two_tier_12warps"])
@pytest.mark.parametrize("kind", ["dc", "dc_extreme", "mixed", "rowwise", "low2", "low4", "low24", "low42", "lowmix"])
def test_two_tier_decode_branches(kind, cfg, torch_cuda, monkeypatch):
    """Every branch of kernels_dec2t.cu (and the DC-only shortcut of the exact kernel) against the
    oracle's float64 chain, bit for bit.  The decoder is forced: at these sizes the library would
    not probe and would use the plain exact kernel."""
    from vcf_b200 import _lib
    t = torch_cuda
    monkeypatch.setenv("VCFB_DEC_CFG", cfg)
    rng = np.random.default_rng(len(kind))
    for (H, W) in ((64, 256), (120, 640)):
        for q in (1, 3, 8, 12, 16, 31, 32, 64, 255):
            idx = _crafted_indices(rng, H, W, kind)
            ref = O.decode_array(idx, (H, W, 3), 8, q)
            got = _codec(block_size=8, q=q, fp64=True).decode(t.from_numpy(idx).cuda(), (H, W))
            assert _lib.last_kernel() == "dec8_fast"
            bad = int((got.cpu().numpy() != ref).sum())
            assert bad == 0, (kind, H, W, q, bad)


def test_float64_decoders_agree_full_size(torch_cuda, monkeypatch):
    """4K frames, natural and noise content, q from dense to DC-only indices: the probed default
    (device-side choice), the two-tier decoder and the exact kernel with the DC-only shortcut
    against the plain exact kernel (itself checked against the oracle), whole batch."""
    t = torch_cuda
    H, W = 2160, 3840
    frames = np.stack([O.synthetic_frame(H, W, 40 + i, "noise" if i == 2 else "natural") for i in range(3)])
    for x in (t.from_numpy(frames).cuda(), t.from_numpy(frames[2:]).cuda(), t.from_numpy(frames[:1]).cuda()):
        for q in (2, 5, 8, 12, 16, 24, 32, 48, 64):
            idx = _codec(block_size=8, q=q).encode(x)
            monkeypatch.setenv("VCFB_DEC_CFG", "9x2")
            ref = _codec(block_size=8, q=q, fp64=True).decode(idx, (H, W))
            for cfg in (None, "8x1", "9x1"):
                if cfg:
                    monkeypatch.setenv("VCFB_DEC_CFG", cfg)
                else:
                    monkeypatch.delenv("VCFB_DEC_CFG")
                got = _codec(block_size=8, q=q, fp64=True).decode(idx, (H, W))
                assert bool(t.equal(got, ref)), (q, cfg, int((got != ref).sum().item()))
            if q == 16 and x.shape[0] == 3:
                refd = O.decode_array(idx[0].cpu().numpy(), (H, W, 3), 8, q)
                assert np.array_equal(ref[0].cpu().numpy(), refd)
        monkeypatch.delenv("VCFB_DEC_CFG", raising=False)


@pytest.mark.parametrize("tier", ["tensor_core", "cuda_core"])
def test_fast_mode_float32_decoder_tolerances(tier, torch_cuda, monkeypatch):
    """BASELINE north star, fast mode: decoded pixels within +-1 LSB of the reference and PSNR
    matching to 0.01 dB -- natural, noise and smooth content, dense to DC-only indices, plus
    arbitrary (non-encoder) index arrays.  Both float32 decoders: the tensor-core tier
    (kernels_tc.cu: 64 x 64 Kronecker inverse DCT on tcgen05, the default) and the CUDA-core kernel
    (kernels_dec32.cu: scaled AAN transform); in both, blocks without AC indices go through the
    reference's float64 chain."""
    from vcf_b200 import _lib
    if tier == "cuda_core":
        monkeypatch.setenv("VCFB_DEC32_CFG", "4x3")
    expect_kernel = "dec8_tc" if tier == "tensor_core" else "dec8_fast"
    t = torch_cuda
    H, W = 1080, 1920
    yy, xx = np.mgrid[0:H, 0:W]
    smooth = np.stack([(128 + 100 * np.sin(xx / 300.0 + c) * np.cos(yy / 200.0)).astype(np.uint8) for c in range(3)], -1)
    frames = {"natural": O.synthetic_frame(H, W, 70, "natural"), "noise": O.synthetic_frame(H, W, 71, "noise"),
              "smooth": np.ascontiguousarray(smooth)}
    for name, img in frames.items():
        for q in (1, 4, 8, 12, 16, 24, 32, 64, 128):
            idx = O.encode_array(img, 8, q)
            ref = O.decode_array(idx, img.shape, 8, q)
            got = _codec(block_size=8, q=q).decode(t.from_numpy(idx).cuda(), (H, W)).cpu().numpy()
            assert _lib.last_kernel() == expect_kernel
            d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
            assert d.max() <= 1, (name, q, int(d.max()))
            p_ref, p_got = O.psnr(img, ref), O.psnr(img, got)
            assert abs(p_ref - p_got) < 0.01, (name, q, p_ref, p_got, float((d != 0).mean()))
    rng = np.random.default_rng(9)
    for q in (3, 32, 255):
        idx = rng.integers(0, 256, size=(64, 256, 3), dtype=np.uint8)
        ref = O.decode_array(idx, (64, 256, 3), 8, q)
        got = _codec(block_size=8, q=q).decode(t.from_numpy(idx).cuda(), (64, 256)).cpu().numpy()
        assert np.abs(got.astype(np.int16) - ref.astype(np.int16)).max() <= 1, q
    for kind in ("dc", "dc_extreme", "mixed", "rowwise"):            # exact where it must be: DC-only blocks
        idx = _crafted_indices(rng, 64, 256, kind)
        for q in (8, 12, 31, 64):
            ref = O.decode_array(idx, (64, 256, 3), 8, q)
            got = _codec(block_size=8, q=q).decode(t.from_numpy(idx).cuda(), (64, 256)).cpu().numpy()
            assert np.abs(got.astype(np.int16) - ref.astype(np.int16)).max() <= 1, (kind, q)
            if kind.startswith("dc"):
                assert np.array_equal(got, ref), (kind, q)


@pytest.mark.parametrize("color", ["YCoCg", "YCrCb"])
def test_standalone_colour_codecs(color, torch_cuda, golden_dir):
    """src/YCoCg.py:33-85 and src/YCrCb.py:33-69 as codecs of their own (SURVEY 8a rows A11,
    A12): bit-exact against the vectors recorded from the unmodified reference scripts and
    against the oracle on seeded inputs, incl. odd sizes and unaligned device pointers."""
    from vcf_b200 import ColorCodec, _lib
    t = torch_cuda
    g = np.load(os.path.join(golden_dir, "ref_flow_sa_%s.npz" % color.lower()))
    q = int([str(x) for x in g["flags"]][1])
    cc = ColorCodec(color, q)
    assert np.array_equal(cc.encode(g["img"]), g["idx"])
    assert np.array_equal(cc.decode(g["idx"]), g["decoded"])
    enc_o = O.ycocg_standalone_encode if color == "YCoCg" else O.ycrcb_standalone_encode
    dec_o = O.ycocg_standalone_decode if color == "YCoCg" else O.ycrcb_standalone_decode
    rng = np.random.default_rng(3)
    for q in (1, 3, 8, 32, 200):
        cc = ColorCodec(color, q)
        for shape in ((1, 1, 3), (7, 5, 3), (64, 96, 3), (3, 217, 131, 3)):
            img = rng.integers(0, 256, size=shape, dtype=np.uint8)
            k = cc.encode(img)
            assert k.dtype == np.uint16 and np.array_equal(k, enc_o(img.reshape(-1, 1, 3), q).reshape(shape))
            assert _lib.last_kernel() == "color_encode"
            assert np.array_equal(cc.decode(k), dec_o(k.reshape(-1, 1, 3), q).reshape(shape))
            kr = rng.integers(0, 65536, size=shape, dtype=np.uint16)      # arbitrary indices: wrap paths
            assert np.array_equal(cc.decode(kr), dec_o(kr.reshape(-1, 1, 3), q).reshape(shape))
        # torch path, unaligned pointer
        img = rng.integers(0, 256, size=(50, 70, 3), dtype=np.uint8)
        buf = t.zeros(img.size + 16, dtype=t.uint8, device="cuda")
        v = buf[1:1 + img.size].view(50, 70, 3)
        v.copy_(t.from_numpy(img))
        kt = cc.encode(v)
        assert np.array_equal(kt.cpu().numpy(), enc_o(img, q))
        assert np.array_equal(cc.decode(kt).cpu().numpy(), dec_o(enc_o(img, q), q))


def test_config3_rde_sweep_full_size(torch_cuda):
    """BASELINE config 3: 4K frame, B in {4,8,16,32} x 8 q values.  Exact index parity on
    every point would take minutes of oracle time, so: exact parity for one q per B
    (incl. a non-power-of-two), and for all 32 points the RD statistics of the GPU sweep
    are checked against properties (RMSE monotone in q, rate estimate falling with q) and
    against the oracle's RMSE on the points that were decoded."""
    from vcf_b200.rd import rd_sweep
    t = torch_cuda
    H, W = 2160, 3840
    img = O.synthetic_frame(H, W, 3, "natural")
    x = t.from_numpy(img).cuda()
    qs = (4, 8, 12, 16, 24, 32, 48, 64)
    pts = rd_sweep(x, (4, 8, 16, 32), qs)
    assert len(pts) == 32
    for B in (4, 8, 16, 32):
        row = [p for p in pts if p["B"] == B]
        # over the power-of-two steps (nested dead zones) distortion grows and rate falls --
        # for q >= B only: below that the DC index leaves [-128,127] and the reference's
        # uint8 cast wraps (src/2D-DCT.py:361), e.g. RMSE 25.0 at B=8, q=4 vs 5.5 at q=8
        p2 = [p for p in row if p["q"] in (4, 8, 16, 32, 64) and p["q"] >= B]
        rm = [p["rmse"] for p in p2]
        assert all(rm[i] < rm[i + 1] for i in range(len(rm) - 1)), (B, rm)
        br = [p["bpp_entropy"] for p in p2]
        assert all(br[i] > br[i + 1] for i in range(len(br) - 1)), (B, br)
        assert all(0 < p["bpp_entropy"] < 24 and p["rmse"] > 0 for p in row)
    for B, q in ((4, 12), (16, 24), (32, 48)):
        ref = O.encode_array(img, B, q)
        got = _codec(block_size=B, q=q).encode(x).cpu().numpy()
        assert np.array_equal(got, ref), (B, q)
        refd = O.decode_array(ref, img.shape, B, q)
        dec = _codec(block_size=B, q=q, fp64=True).decode(t.from_numpy(ref).cuda(), (H, W)).cpu().numpy()
        assert np.array_equal(dec, refd), (B, q)
        p = [p for p in pts if p["B"] == B and p["q"] == q][0]
        assert p["sse"] == O.sse_int(img, refd)
        assert abs(p["rmse"] - float(O.rmse(img, refd))) < 1e-3


def test_config5_8k_ycrcb_b16_full_size(torch_cuda):
    """BASELINE config 5 shape: one 7680x4320 frame, YCrCb (float extension) + B=16, q=32:
    exact parity with the oracle (float32 encode, float64 decode) and statistics."""
    t = torch_cuda
    H, W = 4320, 7680
    img = O.synthetic_frame(H, W, 2000, "natural")
    x = t.from_numpy(img).cuda()
    enc = _codec(block_size=16, q=32, color="YCrCb")
    got, st = enc.encode(x, stats=True)
    ref = O.encode_array(img, 16, 32, color="YCrCb")
    assert np.array_equal(got.cpu().numpy(), ref)
    dec = _codec(block_size=16, q=32, color="YCrCb", fp64=True)
    y, sd = dec.decode(got, (H, W), original=x, stats=True)
    refd = O.decode_array(ref, img.shape, 16, 32, color="YCrCb")
    assert np.array_equal(y.cpu().numpy(), refd)
    from vcf_b200.codec import stats_dict
    s = stats_dict((st + sd).cpu().numpy())
    assert int(s["sse"].sum()) == O.sse_int(img, refd) and s["nsamples"] == img.size
    nz, sabs, hist = O.index_stats(ref)
    assert s["nonzero"] == nz and np.array_equal(s["hist"], hist)


def test_config4_1080p_sequence_batch(torch_cuda):
    """BASELINE config 4 shape: a batch of 1920x1080 frames (B=8, q=32; 1080 = 135*8).
    Frame i of a batch must equal frame i coded alone (frames are independent,
    src/III.py:132-144), and a few frames are checked exactly against the oracle."""
    t = torch_cuda
    n, H, W = 24, 1080, 1920
    frames = np.stack([O.synthetic_frame(H, W, 1000 + i, "natural") for i in range(4)])
    frames = np.concatenate([frames] * (n // 4))
    frames[5] = O.synthetic_frame(H, W, 5, "noise")
    x = t.from_numpy(frames).cuda()
    enc, dec = _codec(block_size=8, q=32), _codec(block_size=8, q=32, fp64=True)
    idx = enc.encode(x)
    y = dec.decode(idx, (H, W))
    for i in (0, 5, 23):
        assert t.equal(idx[i], enc.encode(x[i]))
        assert t.equal(y[i], dec.decode(idx[i], (H, W)))
    for i in (1, 5):
        ref = O.encode_array(frames[i], 8, 32)
        assert np.array_equal(idx[i].cpu().numpy(), ref)
        assert np.array_equal(y[i].cpu().numpy(), O.decode_array(ref, frames[i].shape, 8, 32))
    assert t.equal(idx[0], idx[4]) and t.equal(y[1], y[9])          # identical frames, identical streams


def test_fast_path_statistics(torch_cuda):
    """Statistics requested on a fast-path shape: the TMA kernels run and accumulate the
    statistics themselves (the histogram and the float32 decoder take the streaming statistics
    kernels); every number must equal the oracle's."""
    from vcf_b200 import _lib
    from vcf_b200.codec import stats_dict
    t = torch_cuda
    # (2160x3840 has 8100 tiles: the probed float64 decode; q = 2 -> two-tier, 16 -> exact chain,
    #  48 -> exact chain with the DC-only shortcut; all with the distortion statistics fused in)
    for (n, H, W, q) in ((3, 64, 256, 16), (1, 1080, 1920, 8), (1, 2160, 3840, 2), (1, 2160, 3840, 16), (1, 2160, 3840, 48)):
        frames = np.stack([O.synthetic_frame(H, W, 1200 + i, "natural" if i % 2 == 0 else "noise") for i in range(n)])
        x = t.from_numpy(frames).cuda()
        ref = np.stack([O.encode_array(f, 8, q) for f in frames])
        refd = np.stack([O.decode_array(k, (H, W, 3), 8, q) for k in ref])
        nz, sabs, hist = O.index_stats(ref)
        for use_hist in (True, False):
            enc = _codec(block_size=8, q=q, hist=use_hist)
            idx, st = enc.encode(x, stats=True)
            assert _lib.last_kernel() == "enc8_fast"
            assert np.array_equal(idx.cpu().numpy(), ref)
            s = stats_dict(st.cpu().numpy())
            assert s["nonzero"] == nz and s["sumabs"] == sabs and s["nindices"] == ref.size
            assert np.array_equal(s["hist"], hist if use_hist else np.zeros_like(hist))
        dec = _codec(block_size=8, q=q, fp64=True)
        y, sd = dec.decode(idx, (H, W), original=x, stats=True)
        assert _lib.last_kernel() == "dec8_fast"
        assert np.array_equal(y.cpu().numpy(), refd)
        s = stats_dict(sd.cpu().numpy())
        for c in range(3):
            assert int(s["sse"][c]) == O.sse_int(frames[..., c], refd[..., c])
        assert s["nsamples"] == frames.size
        assert s["sumdiff"] == int((frames.astype(np.int64) - refd.astype(np.int64)).sum())
        # float32 decoder: separate streaming pass over (original, decoded)
        y32, s32 = _codec(block_size=8, q=q).decode(idx, (H, W), original=x, stats=True)
        s32 = stats_dict(s32.cpu().numpy())
        y32 = y32.cpu().numpy()
        for c in range(3):
            assert int(s32["sse"][c]) == O.sse_int(frames[..., c], y32[..., c])
        assert s32["nsamples"] == frames.size


def test_no_writes_outside_the_output_arrays(torch_cuda):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are hunted with
    canaries: every output lives inside a 0xAB-filled buffer at an odd offset and the bytes
    around it must be untouched after each kernel family (general at every B, TMA fast
    path, colour codecs); results are checked against the oracle as well."""
    from vcf_b200 import ColorCodec
    t = torch_cuda
    PAD = 4096

    def guarded(nbytes, off):
        buf = t.full((PAD + off + nbytes + PAD,), 0xAB, dtype=t.uint8, device="cuda")
        return buf, buf[PAD + off:PAD + off + nbytes]

    def intact(buf, nbytes, off):
        return bool((buf[:PAD + off] == 0xAB).all()) and bool((buf[PAD + off + nbytes:] == 0xAB).all())

    cases = [(8, (2, 16, 128), 0), (8, (1, 40, 1024), 0),              # fast path (aligned by construction)
             (8, (1, 37, 53), 1), (4, (2, 30, 70), 3), (16, (1, 50, 90), 5), (32, (1, 70, 130), 7),
             (8, (1, 8, 8), 2), (16, (1, 1, 1), 9)]
    for B, (n, H, W), off in cases:
        frames = np.stack([O.synthetic_frame(H, W, 1300 + i, "noise") for i in range(n)])
        Hp, Wp, _, _ = O.padded_shape(H, W, B)
        xbuf, xv = guarded(frames.size, off)
        x = xv.view(n, H, W, 3)
        x.copy_(t.from_numpy(frames))
        ibuf, iv = guarded(n * Hp * Wp * 3, off)
        idx = iv.view(n, Hp, Wp, 3)
        _codec(block_size=B, q=8).encode(x, out=idx)
        t.cuda.synchronize()
        assert intact(ibuf, n * Hp * Wp * 3, off), ("encode", B, H, W, off)
        ref = np.stack([O.encode_array(f, B, 8) for f in frames])
        assert np.array_equal(idx.cpu().numpy(), ref)
        for fp64 in (True, False):
            ybuf, yv = guarded(frames.size, off)
            y = yv.view(n, H, W, 3)
            _codec(block_size=B, q=8, fp64=fp64).decode(idx, (H, W), out=y)
            t.cuda.synchronize()
            assert intact(ybuf, frames.size, off), ("decode", B, H, W, off, fp64)
            if fp64:
                assert np.array_equal(y.cpu().numpy(), np.stack([O.decode_array(k, (H, W, 3), B, 8) for k in ref]))
        assert intact(xbuf, frames.size, off)
    # colour codecs write through plain pointers as well
    img = t.from_numpy(O.synthetic_frame(33, 71, 5, "noise")).cuda()
    for color in ("YCoCg", "YCrCb"):
        k = ColorCodec(color, 7).encode(img)
        assert k.shape == img.shape


@pytest.mark.parametrize("q", [0.5, 2.5, 12.5, 1000, 40000])
def test_unusual_quantisation_steps(q, torch_cuda):
    """Steps the reference's CLI would not produce but its arithmetic defines: fractional
    (numpy divides by a float, decodes in float64 without the int16 product), sub-unit powers
    of two (massive uint8 wrap), and steps whose int16 product q*k wraps (src/2D-DCT.py:410
    keeps int16)."""
    t = torch_cuda
    for B, (H, W) in ((8, (64, 256)), (8, (40, 72)), (16, (48, 80))):
        img = O.synthetic_frame(H, W, 1400 + B, "natural")
        x = t.from_numpy(img).cuda()
        for fp64, dt in ((False, np.float32), (True, np.float64)):
            ref = O.encode_array(img, B, q, dtype=dt)
            got = _codec(block_size=B, q=q, fp64=fp64).encode(x)
            assert np.array_equal(got.cpu().numpy(), ref), (B, H, W, q, fp64)
        idx = np.random.default_rng(int(q * 2)).integers(0, 256, size=ref.shape, dtype=np.uint8)
        if q >= 32768:      # the reference's int16 dequantiser raises; so does the library
            from vcf_b200 import VcfbError
            with pytest.raises(OverflowError):
                O.decode_array(idx, img.shape, B, q)
            with pytest.raises(VcfbError):
                _codec(block_size=B, q=q, fp64=True).decode(t.from_numpy(idx).cuda(), (H, W))
            continue
        with np.errstate(over="ignore", invalid="ignore"):
            refd = O.decode_array(idx, img.shape, B, q)
        got = _codec(block_size=B, q=q, fp64=True).decode(t.from_numpy(idx).cuda(), (H, W))
        assert np.array_equal(got.cpu().numpy(), refd), (B, H, W, q)
