"""The upstream-variant decoder (VCFB_F_SYNTH_F32) against the oracle evaluated the same way, and
full-size uniform-noise frames (the tie-dense case, 1 822 exact DC ties per 4K frame
at q = 8) through every decoder of the B=8 fast path.  Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vcf_oracle as O


def _codec(**kw):
    from vcf_b200 import Codec
    return Codec(**kw)


@pytest.mark.parametrize("B", [4, 8, 16, 32])
def test_synth_f32_variant_bit_exact(B):
    """float64 IDCT, result stored as float32, to_RGB / +128 / truncation in float32."""
    from vcf_b200 import _lib, VcfbError
    for (H, W), kind in (((136, 200), "natural"), ((64, 256), "noise"), ((270, 480), "natural")):
        img = O.synthetic_frame(H, W, 11 + B, kind)
        for q in (8, 32, 12):
            idx = O.encode_array(img, B, q)
            ref = O.decode_array(idx, img.shape, B, q, synth_store_dtype=np.float32)
            got = _codec(block_size=B, q=q, fp64=True, synth_f32=True).decode(idx, (H, W))
            assert _lib.last_kernel() == "decode_general"
            assert np.array_equal(got, ref), (B, q, H, W, int((got != ref).sum()))
            base = O.decode_array(idx, img.shape, B, q)
            assert np.abs(ref.astype(np.int16) - base.astype(np.int16)).max() <= 1
    with pytest.raises(ValueError):
        _codec(block_size=8, q=8, synth_f32=True)
    with pytest.raises(VcfbError):           # the C ABI refuses the flag without VCFB_F_FP64
        c = _codec(block_size=B, q=12, fp64=True, synth_f32=True)
        c.flags &= ~4
        c.decode(idx, (H, W))


@pytest.mark.parametrize("q", [8, 16, 32, 64])
def test_full_size_4k_noise_every_decoder(q, monkeypatch):
    """3840x2160 uniform noise: the packed encoder and each float64 decoder of the fast path (probed
    default, two-tier, exact + shortcuts, plain exact chain) bit for bit against the oracle; the
    float32 decoder within the north star's tolerance."""
    import torch
    from vcf_b200 import _lib
    H, W = 2160, 3840
    img = O.synthetic_frame(H, W, 2, "noise")
    x = torch.from_numpy(img).cuda()
    ref = O.encode_array(img, 8, q)
    got = _codec(block_size=8, q=q).encode(x)
    assert _lib.last_kernel() == "enc8_fast"
    assert np.array_equal(got.cpu().numpy(), ref), int((got.cpu().numpy() != ref).sum())
    refd = O.decode_array(ref, img.shape, 8, q)
    for cfg in (None, "8x1", "9x1", "9x2"):
        if cfg:
            monkeypatch.setenv("VCFB_DEC_CFG", cfg)
        else:
            monkeypatch.delenv("VCFB_DEC_CFG", raising=False)
        dec = _codec(block_size=8, q=q, fp64=True).decode(got, (H, W))
        assert _lib.last_kernel() == "dec8_fast"
        bad = int((dec.cpu().numpy() != refd).sum())
        assert bad == 0, (q, cfg, bad)
    monkeypatch.delenv("VCFB_DEC_CFG", raising=False)
    for cfg32, kern in ((None, "dec8_tc"), ("4x3", "dec8_fast")):          # tensor-core tier, CUDA-core kernel
        if cfg32:
            monkeypatch.setenv("VCFB_DEC32_CFG", cfg32)
        d32 = _codec(block_size=8, q=q).decode(got, (H, W)).cpu().numpy()
        assert _lib.last_kernel() == kern
        assert np.abs(d32.astype(np.int16) - refd.astype(np.int16)).max() <= 1
        assert abs(O.psnr(img, d32) - O.psnr(img, refd)) < 0.01
    monkeypatch.delenv("VCFB_DEC32_CFG", raising=False)


def test_full_size_8k_noise_b16():
    """7680x4320 uniform noise, B=16 (config 5's shape): fast-path encoder and float64 decoder."""
    import torch
    from vcf_b200 import _lib
    H, W = 4320, 7680
    img = O.synthetic_frame(H, W, 5, "noise")
    x = torch.from_numpy(img).cuda()
    for q, color in ((32, "YCoCg"), (8, "YCrCb")):
        ref = O.encode_array(img, 16, q, color=color)
        got = _codec(block_size=16, q=q, color=color).encode(x)
        assert _lib.last_kernel() == "enc16_fast"
        assert np.array_equal(got.cpu().numpy(), ref), (q, color)
        refd = O.decode_array(ref, img.shape, 16, q, color=color)
        dec = _codec(block_size=16, q=q, color=color, fp64=True).decode(got, (H, W))
        assert _lib.last_kernel() == "dec16_fast"
        assert np.array_equal(dec.cpu().numpy(), refd), (q, color)
