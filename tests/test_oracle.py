"""Oracle self-consistency: loop form == vectorised form, round trips, YCrCb
fixed point against the real OpenCV, and the properties SURVEY.md section 4 lists."""
import numpy as np
import pytest

from oracle import vcf_oracle as O


@pytest.mark.parametrize("B", [4, 8, 16, 32])
def test_loop_and_vectorised_forms_are_bit_identical(B):
    img = O.synthetic_frame(64, 96, 1, "noise")
    for dt in (np.float32, np.float64):
        x = img.astype(dt) - 128
        a = O.analyze_image_loop(O.ycocg_from_rgb(x), B, B)
        b = O.analyze_image(O.ycocg_from_rgb(x), B, B)
        assert a.dtype == b.dtype == dt and np.array_equal(a, b)
        assert np.array_equal(O.synthesize_image_loop(a, B, B), O.synthesize_image(a, B, B))
    k = (O.encode_array(img, B, 8).astype(np.int16) - 128) * 8
    assert O.synthesize_image_loop(k, B, B).dtype == np.float64   # scipy promotes integers
    assert np.array_equal(O.synthesize_image_loop(k, B, B), O.synthesize_image(k, B, B))


@pytest.mark.parametrize("B", [4, 8, 16, 32])
def test_subband_permutation_round_trip(B):
    rng = np.random.default_rng(B)
    x = rng.integers(0, 1000, size=(B * 5, B * 7, 3))
    s = O.get_subbands(x, B, B)
    assert np.array_equal(O.get_blocks(s, B, B), x)
    ny, nx = 5, 7
    for (y, xx, j, i) in [(0, 0, 0, 0), (2, 3, 1, B - 1), (4, 6, B - 1, 2)]:
        assert np.array_equal(s[j * ny + y, i * nx + xx], x[y * B + j, xx * B + i])


def test_ycocg_exact_on_8bit_input():
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, size=(64, 64, 3)).astype(np.float32) - 128
    for dt in (np.float32, np.float64):
        y = O.ycocg_from_rgb(rgb.astype(dt))
        assert np.array_equal(O.ycocg_to_rgb(y), rgb.astype(dt))
        assert np.array_equal(y * 4, np.round(y * 4))          # multiples of 1/4


def test_ycrcb_fixed_point_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, size=(512, 512, 3), dtype=np.uint8)
    # all grey levels and the extremes as well
    rgb[0, :256] = np.arange(256, dtype=np.uint8)[:, None]
    rgb[1, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255],
                  [255, 255, 0], [0, 255, 255], [255, 0, 255]]
    assert np.array_equal(O.ycrcb_from_rgb_u8(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2YCrCb))
    assert np.array_equal(O.ycrcb_to_rgb_u8(rgb), cv2.cvtColor(rgb, cv2.COLOR_YCrCb2RGB))


def test_ycrcb_float_extension_close_to_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    x = (rng.integers(0, 256, size=(64, 64, 3)).astype(np.float32) - 128)
    ref = cv2.cvtColor(x, cv2.COLOR_RGB2YCrCb)
    ref[..., 1:] -= 0.5                                      # OpenCV's float delta
    assert np.abs(O.ycrcb_from_rgb_float(x) - ref).max() < 1e-3
    back = O.ycrcb_to_rgb_float(O.ycrcb_from_rgb_float(x.astype(np.float64)))
    assert np.abs(back - x).max() < 0.5


def test_padding_geometry():
    assert O.padded_shape(2160, 3840, 32) == (2176, 3840, 8, 0)
    assert O.padded_shape(53, 37, 8) == (56, 40, 1, 1)
    img = np.ones((53, 37, 3), np.float32)
    p = O.pad_and_center(img, 8)
    assert p.shape == (56, 40, 3) and p[0].sum() == 0 and p[-2:].sum() == 0 and p[:, 0].sum() == 0
    assert np.array_equal(O.remove_padding(p, img.shape), img)


def test_deadzone_truncates_toward_zero_and_wraps():
    Q = O.DeadzoneQuantizer(32)
    x = np.array([-95.9, -32.0, -31.9, 0.0, 31.9, 32.0, 95.9], np.float32)
    assert list(Q.encode(x)) == [-2, -1, 0, 0, 0, 1, 2]
    assert Q.decode(np.array([-3, 4], np.int16)).dtype == np.int16
    img = O.synthetic_frame(32, 32, 3, "noise")
    k = O.encode_array(img, 16, 1)                            # |DC| up to 2048 -> wraps mod 256
    full = (O.get_subbands(O.analyze_image(O.ycocg_from_rgb(img.astype(np.float32) - 128), 16, 16), 16, 16) / 1).astype(np.int64) + 128
    assert (full > 255).any() or (full < 0).any()
    assert np.array_equal(k, full.astype(np.uint8))


@pytest.mark.parametrize("B,q", [(8, 32), (16, 8), (4, 12), (32, 64)])
def test_round_trip_quality(B, q):
    img = O.synthetic_frame(96, 128, 5, "natural")
    idx = O.encode_array(img, B, q)
    dec = O.decode_array(idx, img.shape, B, q)
    assert dec.shape == img.shape and dec.dtype == np.uint8
    assert O.psnr(img, dec) > 24.0
    lo = O.decode_array(O.encode_array(img, B, 2 * q), img.shape, B, 2 * q)
    assert O.psnr(img, dec) >= O.psnr(img, lo) - 0.5         # finer step, no worse


def test_stats_helpers():
    img = O.synthetic_frame(40, 40, 6, "natural")
    idx = O.encode_array(img, 8, 32)
    nz, sabs, hist = O.index_stats(idx)
    assert hist.shape == (3, 256) and hist.sum() == idx.size
    assert nz == int((idx != 128).sum())
    assert O.entropy_bits(hist[0]) <= 8 * idx[..., 0].size
    dec = O.decode_array(idx, img.shape, 8, 32)
    assert abs(float(O.rmse(img, dec)) - np.sqrt(O.sse_int(img, dec) / img.size)) < 1e-4


def test_dc_only_block_chain():
    """kernels_dec2t.cu evaluates blocks without AC indices as  (X * c0) * c0  with c0 =
    pocketfft's sqrt(2) constant (every other operand of the DAG is an exact zero).  Pinned
    here against the real scipy.fftpack: all 64 samples of idct2(DC-only block) are bitwise
    equal to that product, scaled by the exact 2^-4."""
    import scipy.fftpack as sf
    c0 = float.fromhex("0x1.6a09e667f3bcdp+0")
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.integers(-32768, 32768, size=4000), np.arange(-300, 301)]).astype(np.float64)
    blocks = np.zeros((xs.size, 8, 8))
    blocks[:, 0, 0] = xs
    got = sf.idct(sf.idct(blocks, norm="ortho", axis=1), norm="ortho", axis=2)
    want = (xs * c0) * c0 * 0.0625
    assert np.array_equal(got, np.broadcast_to(want[:, None, None], got.shape))
