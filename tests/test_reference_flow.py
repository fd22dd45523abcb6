"""Oracle vs. vectors produced by the UNMODIFIED reference scripts
(oracle/make_golden.py; /root/reference/src/2D-DCT.py, YCoCg.py, YCrCb.py,
RDE.py run under the shadow packages).  Pins the oracle's control flow --
padding, -128 shift, subband permutation, bias + uint8 wrap, int16/float64
decode chain, crop, clip + truncation -- to src/2D-DCT.py:268-372 / :377-468."""
import glob
import os

import numpy as np
import pytest

from oracle import vcf_oracle as O
from _util import golden_filter, golden_kw, parse_flags as _parse

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "ref_flow_*.npz")))


def test_golden_present():
    assert len(FILES) >= 12


@pytest.mark.parametrize("fn", FILES, ids=[os.path.basename(f)[9:-4] for f in FILES])
def test_oracle_matches_reference_flow(fn):
    g = np.load(fn)
    img, idx, dec = g["img"], g["idx"], g["decoded"]
    script = str(g["script"])
    kw = golden_kw(g)
    if script == "2D-DCT.py":
        assert tuple(g["shape_bin"]) == img.shape
        for loop in (False, True):
            k = O.encode_array(img, loop=loop, **kw)
            assert k.dtype == np.uint8 and k.shape == idx.shape
            assert np.array_equal(k, idx)
            if golden_filter(g) == "gaussian_blur":
                import cv2   # src/gaussian_blur.py: the filter gets the un-clipped float64 image (:461), the clip follows (:466)
                yf = O.decode_array(idx, img.shape, loop=loop, return_float=True, **kw)
                assert yf.dtype == np.float64
                y = np.clip(cv2.GaussianBlur(yf, (5, 5), 0), 0, 255).astype(np.uint8)
            else:
                assert golden_filter(g) is None
                y = O.decode_array(idx, img.shape, loop=loop, **kw)
            assert np.array_equal(y, dec)
    elif script == "YCoCg.py":
        assert np.array_equal(O.ycocg_standalone_encode(img, kw["q"]), idx)
        assert np.array_equal(O.ycocg_standalone_decode(idx, kw["q"]), dec)
    elif script == "YCrCb.py":
        assert np.array_equal(O.ycrcb_standalone_encode(img, kw["q"]), idx)
        assert np.array_equal(O.ycrcb_standalone_decode(idx, kw["q"]), dec)
    else:
        raise AssertionError(script)


def test_rde_rmse_matches_reference_report():
    g = np.load(os.path.join(GOLD, "ref_flow_default_96x80.npz"))
    r = float(O.rmse(g["img"], g["decoded"]))
    assert abs(round(r, 2) - float(g["rde_rmse_2dp"])) < 1e-9
    bpp = int(g["rde_codestream_bytes"]) * 8 / (g["img"].shape[0] * g["img"].shape[1])
    assert abs(round(bpp + r, 2) - float(g["rde_J_2dp"])) <= 0.011
    s = O.sse_int(g["img"], g["decoded"])
    assert abs(np.sqrt(s / g["img"].size) - r) < 1e-4


L_FILES = [f for f in FILES if "ref_flow_L_" in f]


@pytest.mark.parametrize("fn", L_FILES, ids=[os.path.basename(f)[9:-4] for f in L_FILES])
def test_block_size_search_matches_the_references_log(fn):
    """``-L``: the J = rate + Lambda * RMSE the unmodified reference logs for every block size (src/2D-DCT.py:575-576)
    against the oracle's restatement of the loop body -- pins that the loop runs WITHOUT the 128 offset (it is called
    before ``self.offset = 128`` is assigned, :99-110), dequantises the quantiser's own indices, and hands
    ``astype(uint8)`` of the un-biased indices to the entropy coder."""
    import io
    g = np.load(fn)
    img, q, lam = g["img"], _parse(g["flags"])["q"], float(g["L_lambda"])
    assert list(g["L_block_sizes"]) == [2, 4, 8, 16, 32, 64, 128]
    best, bestJ = None, 1000000
    for B, J_ref in zip(g["L_block_sizes"], g["L_J"]):
        k_u8, y, rm = O.optimize_block_size_point(img, int(B), q)
        b = io.BytesIO()
        np.savez_compressed(file=b, a=k_u8)                    # src/z_lib.py:19-23
        J = len(b.getvalue()) + lam * rm
        assert abs(float(J) - float(J_ref)) <= 1e-6 * float(J_ref), (int(B), float(J), float(J_ref))
        if J < bestJ:
            best, bestJ = int(B), J
        # with the offset its author presumably intended the numbers differ: the vector does discriminate
        _, _, rm128 = O.optimize_block_size_point(img, int(B), q, offset=128)
        if int(B) >= 8 and "bright" in fn:
            assert abs(float(rm128) - float(rm)) > 1e-6
    assert best == int(g["L_chosen"])
