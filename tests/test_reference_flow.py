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
from _util import parse_flags as _parse

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "ref_flow_*.npz")))


def test_golden_present():
    assert len(FILES) >= 12


@pytest.mark.parametrize("fn", FILES, ids=[os.path.basename(f)[9:-4] for f in FILES])
def test_oracle_matches_reference_flow(fn):
    g = np.load(fn)
    img, idx, dec = g["img"], g["idx"], g["decoded"]
    script = str(g["script"])
    kw = _parse(g["flags"])
    if script == "2D-DCT.py":
        assert tuple(g["shape_bin"]) == img.shape
        for loop in (False, True):
            k = O.encode_array(img, loop=loop, **kw)
            assert k.dtype == np.uint8 and k.shape == idx.shape
            assert np.array_equal(k, idx)
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
