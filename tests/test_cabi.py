"""The C-ABI library loads without a GPU and exports every symbol that
include/vcfb200.h declares.  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vcf_b200", "libvcfb200.so")
HDR = os.path.join(ROOT, "include", "vcfb200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcfb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vcfb200.h but not exported"


def test_version_and_geometry(lib):
    from vcf_b200 import _lib
    assert _lib.lib().vcfb_version() == 150
    assert _lib.padded_dims(2160, 3840, 32) == (2176, 3840, 8, 0)
    assert _lib.padded_dims(53, 37, 8) == (56, 40, 1, 1)
    assert _lib.padded_dims(1080, 1920, 16) == (1088, 1920, 4, 0)


def test_argument_errors_without_touching_the_gpu(lib):
    from vcf_b200 import _lib
    L = _lib.lib()
    assert L.vcfb_encode_dev(None, 1, 16, 16, 8, 32.0, 0, 0, None, None, None, None) == -1
    assert b"NULL" in L.vcfb_last_error()
    buf = (ctypes.c_uint8 * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.vcfb_encode_dev(p, 1, 16, 16, 7, 32.0, 0, 0, None, p, None, None) == -1
    assert b"block size" in L.vcfb_last_error()
    assert L.vcfb_encode_dev(p, 1, 16, 16, 8, 0.0, 0, 0, None, p, None, None) == -1
    assert L.vcfb_encode_dev(p, 1, 16, 16, 8, 32.0, 5, 0, None, p, None, None) == -1
    assert L.vcfb_encode_dev(p, 1, 16, 16, 8, 32.0, 0, 2, None, p, None, None) == -1   # -p without weights
    assert L.vcfb_decode_dev(p, 1, 16, 16, 8, 32.0, 0, 0, None, None, None, None, None, None) == -1


def test_no_cpu_fallback_when_no_device():
    """On a box without a GPU the numpy entry point must fail loudly."""
    import numpy as np
    from vcf_b200 import Codec, VcfbError, _lib
    if _lib.lib().vcfb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(VcfbError):
        Codec().encode(np.zeros((16, 16, 3), np.uint8))


def test_perceptual_weights_match_oracle_tables():
    import numpy as np
    from oracle import vcf_oracle as O
    from vcf_b200 import perceptual_weights
    for B in (4, 8, 16, 32):
        Y, C = O.perceptual_tables(B)
        w = perceptual_weights(B)
        assert np.array_equal(w[0], Y / 121) and np.array_equal(w[1], C / 99)
