"""Host logic of vcf_b200/entropy.py without a GPU: the TIFF and .npz containers around finished
streams (made here with the host's zlib, which is what the GPU stream is interchangeable with),
the CRC-32 combination, and the refusal to compress without a device."""
import io
import zipfile
import zlib

import cv2
import numpy as np
import pytest

from vcf_b200 import entropy as E


def _raw_deflate(b: bytes) -> bytes:
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    return c.compress(b) + c.flush()


def test_crc32_combine_is_zlibs():
    rng = np.random.default_rng(0)
    for la, lb in ((0, 0), (0, 5), (5, 0), (1, 1), (1000, 77777), (3, 1 << 20)):
        a, b = rng.integers(0, 256, la, dtype=np.uint8).tobytes(), rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert E.crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)


def test_npy_header_is_numpys():
    for arr in (np.zeros((3, 4, 5), np.int16), np.zeros((0, 4), np.float32), np.zeros((), np.uint8), np.zeros(7, np.uint8)):
        fh = io.BytesIO()
        np.save(fh, arr)
        assert E._npy_header(arr.dtype, arr.shape) == fh.getvalue()[: len(fh.getvalue()) - arr.nbytes]


def test_npz_container_is_read_by_numpy_and_zipfile():
    rng = np.random.default_rng(1)
    arrays = {"a": rng.integers(0, 256, (40, 50, 3), dtype=np.uint8), "k": rng.integers(-9, 9, (7, 3), dtype=np.int16),
              "empty": np.zeros((0, 4), np.float32)}
    members = [(n, E._npy_header(v.dtype, v.shape), _raw_deflate(v.tobytes()), zlib.crc32(v.tobytes()), v.nbytes)
               for n, v in arrays.items()]
    fh = io.BytesIO()
    E._npz_container(fh, members)
    fh.seek(0)
    assert zipfile.ZipFile(fh).testzip() is None
    fh.seek(0)
    z = np.load(fh)                                   # /root/reference/src/z_lib.py:25-29
    for n, v in arrays.items():
        assert z[n].dtype == v.dtype and z[n].shape == v.shape and np.array_equal(z[n], v)


def test_tiff_container_is_read_by_libtiff_and_pillow():
    from PIL import Image
    rng = np.random.default_rng(2)
    for a in (rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), rng.integers(0, 256, (37, 53), dtype=np.uint8),
              rng.integers(0, 65536, (20, 31, 3), dtype=np.uint16), rng.integers(0, 65536, (5, 4), dtype=np.uint16),
              rng.integers(0, 256, (1, 1, 3), dtype=np.uint8)):
        t = E._tiff_container(a.dtype, a.shape, zlib.compress(a.tobytes()))
        back = cv2.imdecode(np.frombuffer(t, np.uint8), cv2.IMREAD_UNCHANGED)
        assert back is not None and back.dtype == a.dtype
        if back.ndim == 3:
            back = cv2.cvtColor(back, cv2.COLOR_BGR2RGB)
        assert np.array_equal(back.reshape(a.shape), a)
        if a.dtype == np.uint8:
            assert np.array_equal(np.array(Image.open(io.BytesIO(t))).reshape(a.shape), a)


def test_row_geometry_of_the_arrays_the_entropy_stage_gets():
    """Bytes per row and per sample handed to vcfb_deflate_rows_dev: the H x W x 3 uint8 index image of
    src/2D-DCT.py:361-364, 16-bit images (src/TIFF.py:26), batches, planes, flat arrays."""
    from vcf_b200.entropy import row_geometry
    assert row_geometry(np.zeros((4, 10, 3), np.uint8)) == (30, 3)
    assert row_geometry(np.zeros((4, 10, 3), np.uint16)) == (60, 6)
    assert row_geometry(np.zeros((2, 4, 10, 3), np.uint8)) == (30, 3)
    assert row_geometry(np.zeros((4, 10), np.uint8)) == (10, 1)
    assert row_geometry(np.zeros((4, 10), np.int16)) == (20, 2)
    assert row_geometry(np.zeros(40, np.uint8)) == (0, 1)
    import torch
    assert row_geometry(torch.zeros((4, 10, 3), dtype=torch.uint8)) == (30, 3)
    assert row_geometry(torch.zeros((4, 10), dtype=torch.int16)) == (20, 2)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from vcf_b200 import VcfbError
    x = np.zeros(100, np.uint8)
    for call in (E.deflate_raw, E.zlib_compress, E.crc32, E.adler32):
        with pytest.raises(VcfbError):
            call(x)
    with pytest.raises(VcfbError):
        E.tiff_zlib(x.reshape(10, 10))
    with pytest.raises(VcfbError):
        E.savez_compressed(io.BytesIO(), a=x)
