"""Row F4 (entropy front-end): the deflate streams the GPU writes (vcfb_deflate_dev through
vcf_b200.entropy) are read back by the decoders the reference uses -- zlib.decompress and np.load
(/root/reference/src/z_lib.py:25-29) -- and give the input bytes.  The same edge cases as the CPU
emulation in tests/test_deflate_core.py, which runs the identical __host__ __device__ code."""
import io
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _back(data, geometry=None):
    from vcf_b200 import _lib
    from vcf_b200.entropy import deflate_raw
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    raw = deflate_raw(data, geometry)
    assert _lib.last_kernel() in ("deflate_gather", "deflate_scan")
    assert len(raw) <= _lib.lib().vcfb_deflate_bound(data.size)
    assert zlib.decompress(raw, -15) == data.tobytes()
    return len(raw)


def test_edge_cases():
    rng = np.random.default_rng(7)
    cases = [
        np.zeros(0, np.uint8),
        np.array([5], np.uint8), np.array([5, 5], np.uint8), np.array([5, 5, 5], np.uint8), np.array([5, 5, 5, 5], np.uint8),
        np.full(257, 128, np.uint8), np.full(258, 128, np.uint8), np.full(259, 128, np.uint8),
        np.full(260, 128, np.uint8), np.full(261, 128, np.uint8), np.full(262, 128, np.uint8),
        np.full(100000, 128, np.uint8),
        np.full(258 * 512, 9, np.uint8), np.full(258 * 512 + 1, 9, np.uint8), np.full(258 * 512 - 1, 9, np.uint8),
        np.arange(256, dtype=np.uint8),
        np.tile(np.arange(256, dtype=np.uint8), 300),
        rng.integers(0, 256, 70001, dtype=np.uint8),
        rng.integers(0, 256, 65536 * 3 + 17, dtype=np.uint8),
        rng.integers(0, 2, 50000, dtype=np.uint8),
        np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 600, 3000)),
    ]
    for data in cases:
        _back(data)


def test_every_run_length_at_every_alignment():
    for L in list(range(1, 40)) + [255, 256, 257, 258, 259, 260, 300, 515, 516, 517, 518]:
        for lead in range(0, 9):
            data = np.concatenate([np.arange(1, lead + 1, dtype=np.uint8), np.full(L, 200, np.uint8), np.array([3, 3, 9], np.uint8)])
            _back(data)


def test_noise_falls_back_to_stored_blocks():
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 2_000_000, dtype=np.uint8)
    size = _back(data)
    nseg = -(-data.size // (258 * 512))
    assert size <= data.size + 5 * 3 * nseg + 2


def test_skewed_alphabet_length_limit():
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    rng = np.random.default_rng(5)
    sym = np.concatenate([np.full(f, i, np.uint8) for i, f in enumerate(fib[:24])])
    rng.shuffle(sym)
    sym = sym[sym != np.roll(sym, 1)]
    _back(sym)


def test_random_fuzz():
    rng = np.random.default_rng(11)
    for it in range(40):
        n = int(rng.integers(1, 3_000_000 if it % 5 == 0 else 300_000))
        k = int(rng.integers(1, 256))
        p = rng.dirichlet(np.full(k, 0.3))
        data = rng.choice(k, size=n, p=p).astype(np.uint8)
        if it % 3 == 0:
            data = np.repeat(data[: n // 8 + 1], rng.integers(1, 20, n // 8 + 1))
        _back(data)


def test_row_candidates_fuzz_and_emulation():
    """Rows as match candidates (vcfb_deflate_rows_dev): round trip on repeated / perturbed rows of
    every kind of length, and the stream is byte for byte the one the host emulation of the same
    __host__ __device__ code writes (tests/emul/deflate_emul.cpp, built here with g++)."""
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "build", "deflate_emul.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", so,
                           os.path.join(root, "tests", "emul", "deflate_emul.cpp")])
    E = ctypes.CDLL(so)
    E.dfl_emul.restype = ctypes.c_longlong
    E.dfl_emul.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_int,
                           ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
    from vcf_b200.entropy import deflate_raw
    rng = np.random.default_rng(21)
    for it in range(48):
        row = int(rng.choice([1, 2, 3, 5, 7, 48, 255, 258, 300, 777, 5760, 11520, 32765, 32767, 32768, 40000]))
        px = int(rng.choice([1, 1, 2, 3, 4]))
        nrows = int(rng.integers(1, max(2, 600000 // row)))
        k = int(rng.integers(1, 40))
        base = rng.choice(k, size=row, p=rng.dirichlet(np.full(k, 0.4))).astype(np.uint8)
        rows = np.tile(base, (nrows, 1))
        flips = rng.random(rows.shape) < float(rng.choice([0.0, 0.002, 0.05, 0.5]))
        rows[flips] = rng.integers(0, 256, int(flips.sum()), dtype=np.uint8)
        if it % 4 == 0:
            rows = np.cumsum(rows, axis=0, dtype=np.uint8)
        data = np.ascontiguousarray(rows.ravel()[: int(rng.integers(1, rows.size + 1))])
        raw = deflate_raw(data, (row, px))
        assert zlib.decompress(raw, -15) == data.tobytes(), (it, row, px)
        out = np.empty(data.size + data.size // 50 + 4096, np.uint8)
        n = E.dfl_emul(data.ctypes.data, data.size, 258, 512, row, px, None, 0, 1, out.ctypes.data, out.size, None)
        assert n == len(raw) and out[:n].tobytes() == raw, (it, row, px, n, len(raw))


@pytest.mark.parametrize("q", [4, 16, 32, 64])
def test_index_planes_size_against_zlib(q):
    """Indices of the transform path (one 4K frame, produced on the GPU): round trip; the size is at
    most that of zlib's own run-length strategy (Z_RLE), and within a stated factor of zlib level 6
    with its 32 KB hash-chain matcher -- the rate the reference's RD curves use
    (profiles/r2_deflate_candidates.json; round 1's run-length parse: 1.22x at q=16, 1.39x at q=32,
    1.70x at q=64)."""
    import torch
    from oracle import vcf_oracle as O
    from vcf_b200 import Codec
    from vcf_b200.entropy import deflate_raw_dev
    img = O.synthetic_frame(2160, 3840, 2, "natural")
    k = Codec(8, q).encode(torch.from_numpy(img[None]).cuda())[0]
    dst, n = deflate_raw_dev(k)
    n = int(n.item())
    raw = dst[:n].cpu().numpy().tobytes()
    host = k.cpu().numpy()
    assert zlib.decompress(raw, -15) == host.tobytes()
    ref = len(zlib.compress(host.tobytes(), 6))
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_RLE)
    rle = len(c.compress(host.tobytes()) + c.flush())
    nseg = -(-host.size // (258 * 512))
    assert n <= 1.01 * rle + 100 * nseg, (q, n, rle)
    assert n <= {4: 1.0, 16: 1.2, 32: 1.2, 64: 1.2}[q] * ref + 100 * nseg, (q, n, ref)
    flat = deflate_raw_dev(k.reshape(-1))                     # no geometry: runs only
    assert int(flat[1].item()) >= n


def test_large_input():
    """A multi-frame batch in one call (70 MiB, 556 segments)."""
    import torch
    from vcf_b200.entropy import deflate_raw
    rng = np.random.default_rng(1)
    base = np.repeat(rng.integers(120, 136, 2_000_000, dtype=np.uint8), rng.integers(1, 80, 2_000_000))[: 70 * (1 << 20)]
    x = torch.from_numpy(base).cuda()
    raw = deflate_raw(x)
    assert zlib.decompress(raw, -15) == base.tobytes()


def test_containers_are_read_by_the_reference_decoders():
    from vcf_b200.entropy import savez_compressed, zlib_compress
    rng = np.random.default_rng(2)
    a = np.repeat(rng.integers(100, 150, 60000, dtype=np.uint8), rng.integers(1, 30, 60000))[:600000].reshape(200, 1000, 3)
    assert zlib.decompress(zlib_compress(a)) == a.tobytes()
    b = (a.astype(np.int16) - 128)
    fh = io.BytesIO()
    savez_compressed(fh, a=a, b=b)
    fh.seek(0)
    z = np.load(fh)                      # src/z_lib.py:25-29
    assert z["a"].dtype == np.uint8 and np.array_equal(z["a"], a)
    assert z["b"].dtype == np.int16 and np.array_equal(z["b"], b)


def test_argument_errors():
    import torch
    from vcf_b200 import VcfbError, _lib
    L = _lib.lib()
    x = torch.zeros(1000, dtype=torch.uint8, device="cuda")
    dst = torch.zeros(10, dtype=torch.uint8, device="cuda")
    ws = torch.zeros(L.vcfb_deflate_workspace(1000), dtype=torch.uint8, device="cuda")
    n = torch.zeros(1, dtype=torch.int64, device="cuda")
    with pytest.raises(VcfbError):
        _lib.check(L.vcfb_deflate_dev(x.data_ptr(), 1000, dst.data_ptr(), dst.numel(), n.data_ptr(), ws.data_ptr(), ws.numel(), None))
    big = torch.zeros(L.vcfb_deflate_bound(1000), dtype=torch.uint8, device="cuda")
    with pytest.raises(VcfbError):
        _lib.check(L.vcfb_deflate_dev(x.data_ptr(), 1000, big.data_ptr(), big.numel(), n.data_ptr(), ws.data_ptr(), 8, None))
    for bad_sample in (0, -1, 17):
        with pytest.raises(VcfbError):
            _lib.check(L.vcfb_deflate_rows_dev(x.data_ptr(), 1000, 100, bad_sample, big.data_ptr(), big.numel(), n.data_ptr(),
                                               ws.data_ptr(), ws.numel(), None))
    # a row beyond deflate's window is not an error: the candidates above are dropped
    _lib.check(L.vcfb_deflate_rows_dev(x.data_ptr(), 1000, 1 << 20, 3, big.data_ptr(), big.numel(), n.data_ptr(), ws.data_ptr(),
                                       ws.numel(), None))
    torch.cuda.synchronize()
    assert zlib.decompress(big[: int(n.item())].cpu().numpy().tobytes(), -15) == bytes(1000)


def test_crc32_equals_zlib():
    import torch
    from vcf_b200 import _lib
    from vcf_b200.entropy import crc32, crc32_combine
    rng = np.random.default_rng(4)
    for n in (0, 1, 7, 8, 9, 511, 512, 513, 4096, 100003, 512 * 256 * 3 + 5, 30_000_001):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        assert crc32(data) == (zlib.crc32(data.tobytes()) & 0xFFFFFFFF), n
        if n:
            assert _lib.last_kernel() == "crc32"
    x = torch.from_numpy(rng.integers(-500, 500, (33, 77, 3), dtype=np.int16)).cuda()
    assert crc32(x) == (zlib.crc32(x.cpu().numpy().tobytes()) & 0xFFFFFFFF)
    a, b = rng.integers(0, 256, 1000, dtype=np.uint8), rng.integers(0, 256, 77777, dtype=np.uint8)
    assert crc32_combine(crc32(a), crc32(b), b.size) == (zlib.crc32(np.concatenate([a, b]).tobytes()) & 0xFFFFFFFF)


def test_savez_compressed_from_cuda_tensors():
    """The batched driver hands CUDA tensors over: no host copy of the array is made, and the zip
    member's CRC-32 (checked by np.load / zipfile on reading) comes from the GPU."""
    import zipfile
    import torch
    from vcf_b200.entropy import savez_compressed
    rng = np.random.default_rng(6)
    a = np.repeat(rng.integers(120, 136, 90000, dtype=np.uint8), rng.integers(1, 40, 90000))[:1_500_000].reshape(500, 1000, 3)
    t = torch.from_numpy(a).cuda()
    e = torch.zeros((0, 4), dtype=torch.float32, device="cuda")
    fh = io.BytesIO()
    savez_compressed(fh, a=t, k=(t.to(torch.int16) - 128), empty=e)
    fh.seek(0)
    assert zipfile.ZipFile(fh).testzip() is None            # every member's CRC-32 verifies
    fh.seek(0)
    z = np.load(fh)
    assert z["a"].dtype == np.uint8 and np.array_equal(z["a"], a)
    assert z["k"].dtype == np.int16 and np.array_equal(z["k"], a.astype(np.int16) - 128)
    assert z["empty"].shape == (0, 4) and z["empty"].dtype == np.float32


def test_adler32_equals_zlib_and_zlib_stream_from_cuda_tensor():
    import torch
    from vcf_b200 import _lib
    from vcf_b200.entropy import adler32, zlib_compress
    rng = np.random.default_rng(8)
    for n in (0, 1, 7, 8, 9, 15, 16, 17, 4099, 100003, 65521 * 8 + 3, 30_000_001):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        assert adler32(data) == (zlib.adler32(data.tobytes()) & 0xFFFFFFFF), n
        assert _lib.last_kernel() == "adler32"
    ones = np.full(20_000_000, 255, np.uint8)               # the largest sums
    assert adler32(ones) == (zlib.adler32(ones.tobytes()) & 0xFFFFFFFF)
    x = torch.from_numpy(np.repeat(rng.integers(0, 256, 5000, dtype=np.uint8), rng.integers(1, 100, 5000))).cuda()
    assert zlib.decompress(zlib_compress(x)) == x.cpu().numpy().tobytes()      # inflate checks the Adler-32


def test_tiff_zlib_is_read_by_libtiff_and_pillow():
    """The TIFF container of the reference's default entropy stage (src/TIFF.py:23-31); tifffile is
    not installed here, libtiff (through OpenCV) and Pillow are the independent readers."""
    import cv2
    import torch
    from vcf_b200.entropy import tiff_zlib
    rng = np.random.default_rng(10)
    smooth = np.repeat(rng.integers(100, 160, 70000, dtype=np.uint8), rng.integers(1, 30, 70000))
    cases = [smooth[: 301 * 457 * 3].reshape(301, 457, 3), smooth[: 64 * 64].reshape(64, 64),
             rng.integers(0, 256, (1, 1, 3), dtype=np.uint8),
             (smooth[: 90 * 70 * 3].astype(np.uint16) * 257).reshape(90, 70, 3)]
    for a in cases:
        for src in ((a, torch.from_numpy(a).cuda()) if a.dtype == np.uint8 else (a,)):
            t = tiff_zlib(src)
            back = cv2.imdecode(np.frombuffer(t, np.uint8), cv2.IMREAD_UNCHANGED)
            assert back is not None and back.dtype == a.dtype
            if back.ndim == 3:
                back = cv2.cvtColor(back, cv2.COLOR_BGR2RGB)
            assert np.array_equal(back.reshape(a.shape), a)
        if a.dtype == np.uint8:
            from PIL import Image
            assert np.array_equal(np.array(Image.open(io.BytesIO(tiff_zlib(a)))).reshape(a.shape), a)
    with pytest.raises(ValueError):
        tiff_zlib(np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        tiff_zlib(np.zeros((4, 4, 2), np.uint8))
