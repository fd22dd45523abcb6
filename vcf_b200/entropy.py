"""Entropy front-end on the GPU: deflate streams the reference's own
decoders read.

The reference compresses the uint8 index planes of the transform path with zlib on the host --
``np.savez_compressed`` in /root/reference/src/z_lib.py:19-23, tifffile's zlib codec in
src/TIFF.py:23-31 -- and once the transform runs on the GPU that call is the slowest stage left
(``entropy_stage`` in bench.py).  ``vcfb_deflate_dev`` (csrc/kernels_deflate.cu) produces a raw
deflate stream on the device; this module wraps it in the containers the reference uses:

* ``deflate_raw(x)``          raw RFC 1951 stream        -> ``zlib.decompress(s, -15)``
* ``zlib_compress(x)``        RFC 1950 (78 9C ... adler) -> ``zlib.decompress(s)``
* ``savez_compressed(f, a=x)`` .npz (zip, method 8)      -> ``np.load(f)['a']`` as in src/z_lib.py:25-29
* ``tiff_zlib(x)``            TIFF, one zlib strip       -> ``tifffile.imread`` as in src/TIFF.py:33-39

The streams are not byte-identical with zlib's (a different, run-length parse); what is kept is
that the reference's decoder returns the same array.  The CRC-32 of a zip member is taken on the
GPU as well (``vcfb_crc32_dev``: a CUDA tensor is never copied to the host, only its stream and
four bytes of checksum are), and so is the Adler-32 that closes a zlib stream
(``vcfb_adler32_dev``).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import io
import struct
import zlib

import numpy as np

from . import _lib
from ._lib import check


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _as_device_bytes(x):
    """uint8 view (1-D, contiguous, 8-byte aligned) of a numpy array or torch tensor on the GPU."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.VcfbError("no CUDA device: vcf_b200 has no CPU fallback")
    if not _is_torch(x):
        x = torch.from_numpy(np.ascontiguousarray(x).reshape(-1).view(np.uint8))
    if not x.is_cuda:
        x = x.cuda(non_blocking=True)
    x = x.contiguous().reshape(-1).view(torch.uint8)
    if x.data_ptr() % 8:
        x = x.clone()
    return x


def row_geometry(x):
    """``(row_bytes, sample_bytes)`` of an image-like array for the match candidates of the parse
    (``vcfb_deflate_rows_dev``): for (..., H, W, C) the bytes of a row
    and of a pixel, for (H, W) the bytes of a row and of a sample; ``(0, 1)`` (runs only) otherwise."""
    shape = tuple(x.shape)
    item = x.element_size() if _is_torch(x) else x.dtype.itemsize
    if len(shape) == 2:
        return shape[1] * item, item
    if len(shape) >= 3:
        return shape[-2] * shape[-1] * item, shape[-1] * item
    return 0, 1


def deflate_raw_dev(x, geometry=None):
    """Asynchronous on torch's current stream: returns ``(stream, n)`` -- a uint8 CUDA tensor of
    capacity ``vcfb_deflate_bound`` and a one-element int64 CUDA tensor with the stream length.
    ``geometry``: ``(row_bytes, sample_bytes)``; by default from the shape of ``x``."""
    import torch
    row_bytes, sample_bytes = geometry if geometry is not None else row_geometry(x)
    if not 1 <= sample_bytes <= 16:
        row_bytes, sample_bytes = 0, 1
    x = _as_device_bytes(x)
    L = _lib.lib()
    n = x.numel()
    dst = torch.empty(L.vcfb_deflate_bound(n), dtype=torch.uint8, device=x.device)
    ws = torch.empty(L.vcfb_deflate_workspace(n), dtype=torch.uint8, device=x.device)
    out_n = torch.zeros(1, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        check(L.vcfb_deflate_rows_dev(x.data_ptr(), n, row_bytes, sample_bytes, dst.data_ptr(), dst.numel(),
                                      out_n.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    return dst, out_n


def crc32_dev(x):
    """Asynchronous on torch's current stream: one-element CUDA tensor (int64 holding the uint32)
    with ``zlib.crc32`` of the bytes of ``x``."""
    import torch
    x = _as_device_bytes(x)
    out = torch.zeros(1, dtype=torch.int64, device=x.device)      # the low 4 bytes are written
    with torch.cuda.device(x.device):
        check(_lib.lib().vcfb_crc32_dev(x.data_ptr(), x.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


def crc32(x) -> int:
    """``zlib.crc32(bytes of x)`` computed on the GPU."""
    return int(crc32_dev(x).item()) & 0xFFFFFFFF


_CRC_POLY = 0xEDB88320


def _multmodp(a: int, b: int) -> int:
    """a(x) * b(x) mod the CRC-32 polynomial; bit 31 is the coefficient of x^0 (csrc/crc32_core.cuh)."""
    p = 0
    m = 1 << 31
    while m:
        if a & m:
            p ^= b
        b = (b >> 1) ^ _CRC_POLY if b & 1 else b >> 1
        m >>= 1
    return p


def crc32_combine(crc1: int, crc2: int, len2: int) -> int:
    """CRC-32 of A || B from crc32(A), crc32(B) and len(B) (zlib's crc32_combine)."""
    p, sq, n = 1 << 31, 1 << 30, 8 * len2          # x^0, x^1, exponent in bits
    while n:
        if n & 1:
            p = _multmodp(sq, p)
        sq = _multmodp(sq, sq)
        n >>= 1
    return _multmodp(p, crc1) ^ crc2


def deflate_raw(x, geometry=None) -> bytes:
    """Raw deflate stream of the bytes of ``x`` (numpy array or torch tensor)."""
    dst, out_n = deflate_raw_dev(x, geometry)
    n = int(out_n.item())
    return dst[:n].cpu().numpy().tobytes()


def adler32_dev(x):
    """Asynchronous on torch's current stream: one-element CUDA tensor (int64 holding the uint32)
    with ``zlib.adler32`` of the bytes of ``x``."""
    import torch
    x = _as_device_bytes(x)
    out = torch.zeros(3, dtype=torch.int64, device=x.device)      # [result, 16 bytes of workspace]
    with torch.cuda.device(x.device):
        check(_lib.lib().vcfb_adler32_dev(x.data_ptr(), x.numel(), out.data_ptr(), out.data_ptr() + 8,
                                          torch.cuda.current_stream().cuda_stream))
    return out[:1]


def adler32(x) -> int:
    """``zlib.adler32(bytes of x)`` computed on the GPU."""
    return int(adler32_dev(x).item()) & 0xFFFFFFFF


def zlib_compress(x) -> bytes:
    """zlib-format stream (what ``zlib.compress`` returns): ``zlib.decompress`` reads it.  Deflate
    stream and Adler-32 both come from the GPU; ``x``: numpy array or torch tensor."""
    dev = _as_device_bytes(x)
    dst, out_n = deflate_raw_dev(dev, row_geometry(x))
    ad = adler32_dev(dev)
    n = int(out_n.item())
    return b"\x78\x9c" + dst[:n].cpu().numpy().tobytes() + struct.pack(">I", int(ad.item()) & 0xFFFFFFFF)


def tiff_zlib(x) -> bytes:
    """A TIFF file holding the image ``x`` -- uint8 or uint16, (H, W) or (H, W, 3), numpy array or
    CUDA tensor -- as one strip compressed with zlib (Compression = 8, "Adobe deflate"): what
    ``tifffile.imwrite(f, data=x, compression='zlib')`` produces in src/TIFF.py:23-31, with the
    strip's zlib stream made on the GPU.  Baseline TIFF 6.0 + the deflate extension, little endian;
    tifffile, libtiff (OpenCV) and Pillow read it."""
    if _is_torch(x):
        dtype, shape = torch_empty_numpy_dtype(x), tuple(x.shape)
        src = x.detach().contiguous()
    else:
        src = np.ascontiguousarray(x)
        dtype, shape = src.dtype, src.shape
    if dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
        raise ValueError(f"tiff_zlib: uint8 or uint16 images, got {dtype}")     # the assert of src/TIFF.py:26
    if not (len(shape) == 2 or (len(shape) == 3 and shape[2] == 3)) or 0 in shape:
        raise ValueError(f"tiff_zlib: (H, W) or (H, W, 3) images, got shape {shape}")
    return _tiff_container(dtype, shape, zlib_compress(src))


def _tiff_container(dtype, shape, data: bytes) -> bytes:
    """Little-endian baseline TIFF around one zlib-compressed strip ``data`` holding a C-contiguous
    image of this dtype (uint8 / uint16) and shape ((H, W) or (H, W, 3)).  Host logic only."""
    H, W = shape[:2]
    spp = 1 if len(shape) == 2 else 3
    bits = 8 * np.dtype(dtype).itemsize
    if 8 + len(data) + 512 >= 1 << 32:
        raise ValueError("image too large for a classic TIFF")
    pad = len(data) & 1                                    # the IFD starts on a word boundary
    ifd_off = 8 + len(data) + pad
    SHORT, LONG = 3, 4
    entries = []

    def entry(tag, typ, count, value):
        entries.append(struct.pack("<HHI", tag, typ, count) + value)

    n_entries = 11
    extra_off = ifd_off + 2 + 12 * n_entries + 4
    entry(256, LONG, 1, struct.pack("<I", W))                                  # ImageWidth
    entry(257, LONG, 1, struct.pack("<I", H))                                  # ImageLength
    if spp == 1:
        entry(258, SHORT, 1, struct.pack("<HH", bits, 0))                      # BitsPerSample
        extra = b""
    else:
        entry(258, SHORT, spp, struct.pack("<I", extra_off))
        extra = struct.pack("<%dH" % spp, *([bits] * spp))
    entry(259, SHORT, 1, struct.pack("<HH", 8, 0))                             # Compression: deflate
    entry(262, SHORT, 1, struct.pack("<HH", 2 if spp == 3 else 1, 0))          # Photometric: RGB / min-is-black
    entry(273, LONG, 1, struct.pack("<I", 8))                                  # StripOffsets
    entry(277, SHORT, 1, struct.pack("<HH", spp, 0))                           # SamplesPerPixel
    entry(278, LONG, 1, struct.pack("<I", H))                                  # RowsPerStrip
    entry(279, LONG, 1, struct.pack("<I", len(data)))                          # StripByteCounts
    entry(284, SHORT, 1, struct.pack("<HH", 1, 0))                             # PlanarConfiguration: chunky
    entry(339, SHORT, 1, struct.pack("<HH", 1, 0))                             # SampleFormat: unsigned
    assert len(entries) == n_entries
    return (b"II*\x00" + struct.pack("<I", ifd_off) + data + b"\x00" * pad + struct.pack("<H", n_entries)
            + b"".join(entries) + struct.pack("<I", 0) + extra)


def _npy_header(dtype: np.dtype, shape) -> bytes:
    """The .npy header np.save writes for a C-contiguous array of this dtype and shape."""
    fh = io.BytesIO()
    np.lib.format.write_array_header_1_0(fh, {"descr": np.lib.format.dtype_to_descr(np.dtype(dtype)),
                                              "fortran_order": False, "shape": tuple(int(d) for d in shape)})
    return fh.getvalue()


def torch_empty_numpy_dtype(t) -> np.dtype:
    """numpy dtype of a torch tensor, without touching its data"""
    import torch
    return torch.empty(0, dtype=t.dtype).numpy().dtype


def _stored_block(data: bytes) -> bytes:
    """Non-final stored deflate block(s) holding ``data`` (RFC 1951 3.2.4)."""
    out = []
    for i in range(0, len(data), 65535):
        part = data[i:i + 65535]
        out.append(b"\x00" + struct.pack("<HH", len(part), len(part) ^ 0xFFFF) + part)
    return b"".join(out)


def savez_compressed(file, **arrays) -> None:
    """``np.savez_compressed(file, **arrays)`` with the deflate streams produced on the GPU.
    ``file``: path or binary file object.  ``np.load`` reads the result (src/z_lib.py:25-29)."""
    members = []
    for name, arr in arrays.items():
        if _is_torch(arr):
            dtype = torch_empty_numpy_dtype(arr)
            shape = tuple(arr.shape)
            src = arr.detach().contiguous()
        else:
            src = np.ascontiguousarray(arr)
            dtype, shape = src.dtype, src.shape
        if dtype.hasobject:
            raise ValueError("object arrays are not supported")
        header = _npy_header(dtype, shape)
        dev = _as_device_bytes(src)                    # one upload (numpy) or none (CUDA tensor)
        dst, out_n = deflate_raw_dev(dev, row_geometry(src))
        crc_body = crc32_dev(dev)
        nbytes = int(out_n.item())
        members.append((name, header, dst[:nbytes].cpu().numpy().tobytes(), int(crc_body.item()) & 0xFFFFFFFF, dev.numel()))
    _npz_container(file, members)


def _npz_container(file, members) -> None:
    """The zip archive np.savez_compressed writes, from finished parts.  ``members``: (name, .npy
    header bytes, raw deflate stream of the array's bytes, crc32 of the array's bytes, their count).
    Host logic only."""
    own = isinstance(file, (str, bytes)) or hasattr(file, "__fspath__")
    fh = open(file, "wb") if own else file
    try:
        start = fh.tell()
        central = []
        for name, header, stream, crc_body, nbody in members:
            # the member's stream: the .npy header as a stored block, then the array's blocks
            comp = _stored_block(header) + stream
            crc = crc32_combine(zlib.crc32(header) & 0xFFFFFFFF, crc_body, nbody)
            usize = len(header) + nbody
            if usize >= 0xFFFFFFFF or len(comp) >= 0xFFFFFFFF:
                raise ValueError("array too large for a zip member without zip64")
            fname = (name + ".npy").encode()
            offset = fh.tell() - start
            # local file header: version 20, no flags, method 8 (deflate), DOS time/date 0 / 1980-01-01
            fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, 8, 0, 0x21, crc, len(comp), usize, len(fname), 0))
            fh.write(fname)
            fh.write(comp)
            central.append(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0, 8, 0, 0x21, crc, len(comp), usize,
                                       len(fname), 0, 0, 0, 0, 0, offset) + fname)
        cd_off = fh.tell() - start
        cd = b"".join(central)
        fh.write(cd)
        fh.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, len(central), len(central), len(cd), cd_off, 0))
    finally:
        if own:
            fh.close()


_CS_STREAMS = {}


def encode_to_codestream(codec, rgb, n_streams: int = 3):
    """Host RGB frames -> per-frame raw deflate code-streams, with everything between on the GPU.

    ``codec``: a :class:`vcf_b200.Codec`; ``rgb``: numpy uint8 (n,H,W,3), ideally in pinned memory
    (``vcf_b200.pinned_empty`` / a pinned torch tensor's ``.numpy()``).  Frame f is uploaded,
    transformed (src/2D-DCT.py:276-361) and deflated (the zlib call of src/z_lib.py:19-23 /
    src/TIFF.py:23-31) on CUDA stream f mod ``n_streams``, so uploads overlap kernels; only the
    code-streams cross PCIe on the way back (3 B/pixel up, ~0.1 B/pixel down instead of 3 + 3).
    Returns a list of ``bytes``, one complete deflate stream per frame
    (``zlib.decompress(s, -15)`` gives the frame's uint8 index array)."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.VcfbError("no CUDA device: vcf_b200 has no CPU fallback")
    if rgb.ndim == 3:
        rgb = rgb[None]
    dev = torch.device("cuda", int(codec.device) if codec.device is not None else torch.cuda.current_device())
    key = (dev.index, n_streams)
    if key not in _CS_STREAMS:
        with torch.cuda.device(dev):
            _CS_STREAMS[key] = [torch.cuda.Stream() for _ in range(n_streams)]
    streams = _CS_STREAMS[key]
    cur = torch.cuda.current_stream(dev)
    pend = []
    with torch.cuda.device(dev):
        for st in streams:
            st.wait_stream(cur)
        for f in range(rgb.shape[0]):
            st = streams[f % n_streams]
            with torch.cuda.stream(st):
                xd = torch.from_numpy(rgb[f]).to(dev, non_blocking=True)
                k = codec.encode(xd)
                dst, nb = deflate_raw_dev(k)
                pend.append((dst, nb))
        for st in streams:
            st.synchronize()
        lens = torch.cat([nb for _, nb in pend]).cpu().tolist()
        out = []
        for (dst, _), n in zip(pend, lens):
            out.append(dst[:int(n)].cpu().numpy().tobytes())
    return out
