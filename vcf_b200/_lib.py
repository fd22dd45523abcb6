"""ctypes binding of libvcfb200.so (include/vcfb200.h).

The shared object is built in-tree by ``make`` / ``__graft_entry__.build()``.
There is no fallback of any kind: if the library is missing, or no CUDA device
is present when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvcfb200.so")

STAT_SSE_R, STAT_SSE_G, STAT_SSE_B, STAT_NSAMPLES = 0, 1, 2, 3
STAT_NONZERO, STAT_SUMABS, STAT_NINDICES, STAT_SUMDIFF, STAT_HIST = 4, 5, 6, 7, 8
STAT_LEN = 8 + 3 * 256

COLOR_YCOCG, COLOR_YCRCB = 0, 1
F_NO_SUBBANDS, F_PERCEPTUAL, F_FP64, F_CONTRACT, F_HIST, F_SYNTH_F32, F_FAST, F_NOWRAP, F_NO_OFFSET = 1, 2, 4, 8, 16, 32, 64, 128, 256
RD_MAX_STEPS = 16


class VcfbError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VcfbError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "vcf_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, d, u = C.c_void_p, C.c_int, C.c_double, C.c_uint
    L.vcfb_version.restype = i
    L.vcfb_last_error.restype = C.c_char_p
    L.vcfb_device_count.restype = i
    L.vcfb_last_kernel.restype = C.c_char_p
    L.vcfb_launch_count.restype = C.c_longlong
    L.vcfb_padded_dims.argtypes = [i, i, i] + [C.POINTER(i)] * 4
    L.vcfb_padded_dims.restype = i
    L.vcfb_encode_dev.argtypes = [vp, i, i, i, i, d, i, u, vp, vp, vp, vp]
    L.vcfb_encode_dev.restype = i
    L.vcfb_decode_dev.argtypes = [vp, i, i, i, i, d, i, u, vp, vp, vp, vp, vp, vp]
    L.vcfb_decode_dev.restype = i
    L.vcfb_rd_sweep_dev.argtypes = [vp, i, i, i, i, C.POINTER(d), i, i, u, vp, vp]
    L.vcfb_rd_sweep_dev.restype = i
    L.vcfb_ctx_create.argtypes = [i, C.POINTER(vp)]
    L.vcfb_ctx_create.restype = i
    L.vcfb_ctx_destroy.argtypes = [vp]
    L.vcfb_ctx_destroy.restype = None
    L.vcfb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.vcfb_host_alloc.restype = i
    L.vcfb_host_free.argtypes = [vp]
    L.vcfb_host_free.restype = None
    ll = C.c_longlong
    L.vcfb_color_encode_dev.argtypes = [vp, ll, d, i, vp, vp]
    L.vcfb_color_encode_dev.restype = i
    L.vcfb_color_decode_dev.argtypes = [vp, ll, d, i, vp, vp]
    L.vcfb_color_decode_dev.restype = i
    L.vcfb_color_encode_host.argtypes = [vp, vp, ll, d, i, vp]
    L.vcfb_color_encode_host.restype = i
    L.vcfb_color_decode_host.argtypes = [vp, vp, ll, d, i, vp]
    L.vcfb_color_decode_host.restype = i
    L.vcfb_encode_host.argtypes = [vp, vp, i, i, i, i, d, i, u, vp, vp, vp]
    L.vcfb_encode_host.restype = i
    L.vcfb_decode_host.argtypes = [vp, vp, i, i, i, i, d, i, u, vp, vp, vp, vp, vp]
    L.vcfb_decode_host.restype = i
    L.vcfb_gray_dev.argtypes = [vp, ll, vp, vp]
    L.vcfb_gray_dev.restype = i
    L.vcfb_block_match_dev.argtypes = [vp, vp, i, i, i, i, i, vp, vp]
    L.vcfb_block_match_dev.restype = i
    L.vcfb_block_match_tss_dev.argtypes = [vp, vp, i, i, i, i, i, vp, vp]
    L.vcfb_block_match_tss_dev.restype = i
    sz = C.c_size_t
    L.vcfb_deflate_bound.argtypes = [sz]
    L.vcfb_deflate_bound.restype = sz
    L.vcfb_deflate_workspace.argtypes = [sz]
    L.vcfb_deflate_workspace.restype = sz
    L.vcfb_deflate_dev.argtypes = [vp, sz, vp, sz, vp, vp, sz, vp]
    L.vcfb_deflate_dev.restype = i
    L.vcfb_deflate_rows_dev.argtypes = [vp, sz, sz, i, vp, sz, vp, vp, sz, vp]
    L.vcfb_deflate_rows_dev.restype = i
    L.vcfb_crc32_dev.argtypes = [vp, sz, vp, vp]
    L.vcfb_crc32_dev.restype = i
    L.vcfb_adler32_dev.argtypes = [vp, sz, vp, vp, vp]
    L.vcfb_adler32_dev.restype = i
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise VcfbError(f"libvcfb200 error {rc}: {lib().vcfb_last_error().decode()}")


def padded_dims(H: int, W: int, B: int):
    """(Hp, Wp, top, left) of src/2D-DCT.py:208-219."""
    a = [C.c_int() for _ in range(4)]
    check(lib().vcfb_padded_dims(H, W, B, *[C.byref(x) for x in a]))
    return tuple(x.value for x in a)


def launch_count() -> int:
    """Kernels launched so far by the calling thread (vcfb_launch_count)."""
    return int(lib().vcfb_launch_count())


def last_kernel() -> str:
    """Kernel family the calling thread launched last (vcfb_last_kernel)."""
    return lib().vcfb_last_kernel().decode()
