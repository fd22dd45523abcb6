"""Array-level host API of the path: uint8 RGB frames <-> uint8 quantisation indices.

Mirrors the arithmetic section of the reference's ``CoDec.encode_fn`` /
``decode_fn`` (/root/reference/src/2D-DCT.py:276-361 and :398-466) with the
same knobs as its CLI flags: ``block_size`` (-B), ``q`` (-q / QSS),
``perceptual`` (-p), ``disable_subbands`` (-x).  All arithmetic happens in
libvcfb200.so on the GPU; this module only moves pointers.

Two calling conventions:

* torch CUDA tensors in -> torch CUDA tensors out (zero copy, asynchronous on
  torch's current stream);
* numpy arrays in -> numpy arrays out (the library stages through pinned memory;
  the call returns when the result is in the output array).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (COLOR_YCOCG, COLOR_YCRCB, F_CONTRACT, F_FP64, F_HIST, F_NO_SUBBANDS, F_PERCEPTUAL, F_FAST, F_SYNTH_F32, F_NO_OFFSET,
                   STAT_HIST, STAT_LEN, VcfbError, check, padded_dims)

_COLORS = {"YCoCg": COLOR_YCOCG, "YCrCb": COLOR_YCRCB}

# JPEG luma / chroma tables of src/2D-DCT.py:66-82
_Y_QSS = np.array([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55],
                   [14, 13, 16, 24, 40, 57, 69, 56], [14, 17, 22, 29, 51, 87, 80, 62],
                   [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                   [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]],
                  dtype=np.uint8)
_C_QSS = np.array([[17, 18, 24, 47, 99, 99, 99, 99], [18, 21, 26, 66, 99, 99, 99, 99],
                   [24, 26, 56, 99, 99, 99, 99, 99], [47, 66, 99, 99, 99, 99, 99, 99],
                   [99] * 8, [99] * 8, [99] * 8, [99] * 8], dtype=np.uint8)


def perceptual_weights(B: int) -> np.ndarray:
    """(2,B,B) float64: Y_QSSs/121 and C_QSSs/99 exactly as the reference builds
    them (src/2D-DCT.py:83-90, :322-324): uint8 tables, ``cv2.resize`` to BxB with
    INTER_AREA below 8 and INTER_LINEAR otherwise."""
    import cv2
    inter = cv2.INTER_AREA if B < 8 else cv2.INTER_LINEAR
    Cq = cv2.resize(_C_QSS, (B, B), interpolation=inter)
    Yq = cv2.resize(_Y_QSS, (B, B), interpolation=inter)
    return np.stack([Yq / 121, Cq / 99]).astype(np.float64)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def stats_dict(vec: np.ndarray) -> dict:
    """Decode the int64 statistics vector (include/vcfb200.h VCFB_STAT_*)."""
    v = np.asarray(vec, dtype=np.int64)
    sse = v[0:3].copy()
    n = int(v[3])
    hist = v[STAT_HIST:STAT_HIST + 768].reshape(3, 256).copy()
    out = dict(sse=sse, nsamples=n, nonzero=int(v[4]), sumabs=int(v[5]), nindices=int(v[6]),
               sumdiff=int(v[7]), hist=hist)
    if n:
        mse = float(sse.sum()) / n
        out["mse"] = mse
        out["rmse"] = math.sqrt(mse)                         # src/RDE.py:49-53
        out["psnr"] = float("inf") if mse == 0 else 10.0 * math.log10(255.0 * 255.0 / mse)
    tot = hist.sum()
    if tot:
        h = hist.astype(np.float64)
        bits = 0.0
        for c in range(3):
            nc = h[c].sum()
            p = h[c][h[c] > 0] / nc
            bits += float(-(p * np.log2(p)).sum() * nc)
        out["entropy_bits"] = bits                             # zero-order estimate
    return out


class Codec:
    """One configuration of the path.  Thread-compatible; one host staging
    context per instance (created on first numpy call)."""

    def __init__(self, block_size: int = 8, q=32, color: str = "YCoCg", perceptual: bool = False,
                 disable_subbands: bool = False, fp64: bool = False, contract: bool = False,
                 device: Optional[int] = None, hist: bool = True, synth_f32: bool = False, fast: bool = False,
                 no_offset: bool = False):
        if color not in _COLORS:
            raise ValueError(f"color must be one of {list(_COLORS)}")
        if block_size not in (2, 4, 8, 16, 32, 64, 128):      # the reference's -L search set (src/2D-DCT.py:538)
            raise ValueError("block_size must be a power of two in [2, 128]")
        if not (float(q) > 0):
            raise ValueError("q must be > 0")
        self.B = int(block_size)
        self.q = float(q)
        self.color = _COLORS[color]
        self.flags = ((F_PERCEPTUAL if perceptual else 0) | (F_NO_SUBBANDS if disable_subbands else 0)
                      | (F_FP64 if fp64 else 0) | (F_CONTRACT if contract else 0)
                      | (F_HIST if hist else 0)       # histogram of the indices in the statistics
                      | (F_FAST if fast else 0)       # encode: tensor-core fast mode (< 1e-6 of the indices differ)
                      | (F_NO_OFFSET if no_offset else 0))   # encode only: the offset-free loop of optimize_block_size
        # decode only: upstream variant "synthesize_image stores float32" (include/vcfb200.h)
        self.synth_f32 = bool(synth_f32)
        if synth_f32 and not fp64:
            raise ValueError("synth_f32 is a variant of the float64 decoder (fp64=True)")
        self.fp64 = bool(fp64)
        self.device = device
        self._weights_np = perceptual_weights(self.B) if perceptual else None
        self._weights_dev = {}
        self._ctx = None

    def _dec_flags(self):
        return (self.flags & ~F_NO_OFFSET) | (F_SYNTH_F32 if self.synth_f32 else 0)

    # -- plumbing ---------------------------------------------------------------
    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.lib().vcfb_ctx_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def _host_ctx(self):
        if self._ctx is None:
            L = _lib.lib()
            if L.vcfb_device_count() < 1:
                raise VcfbError("no CUDA device: vcf_b200 has no CPU fallback")
            h = C.c_void_p()
            check(L.vcfb_ctx_create(int(self.device or 0), C.byref(h)))
            self._ctx = h
        return self._ctx

    def _dev_weights(self, dev):
        if self._weights_np is None:
            return None
        import torch
        t = self._weights_dev.get(dev)
        if t is None:
            t = torch.from_numpy(self._weights_np).to(dev)
            self._weights_dev[dev] = t
        return t

    @staticmethod
    def _frames(x, what):
        if x.ndim == 3:
            x = x[None]
        if x.ndim != 4 or x.shape[-1] != 3:
            raise ValueError(f"{what} must have shape (n,H,W,3) or (H,W,3)")
        return x

    # -- encode -------------------------------------------------------------------
    def encode(self, rgb, stats: bool = False, out=None):
        """uint8 RGB (n,H,W,3)|(H,W,3) -> uint8 indices (n,Hp,Wp,3)|(Hp,Wp,3)
        [, statistics dict].  Replaces src/2D-DCT.py:276-361.
        ``out``: optional preallocated (n,Hp,Wp,3) uint8 result (same kind as the
        input; a pinned numpy array is transferred without a staging copy)."""
        L = _lib.lib()
        single = rgb.ndim == 3
        x = self._frames(rgb, "rgb")
        n, H, W, _ = x.shape
        Hp, Wp, _, _ = padded_dims(H, W, self.B)
        if _is_torch(x):
            import torch
            if x.dtype != torch.uint8 or not x.is_cuda:
                raise ValueError("torch input must be a CUDA uint8 tensor")
            x = x.contiguous()
            if out is None:
                out = torch.empty((n, Hp, Wp, 3), dtype=torch.uint8, device=x.device)
            elif (tuple(out.shape) != (n, Hp, Wp, 3) or out.dtype != torch.uint8 or out.device != x.device
                  or not out.is_contiguous()):
                raise ValueError("out must be a contiguous uint8 tensor (n,Hp,Wp,3) on the input's device")
            st = torch.zeros(STAT_LEN, dtype=torch.int64, device=x.device) if stats else None
            w = self._dev_weights(x.device)
            with torch.cuda.device(x.device):
                stream = torch.cuda.current_stream().cuda_stream
                check(L.vcfb_encode_dev(x.data_ptr(), n, H, W, self.B, self.q, self.color, self.flags,
                                        w.data_ptr() if w is not None else None, out.data_ptr(),
                                        st.data_ptr() if st is not None else None, stream))
            res = out[0] if single else out
            return (res, st) if stats else res
        x = np.ascontiguousarray(x)
        if x.dtype != np.uint8:
            raise ValueError("rgb must be uint8")
        if out is None:
            out = np.empty((n, Hp, Wp, 3), dtype=np.uint8)
        elif out.shape != (n, Hp, Wp, 3) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint8 array (n,Hp,Wp,3)")
        st = np.zeros(STAT_LEN, dtype=np.int64) if stats else None
        w = self._weights_np
        check(L.vcfb_encode_host(self._host_ctx(), x.ctypes.data, n, H, W, self.B, self.q, self.color,
                                 self.flags, w.ctypes.data if w is not None else None,
                                 out.ctypes.data, st.ctypes.data if st is not None else None))
        res = out[0] if single else out
        return (res, stats_dict(st)) if stats else res

    # -- decode -------------------------------------------------------------------
    def decode(self, idx, shape, original=None, stats: bool = False, return_float: bool = False,
               want_rgb: bool = True, out=None):
        """uint8 indices -> uint8 RGB of un-padded ``shape`` = (H, W).
        Replaces src/2D-DCT.py:398-466.

        return_float: also return the un-clipped image the reference hands to
        ``CT.CoDec.filter`` (:461) -- float64 in fp64 mode, float32 otherwise.
        original + stats: accumulate the SSE against ``original`` (src/RDE.py)."""
        L = _lib.lib()
        single = idx.ndim == 3
        k = self._frames(idx, "idx")
        n = k.shape[0]
        H, W = int(shape[0]), int(shape[1])
        Hp, Wp, _, _ = padded_dims(H, W, self.B)
        if tuple(k.shape[1:3]) != (Hp, Wp):
            raise ValueError(f"index array is {tuple(k.shape[1:3])}, expected {(Hp, Wp)} for shape {(H, W)} and B={self.B}")
        if _is_torch(k):
            import torch
            if k.dtype != torch.uint8 or not k.is_cuda:
                raise ValueError("torch input must be a CUDA uint8 tensor")
            k = k.contiguous()
            dev = k.device
            rgb = None
            if want_rgb:
                rgb = out if out is not None else torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
                if (tuple(rgb.shape) != (n, H, W, 3) or rgb.dtype != torch.uint8 or rgb.device != dev
                        or not rgb.is_contiguous()):
                    raise ValueError("out must be a contiguous uint8 tensor (n,H,W,3) on the input's device")
            yf = torch.empty((n, H, W, 3), dtype=torch.float64 if self.fp64 else torch.float32,
                             device=dev) if return_float else None
            org = None
            if original is not None:
                org = self._frames(original, "original").contiguous()
                if tuple(org.shape) != (n, H, W, 3) or org.dtype != torch.uint8:
                    raise ValueError("original must be uint8 with the decoded shape")
            st = torch.zeros(STAT_LEN, dtype=torch.int64, device=dev) if stats else None
            w = self._dev_weights(dev)
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream().cuda_stream
                check(L.vcfb_decode_dev(k.data_ptr(), n, H, W, self.B, self.q, self.color, self._dec_flags(),
                                        w.data_ptr() if w is not None else None,
                                        rgb.data_ptr() if rgb is not None else None,
                                        yf.data_ptr() if yf is not None else None,
                                        org.data_ptr() if org is not None else None,
                                        st.data_ptr() if st is not None else None, stream))
            res = [rgb[0] if (single and rgb is not None) else rgb]
            if return_float:
                res.append(yf[0] if single else yf)
            if stats:
                res.append(st)
            return res[0] if len(res) == 1 else tuple(res)
        k = np.ascontiguousarray(k)
        if k.dtype != np.uint8:
            raise ValueError("idx must be uint8")
        rgb = None
        if want_rgb:
            rgb = out if out is not None else np.empty((n, H, W, 3), dtype=np.uint8)
            if rgb.shape != (n, H, W, 3) or rgb.dtype != np.uint8 or not rgb.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous uint8 array (n,H,W,3)")
        yf = np.empty((n, H, W, 3), dtype=np.float64 if self.fp64 else np.float32) if return_float else None
        org = None
        if original is not None:
            org = np.ascontiguousarray(self._frames(original, "original"))
            if org.shape != (n, H, W, 3) or org.dtype != np.uint8:
                raise ValueError("original must be uint8 with the decoded shape")
        st = np.zeros(STAT_LEN, dtype=np.int64) if stats else None
        w = self._weights_np
        check(L.vcfb_decode_host(self._host_ctx(), k.ctypes.data, n, H, W, self.B, self.q, self.color,
                                 self._dec_flags(), w.ctypes.data if w is not None else None,
                                 rgb.ctypes.data if rgb is not None else None,
                                 yf.ctypes.data if yf is not None else None,
                                 org.ctypes.data if org is not None else None,
                                 st.ctypes.data if st is not None else None))
        res = [rgb[0] if (single and rgb is not None) else rgb]
        if return_float:
            res.append(yf[0] if single else yf)
        if stats:
            res.append(stats_dict(st))
        return res[0] if len(res) == 1 else tuple(res)


class ColorCodec:
    """Stand-alone colour codecs of the reference: ``python YCoCg.py encode|decode``
    (src/YCoCg.py:33-85) and ``python YCrCb.py encode|decode`` (src/YCrCb.py:33-69):
    colour transform + deadzone quantiser, uint8 RGB <-> uint16 indices, integer
    arithmetic, bit-exact.  ``q`` must be integral (the reference's ``-q`` is)."""

    def __init__(self, color: str = "YCoCg", q: int = 32, device: Optional[int] = None):
        if color not in _COLORS:
            raise ValueError(f"color must be one of {list(_COLORS)}")
        if int(q) != q or not (0 < q < 32768):
            raise ValueError("q must be an integer in [1, 32767]")
        self.color, self.q, self.device, self._ctx = _COLORS[color], float(q), device, None

    __del__ = Codec.__del__
    _host_ctx = Codec._host_ctx

    def _run(self, x, encode: bool):
        L = _lib.lib()
        if x.shape[-1] != 3:
            raise ValueError("last dimension must be 3")
        npx = int(np.prod(x.shape[:-1]))
        if _is_torch(x):
            import torch
            want = torch.uint8 if encode else torch.uint16
            if x.dtype != want or not x.is_cuda:
                raise ValueError(f"torch input must be a CUDA {want} tensor")
            x = x.contiguous()
            out = torch.empty(x.shape, dtype=torch.uint16 if encode else torch.uint8, device=x.device)
            with torch.cuda.device(x.device):
                stream = torch.cuda.current_stream().cuda_stream
                fn = L.vcfb_color_encode_dev if encode else L.vcfb_color_decode_dev
                check(fn(x.data_ptr(), npx, self.q, self.color, out.data_ptr(), stream))
            return out
        want = np.uint8 if encode else np.uint16
        if x.dtype != want:
            raise ValueError(f"input must be {np.dtype(want)}")
        x = np.ascontiguousarray(x)
        out = np.empty(x.shape, dtype=np.uint16 if encode else np.uint8)
        fn = L.vcfb_color_encode_host if encode else L.vcfb_color_decode_host
        check(fn(self._host_ctx(), x.ctypes.data, npx, self.q, self.color, out.ctypes.data))
        return out

    def encode(self, rgb):
        """uint8 RGB (...,3) -> uint16 indices (...,3)."""
        return self._run(rgb, True)

    def decode(self, k):
        """uint16 indices (...,3) -> uint8 RGB (...,3)."""
        return self._run(k, False)


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            _lib.lib().vcfb_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array in page-locked host memory (vcfb_host_alloc).  The host entry
    points transfer such arrays in place, without the pageable staging copy."""
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    p = C.c_void_p()
    check(_lib.lib().vcfb_host_alloc(max(nbytes, 1), C.byref(p)))
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    buf._vcfb_owner = _PinnedOwner(p)     # freed when the array (whose base is buf) dies
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)


def encode_frames(rgb, block_size=8, q=32, **kw):
    stats = kw.pop("stats", False)
    return Codec(block_size, q, **kw).encode(rgb, stats=stats)


def decode_frames(idx, shape, block_size=8, q=32, **kw):
    call = {k: kw.pop(k) for k in ("original", "stats", "return_float", "want_rgb", "out") if k in kw}
    return Codec(block_size, q, **kw).decode(idx, shape, **call)


def rd_stats(rgb, block_size=8, q=32, **kw):
    """Encode + decode + statistics in one call: the numbers src/RDE.py reports
    (RMSE) plus the zero-order rate estimate of the index planes."""
    c = Codec(block_size, q, **kw)
    idx, s_enc = c.encode(rgb, stats=True)
    shape = rgb.shape[-3:-1]
    _, s_dec = c.decode(idx, shape, original=rgb, stats=True)
    if _is_torch(rgb):
        return stats_dict((s_enc + s_dec).cpu().numpy())
    merged = np.zeros(STAT_LEN, dtype=np.int64)
    for s in (s_enc, s_dec):
        merged[0:3] += s["sse"]
        merged[3] += s["nsamples"]
        merged[4] += s["nonzero"]
        merged[5] += s["sumabs"]
        merged[6] += s["nindices"]
        merged[7] += s["sumdiff"]
        merged[STAT_HIST:] += s["hist"].ravel()
    return stats_dict(merged)
