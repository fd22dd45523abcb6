"""Rate/distortion sweeps on the GPU.

The reference evaluates one (block size, step) point per process run: ``2D-DCT.py encode``,
``decode``, then ``RDE.py`` reading the files back (src/RDE.py:68-118: RMSE between the
original and decoded images in float32, rate = bytes of the code-stream files, J = bpp +
RMSE), or loops over block sizes inside ``optimize_block_size`` (src/2D-DCT.py:533-579).
Here a frame stays resident on the device and every point is one encode + one decode with
statistics; the distortion is exact (integer SSE), the rate is either the zero-order
entropy of the index planes (an estimate, reported as such) or the size of the real
entropy-coded stream when a ``compress`` callable is given.
"""
from __future__ import annotations

import math
from typing import Callable, Iterable, Optional

import numpy as np

from . import _lib
from .codec import Codec, _COLORS, _is_torch, stats_dict


def rd_stats_fused(frames, block_size: int, qs, color: str = "YCoCg", nowrap: bool = False, hist: bool = True,
                   no_offset: bool = False):
    """Statistics vectors of every step in ``qs`` for one block size from ONE pass over ``frames``
    (vcfb_rd_sweep_dev, csrc/kernels_rd.cu): the forward transform runs once, each step is quantised,
    dequantised and decoded on chip, nothing but the int64 statistics is written.  Row i equals what
    ``Codec(B, qs[i]).encode(stats)`` + ``Codec(B, qs[i], fp64=True).decode(original, stats)`` accumulate.

    nowrap: the dequantiser sees the quantiser's own indices, not the ones wrapped to uint8 -- the
    in-process loop of src/2D-DCT.py:533-579 (optimize_block_size).
    no_offset: neither -128 on the pixels nor +128 on the indices: that loop runs before ``self.offset = 128`` is
    assigned (src/2D-DCT.py:99-110), with the [0, 0, 0] the colour stage left in ``self.offset`` (src/YCoCg.py:28-29).
    frames: CUDA uint8 tensor (n,H,W,3)|(H,W,3) (a numpy array is uploaded).  Returns an int64 tensor
    (len(qs), STAT_LEN) on the device."""
    import ctypes as C
    import torch
    L = _lib.lib()
    if L.vcfb_device_count() < 1:
        raise _lib.VcfbError("no CUDA device: vcf_b200 has no CPU fallback")
    x = frames if _is_torch(frames) else torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    if x.ndim == 3:
        x = x[None]
    if x.ndim != 4 or x.shape[-1] != 3 or x.dtype != torch.uint8 or not x.is_cuda:
        raise ValueError("frames must be uint8 (n,H,W,3) or (H,W,3)")
    x = x.contiguous()
    qs = [float(q) for q in qs]
    out = torch.zeros((len(qs), _lib.STAT_LEN), dtype=torch.int64, device=x.device)
    flags = (_lib.F_HIST if hist else 0) | (_lib.F_NOWRAP if nowrap else 0) | (_lib.F_NO_OFFSET if no_offset else 0)
    n, H, W, _ = x.shape
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream().cuda_stream
        for i in range(0, len(qs), _lib.RD_MAX_STEPS):
            part = qs[i:i + _lib.RD_MAX_STEPS]
            arr = (C.c_double * len(part))(*part)
            _lib.check(L.vcfb_rd_sweep_dev(x.data_ptr(), n, H, W, int(block_size), arr, len(part), _COLORS[color],
                                           flags, out[i:].data_ptr(), stream))
    return out


def _point(B, q, st):
    npx = st["nsamples"] // 3
    res = dict(B=B, q=q, rmse=st["rmse"], psnr=st["psnr"], sse=int(st["sse"].sum()), nonzero=st["nonzero"])
    if "entropy_bits" in st:
        res["bpp_entropy"] = st["entropy_bits"] / npx
    return res


def rd_point(frames, block_size: int, q, compress: Optional[Callable] = None, **codec_kw) -> dict:
    """One point: dict with B, q, rmse, psnr, bpp_entropy (zero-order estimate) and, when
    ``compress(idx_u8_numpy) -> bytes-like/BytesIO`` is given, bytes / bpp / J = bpp + RMSE
    exactly as src/RDE.py:102-117 forms them."""
    enc = Codec(block_size=block_size, q=q, **codec_kw)
    dec = Codec(block_size=block_size, q=q, fp64=True, **{k: v for k, v in codec_kw.items() if k != "contract"})
    idx, s_enc = enc.encode(frames, stats=True)
    shape = frames.shape[-3:-1]
    # (the decoded frames are written although only the statistics are used: the TMA fast paths
    #  always produce them, and they are several times faster than the general kernels that can skip them)
    out = dec.decode(idx, shape, original=frames, stats=True)
    s_dec = out[-1]
    if _is_torch(frames):
        st = stats_dict((s_enc + s_dec).cpu().numpy())
    else:
        from .frames import _dict_to_vec
        st = stats_dict(_dict_to_vec(s_enc) + _dict_to_vec(s_dec))
    npx = st["nsamples"] // 3
    res = dict(B=block_size, q=q, rmse=st["rmse"], psnr=st["psnr"], sse=int(st["sse"].sum()),
               bpp_entropy=st["entropy_bits"] / npx, nonzero=st["nonzero"])
    if compress is not None:
        k = idx.cpu().numpy() if _is_torch(idx) else idx
        k = k.reshape((-1,) + k.shape[-3:])
        nbytes = 0
        for f in k:
            b = compress(f)
            if hasattr(b, "getvalue"):
                b = b.getvalue()
            nbytes += len(b)
        res["bytes"] = nbytes
        res["bpp"] = nbytes * 8 / npx
        res["J"] = res["bpp"] + res["rmse"]                  # src/RDE.py:117
    return res


def rd_sweep(frames, block_sizes: Iterable[int] = (4, 8, 16, 32), qs: Iterable = (4, 8, 12, 16, 24, 32, 48, 64),
             compress: Optional[Callable] = None, fused: bool = True, **codec_kw):
    """BASELINE config 3: every (B, q) point of a frame (or batch) kept on the device.

    fused (default): one pass per block size evaluates all steps (``rd_stats_fused``) -- same numbers as the
    per-point path, which remains for requests the fused kernel does not take (perceptual weights, a real
    entropy coder through ``compress``, float64 / contracted forward transforms)."""
    qs = list(qs)
    plain = compress is None and not any(codec_kw.get(k) for k in ("perceptual", "fp64", "contract", "fast",
                                                                   "disable_subbands", "synth_f32"))
    if not (fused and plain):
        return [rd_point(frames, B, q, compress, **codec_kw) for B in block_sizes for q in qs]
    out = []
    tables = [(B, rd_stats_fused(frames, B, qs, color=codec_kw.get("color", "YCoCg"), hist=codec_kw.get("hist", True)))
              for B in block_sizes]          # all launches first, one synchronising copy at the end
    for B, t in tables:
        t = t.cpu().numpy()
        out += [_point(B, q, stats_dict(t[i])) for i, q in enumerate(qs)]
    return out


def best_block_size(frame, q, Lambda: float, compress: Callable, block_sizes=(4, 8, 16, 32)):
    """arg-min over B of J = rate_bytes + Lambda * RMSE, the objective of
    src/2D-DCT.py:575 (with the distortion between un-shifted images)."""
    best, bestJ = None, math.inf
    for B in block_sizes:
        p = rd_point(frame, B, q, compress)
        J = p["bytes"] + Lambda * p["rmse"]
        if J < bestJ:
            best, bestJ = B, J
    return best, bestJ
