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

from .codec import Codec, _is_torch, stats_dict


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
             compress: Optional[Callable] = None, **codec_kw):
    """BASELINE config 3: every (B, q) point of a frame (or batch) kept on the device."""
    return [rd_point(frames, B, q, compress, **codec_kw) for B in block_sizes for q in qs]


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
