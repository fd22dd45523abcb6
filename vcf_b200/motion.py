"""Motion estimation of the reference's hybrid codec on the GPU.

``block_matching(ref_frame, curr_frame, block_size, search_range)`` has the signature and the
result of ``IPP.block_matching`` (/root/reference/src/IPP_DCT.py:344-373) with ``use_fast=False``:
RGB frames are converted to gray like ``cv2.cvtColor(..., COLOR_RGB2GRAY)`` (:350-352), every
block is matched by full search (:217-244) and the field comes back as float32
``(h // bs, w // bs, 2)`` holding ``(dx, dy)``.  Integer arithmetic in libvcfb200.so
(``vcfb_gray_dev``, ``vcfb_block_match_dev``); no CPU fallback.

numpy in -> numpy out (synchronous); torch CUDA tensors in -> torch CUDA tensor out (asynchronous
on torch's current stream).  A leading batch dimension matches pair ``f`` of ``ref`` with pair
``f`` of ``curr``.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _gray_dev(torch, x):
    """(..., H, W, 3) or (..., H, W) uint8 CUDA tensor -> (..., H, W) gray."""
    if x.dtype != torch.uint8:
        raise ValueError("frames must be uint8")
    x = x.contiguous()
    if x.shape[-1] != 3 or x.dim() < 3:
        return x
    out = torch.empty(x.shape[:-1], dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().vcfb_gray_dev(x.data_ptr(), out.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


def rgb_to_gray(frame):
    """cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY) for uint8 RGB (..., H, W, 3)."""
    import torch
    if _is_torch(frame):
        return _gray_dev(torch, frame)
    if not torch.cuda.is_available():
        raise _lib.VcfbError("no CUDA device: vcf_b200 has no CPU fallback")
    return _gray_dev(torch, torch.from_numpy(np.ascontiguousarray(frame)).cuda()).cpu().numpy()


def block_matching(ref_frame, curr_frame, block_size: int = 16, search_range: int = 8, color: bool | None = None,
                   use_fast: bool = False):
    """Motion vectors of ``curr_frame`` against ``ref_frame``: full search, or with ``use_fast``
    the three-step search of ``_three_step_search`` (src/IPP_DCT.py:159-205, the ``--fast`` flag).

    Frames: uint8, gray ``(H, W)`` / ``(n, H, W)`` or RGB ``(H, W, 3)`` / ``(n, H, W, 3)``
    (``color`` overrides the guess "last dimension == 3 means RGB")."""
    import torch
    as_numpy = not _is_torch(ref_frame)
    if as_numpy:
        if not torch.cuda.is_available():
            raise _lib.VcfbError("no CUDA device: vcf_b200 has no CPU fallback")
        ref_frame = torch.from_numpy(np.ascontiguousarray(ref_frame)).cuda()
        curr_frame = torch.from_numpy(np.ascontiguousarray(curr_frame)).cuda()
    if ref_frame.shape != curr_frame.shape:
        raise ValueError("reference and current frame differ in shape")
    is_rgb = (ref_frame.dim() >= 3 and ref_frame.shape[-1] == 3) if color is None else bool(color)
    if is_rgb:
        ref_frame, curr_frame = _gray_dev(torch, ref_frame), _gray_dev(torch, curr_frame)
    if ref_frame.dtype != torch.uint8 or curr_frame.dtype != torch.uint8:
        raise ValueError("frames must be uint8")
    batched = ref_frame.dim() == 3
    if ref_frame.dim() not in (2, 3):
        raise ValueError(f"unexpected frame shape {tuple(ref_frame.shape)}")
    r = ref_frame.contiguous()
    c = curr_frame.contiguous()
    n = r.shape[0] if batched else 1
    H, W = r.shape[-2], r.shape[-1]
    bs, sr = int(block_size), int(search_range)
    mv = torch.empty((n, H // bs if H >= bs else 0, W // bs if W >= bs else 0, 2), dtype=torch.int16, device=r.device)
    with torch.cuda.device(r.device):
        fn = _lib.lib().vcfb_block_match_tss_dev if use_fast else _lib.lib().vcfb_block_match_dev
        check(fn(r.data_ptr(), c.data_ptr(), n, H, W, bs, sr, mv.data_ptr(), torch.cuda.current_stream().cuda_stream))
    mv = mv.to(torch.float32)                      # the reference's field is float32 (:354)
    if not batched:
        mv = mv[0]
    return mv.cpu().numpy() if as_numpy else mv
