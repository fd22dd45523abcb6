"""CPU restatement of the motion-estimation step of the reference's hybrid codec (SURVEY.md 8f
row F3).  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product (vcf_b200/) never does.

Follows /root/reference/src/IPP_DCT.py:
  * gray_from_rgb       :350-352  cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY) -- the real OpenCV;
    gray_fixed_point restates its 8-bit arithmetic and is pinned against it in tests/test_motion.py
  * block_matching_full :217-244 (`_process_block_row`, use_fast=False) and :344-373
    (`IPP.block_matching`): dy-major scan of [-sr, sr]^2, out-of-frame candidates skipped,
    SAD on int16 differences, strict `<` so the first minimum of the scan wins; the field is
    float32 (h//bs, w//bs, 2) holding (dx, dy).
  * block_matching_tss  :159-205 (`_three_step_search`, use_fast=True) driven by :212-215: the
    centre moves inside the neighbour loop, and the step stays at 1 while rounds keep improving.
Parity status: pinned -- tests/golden/ref_me_*.npz were produced by executing the reference's own
`_process_block_row` (oracle/make_golden_me.py extracts it from the unmodified source file).
"""
from __future__ import annotations

import numpy as np


def gray_from_rgb(frame: np.ndarray) -> np.ndarray:
    import cv2
    return cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY)


def gray_fixed_point(frame: np.ndarray) -> np.ndarray:
    """OpenCV's 8-bit RGB2GRAY: (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (OpenCV 4.x)."""
    f = frame.astype(np.uint32)
    return ((f[..., 0] * 9798 + f[..., 1] * 19235 + f[..., 2] * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def block_matching_full(ref_gray: np.ndarray, cur_gray: np.ndarray, bs: int, sr: int) -> np.ndarray:
    """All blocks at once, one candidate displacement at a time, in the reference's scan order."""
    h, w = ref_gray.shape
    ny, nx = h // bs, w // bs                       # rows: range(0, h-bs+1, bs)  (:357-359)
    ref = ref_gray.astype(np.int16)
    cur = cur_gray.astype(np.int16)[:ny * bs, :nx * bs].reshape(ny, bs, nx, bs)
    best = np.full((ny, nx), np.iinfo(np.int64).max, dtype=np.int64)       # min_sad = inf (:223)
    mv = np.zeros((ny, nx, 2), dtype=np.float32)                           # best_mv = (0, 0)
    ii = np.arange(ny)[:, None] * bs
    jj = np.arange(nx)[None, :] * bs
    for dy in range(-sr, sr + 1):                                          # :227
        for dx in range(-sr, sr + 1):                                      # :232
            ok = (ii + dy >= 0) & (ii + dy + bs <= h) & (jj + dx >= 0) & (jj + dx + bs <= w)   # :229,:234
            if not ok.any():
                continue
            # shifted reference, zero where it would leave the frame (those blocks are masked by `ok`)
            sh = np.zeros((ny * bs, nx * bs), dtype=np.int16)
            y0, y1 = max(0, -dy), min(ny * bs, h - dy)
            x0, x1 = max(0, -dx), min(nx * bs, w - dx)
            sh[y0:y1, x0:x1] = ref[y0 + dy:y1 + dy, x0 + dx:x1 + dx]
            sad = np.abs(cur - sh.reshape(ny, bs, nx, bs)).sum(axis=(1, 3), dtype=np.int64)    # :238
            take = ok & (sad < best)                                       # :240
            best[take] = sad[take]
            mv[take] = (dx, dy)                                            # :242
    return mv


def block_matching_tss(ref_gray: np.ndarray, cur_gray: np.ndarray, bs: int, sr: int) -> np.ndarray:
    """Three-step search, block by block (plain loops: the walk is data dependent)."""
    h, w = ref_gray.shape
    ny, nx = h // bs, w // bs
    ref = ref_gray.astype(np.int32)
    cur = cur_gray.astype(np.int32)
    mv = np.zeros((ny, nx, 2), dtype=np.float32)

    def sad(block, y, x):
        return int(np.abs(block - ref[y:y + bs, x:x + bs]).sum())

    for by in range(ny):
        for bx in range(nx):
            i, j = by * bs, bx * bs
            block = cur[i:i + bs, j:j + bs]
            step = sr // 2                                        # :166
            cx, cy, best = j, i, (0, 0)
            min_sad = sad(block, cy, cx)                          # :171-175 (always inside the frame)
            while step >= 1:                                      # :177
                improved = False
                for dy in (-step, 0, step):                       # :180-181
                    for dx in (-step, 0, step):
                        if dy == 0 and dx == 0:
                            continue
                        ry, rx = cy + dy, cx + dx                 # :185-186: the centre may already have moved
                        if ry < 0 or ry + bs > h or rx < 0 or rx + bs > w:
                            continue
                        s_ = sad(block, ry, rx)
                        if s_ < min_sad:                          # :193
                            min_sad, best = s_, (rx - j, ry - i)
                            cx, cy = rx, ry
                            improved = True
                step = max(1, step // 2) if improved else step // 2   # :199-202
            mv[by, bx] = best
    return mv
