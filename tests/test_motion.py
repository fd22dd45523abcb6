"""Row F3 (motion estimation of the hybrid codec): oracle against the vectors recorded from the
reference's own `_process_block_row`, and the GPU kernels against both."""
import glob
import os

import numpy as np
import pytest

from oracle import me_oracle as M

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_me_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 6


@pytest.mark.parametrize("fn", GOLD, ids=[os.path.basename(f)[7:-4] for f in GOLD])
def test_oracle_matches_reference_vectors(fn):
    g = np.load(fn)
    mv = M.block_matching_full(g["ref"], g["cur"], int(g["bs"]), int(g["sr"]))
    assert mv.dtype == np.float32 and np.array_equal(mv, g["mv"])
    mvf = M.block_matching_tss(g["ref"], g["cur"], int(g["bs"]), int(g["sr"]))
    assert mvf.dtype == np.float32 and np.array_equal(mvf, g["mv_fast"])


def test_gray_fixed_point_is_opencv():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(512, 1024, 3), dtype=np.uint8)
    assert np.array_equal(M.gray_fixed_point(x), M.gray_from_rgb(x))
    # every (R, G) pair with a few B, and every (G, B) pair with a few R
    a = np.arange(256, dtype=np.uint8)
    for k in (0, 1, 127, 128, 254, 255):
        y = np.stack(list(np.meshgrid(a, a, indexing="ij")) + [np.full((256, 256), k, np.uint8)], -1)
        assert np.array_equal(M.gray_fixed_point(y), M.gray_from_rgb(np.ascontiguousarray(y)))
        z = np.ascontiguousarray(y[..., ::-1])
        assert np.array_equal(M.gray_fixed_point(z), M.gray_from_rgb(z))


@pytest.mark.gpu
@pytest.mark.parametrize("fn", GOLD, ids=[os.path.basename(f)[7:-4] for f in GOLD])
def test_gpu_block_matching_matches_reference_vectors(fn):
    from vcf_b200 import _lib
    from vcf_b200.motion import block_matching
    g = np.load(fn)
    mv = block_matching(g["ref"], g["cur"], int(g["bs"]), int(g["sr"]))
    assert _lib.last_kernel() == "block_match"
    assert mv.dtype == np.float32 and np.array_equal(mv, g["mv"])
    mvf = block_matching(g["ref"], g["cur"], int(g["bs"]), int(g["sr"]), use_fast=True)
    assert _lib.last_kernel() == "block_match_tss"
    assert np.array_equal(mvf, g["mv_fast"])


@pytest.mark.gpu
def test_gpu_block_matching_against_oracle():
    import torch
    from vcf_b200.motion import block_matching, rgb_to_gray
    rng = np.random.default_rng(5)
    for (h, w, bs, sr) in ((96, 160, 16, 8), (100, 130, 16, 8), (64, 64, 8, 16), (128, 96, 32, 4), (48, 48, 4, 0),
                           (270, 480, 16, 8)):
        base = rng.integers(0, 256, size=(h + 40, w + 40, 3), dtype=np.uint8)
        ref = np.ascontiguousarray(base[20:20 + h, 20:20 + w])
        cur = np.ascontiguousarray(base[17:17 + h, 24:24 + w])
        cur[: h // 2] = ref[: h // 2] // 2              # an area that does not match anything well
        cur[:, : bs] = 50                               # flat column: ties
        rg, cg = M.gray_from_rgb(ref), M.gray_from_rgb(cur)
        assert np.array_equal(rgb_to_gray(ref), rg)
        want = M.block_matching_full(rg, cg, bs, sr)
        got = block_matching(ref, cur, bs, sr)          # RGB in: gray conversion on the GPU
        assert np.array_equal(got, want), (h, w, bs, sr)
        if h * w <= 16384:                              # the loop-form oracle of the three-step search is slow
            assert np.array_equal(block_matching(rg, cg, bs, sr, use_fast=True), M.block_matching_tss(rg, cg, bs, sr)), (h, w, bs, sr)
        got2 = block_matching(torch.from_numpy(np.stack([rg, cg])).cuda(), torch.from_numpy(np.stack([cg, rg])).cuda(), bs, sr)
        assert np.array_equal(got2[0].cpu().numpy(), want)
        assert np.array_equal(got2[1].cpu().numpy(), M.block_matching_full(cg, rg, bs, sr))


@pytest.mark.gpu
def test_gpu_block_matching_errors():
    from vcf_b200 import VcfbError
    from vcf_b200.motion import block_matching
    x = np.zeros((32, 32), dtype=np.uint8)
    with pytest.raises(VcfbError):
        block_matching(x, x, 64, 8)          # frame smaller than a block
    with pytest.raises(VcfbError):
        block_matching(x, x, 16, 40)         # search range out of bounds
