"""Seeded random configurations through the C ABI against the oracle: shapes on and off the fast
paths, every block size, both colour transforms, integral / fractional / large steps, flags,
batches, content from noise to flat.  Everything exact (float32 encode, float64 decode)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vcf_oracle as O


def _content(rng, H, W, kind):
    if kind == "noise":
        return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    if kind == "flat":
        img = np.empty((H, W, 3), dtype=np.uint8)
        img[:] = rng.integers(0, 256, size=3, dtype=np.uint8)
        img[rng.integers(0, H), rng.integers(0, W)] = rng.integers(0, 256, size=3, dtype=np.uint8)   # one odd pixel
        return img
    if kind == "steps":          # piecewise constant: DC-only blocks next to edges
        img = np.repeat(np.repeat(rng.integers(0, 256, size=((H + 23) // 24, (W + 39) // 40, 3), dtype=np.uint8), 24, 0), 40, 1)
        return np.ascontiguousarray(img[:H, :W])
    return O.synthetic_frame(H, W, int(rng.integers(0, 1 << 30)), "natural")


@pytest.mark.parametrize("seed", range(32))
def test_random_configurations_bit_exact(seed):
    import torch
    from vcf_b200 import Codec
    from vcf_b200.codec import stats_dict
    assert torch.cuda.is_available()
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        B = int(rng.choice([4, 8, 8, 16, 16, 32, 32, 2, 64, 128]))
        # widths: multiples of 128 / 256 (fast paths), of 16, and arbitrary
        W = int(rng.choice([128 * rng.integers(1, 5), 256 * rng.integers(1, 4), 16 * rng.integers(1, 20), rng.integers(1, 300)]))
        H = int(rng.choice([B * rng.integers(1, 6), 8 * rng.integers(1, 12), rng.integers(1, 100)]))
        q = rng.choice([1, 2, 3, 5, 8, 12, 16, 31, 32, 64, 100, 255, 300, 2.5, 12.5])
        q = int(q) if float(q).is_integer() else float(q)
        color = "YCrCb" if (B in (8, 16, 32, 64) and rng.random() < 0.4) else "YCoCg"
        kw = {}
        if rng.random() < 0.15:
            kw["disable_subbands"] = True
        if B == 8 and color == "YCoCg" and rng.random() < 0.15:
            kw["perceptual"] = True
        n = int(rng.integers(1, 4))
        frames = np.stack([_content(rng, H, W, rng.choice(["noise", "natural", "flat", "steps"])) for _ in range(n)])
        okw = dict(color=color, **kw)
        ref = np.stack([O.encode_array(f, B, q, **okw) for f in frames])
        refd = np.stack([O.decode_array(k, (H, W, 3), B, q, **okw) for k in ref])
        x = torch.from_numpy(frames).cuda()
        tag = (seed, B, H, W, q, color, kw, n)
        idx, st = Codec(block_size=B, q=q, color=color, hist=bool(rng.random() < 0.5), **kw).encode(x, stats=True)
        assert np.array_equal(idx.cpu().numpy(), ref), tag
        nz, sabs, _ = O.index_stats(ref)
        s = stats_dict(st.cpu().numpy())
        assert s["nonzero"] == nz and s["sumabs"] == sabs and s["nindices"] == ref.size, tag
        y, sd = Codec(block_size=B, q=q, color=color, fp64=True, **kw).decode(idx, (H, W), original=x, stats=True)
        assert np.array_equal(y.cpu().numpy(), refd), tag
        s = stats_dict(sd.cpu().numpy())
        assert [int(v) for v in s["sse"]] == [O.sse_int(frames[..., c], refd[..., c]) for c in range(3)], tag
        assert s["nsamples"] == frames.size, tag
        y2 = Codec(block_size=B, q=q, color=color, fp64=True, **kw).decode(idx, (H, W))
        assert np.array_equal(y2.cpu().numpy(), refd), tag
        if not kw:
            # the float32 decoders (tensor-core / B=16 / B=32 fast modes or the general kernel): +-1 LSB
            y32 = Codec(block_size=B, q=q, color=color).decode(idx, (H, W)).cpu().numpy()
            assert np.abs(y32.astype(np.int16) - refd.astype(np.int16)).max() <= 1, tag
            # the fused sweep: the statistics of this step and of a second one from one pass
            from vcf_b200.rd import rd_stats_fused
            q2 = 7 if q != 7 else 9
            tab = rd_stats_fused(x, B, (q, q2), color=color).cpu().numpy()
            s1 = stats_dict(tab[0])
            assert s1["nonzero"] == nz and s1["sumabs"] == sabs and s1["nindices"] == ref.size, tag
            assert [int(v) for v in s1["sse"]] == [O.sse_int(frames[..., c], refd[..., c]) for c in range(3)], tag
            ref2 = np.stack([O.encode_array(f, B, q2, color=color) for f in frames])
            refd2 = np.stack([O.decode_array(k, (H, W, 3), B, q2, color=color) for k in ref2])
            s2 = stats_dict(tab[1])
            assert [int(v) for v in s2["sse"]] == [O.sse_int(frames[..., c], refd2[..., c]) for c in range(3)], tag
            assert s2["nonzero"] == O.index_stats(ref2)[0], tag
