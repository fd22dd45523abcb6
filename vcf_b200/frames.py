"""Frame-parallel sharding and rate/distortion statistics across GPUs.

Frames of an intra-coded sequence carry no state from one to the next
(/root/reference/src/III.py:132-144 calls ``decode_fn`` once per frame; the intended
encode loop is :96-104), so the path shards by contiguous frame ranges, one process per
GPU, with no data-path collective.  The only exchange is one all-reduce (sum) of the
int64 statistics vector (include/vcfb200.h VCFB_STAT_*) that ``RDE.py``-style reports
are computed from; integer sums make the result independent of the reduction order.
"""
from __future__ import annotations

import numpy as np

from ._lib import STAT_LEN
from .codec import stats_dict


def frame_range(n_frames: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank ``rank``: sizes differ by at most one and the
    ranges tile [0, n_frames) in rank order."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad rank / world / n_frames")
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(vec, group=None, async_op: bool = False):
    """Sum the statistics vector over all ranks (NCCL for CUDA tensors, gloo for CPU
    tensors / numpy).  Returns the same kind of object it was given.

    async_op=True (CUDA tensors): returns ``(tensor, work)``; the reduction runs on NCCL's own
    stream beside the kernels launched afterwards, and ``work.wait()`` (``None`` on a single
    rank) makes the current stream wait for it -- a batch loop waits for batch i while batch
    i+1 is already being transformed, which takes the collective off the critical path."""
    import torch
    import torch.distributed as dist
    is_np = isinstance(vec, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.int64)) if is_np else vec
    if t.numel() != STAT_LEN or t.dtype != torch.int64:
        raise ValueError("statistics vector must be int64[%d]" % STAT_LEN)
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if async_op:
        return (t.numpy() if is_np else t), work
    return t.numpy() if is_np else t


class FrameParallel:
    """Runs a :class:`vcf_b200.Codec` on this rank's share of a frame sequence.

    ``codec_enc`` / ``codec_dec`` may be the same object.  ``rank`` / ``world`` default
    to the initialised torch.distributed group (or a single rank)."""

    def __init__(self, codec_enc, codec_dec=None, rank=None, world=None, group=None):
        import torch.distributed as dist
        self.enc = codec_enc
        self.dec = codec_dec if codec_dec is not None else codec_enc
        self.group = group
        if rank is None or world is None:
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world = 0, 1
        self.rank, self.world = rank, world

    def my_range(self, n_frames: int):
        return frame_range(n_frames, self.rank, self.world)

    def encode(self, frames, stats: bool = False):
        """``frames``: this rank's frames (n_local,H,W,3)."""
        return self.enc.encode(frames, stats=stats)

    def round_trip_stats(self, frames):
        """Encode + decode this rank's frames and return the *global* statistics
        (dict, see codec.stats_dict): SSE/RMSE/PSNR against the originals plus the
        index histogram / zero-order rate estimate, summed over all ranks."""
        idx, s_enc = self.enc.encode(frames, stats=True)
        shape = frames.shape[-3:-1]
        out = self.dec.decode(idx, shape, original=frames, stats=True)
        s_dec = out[-1]
        if isinstance(s_enc, dict):           # numpy path returns dicts
            vec = _dict_to_vec(s_enc) + _dict_to_vec(s_dec)
            vec = allreduce_stats(vec, self.group)
            return stats_dict(vec)
        vec = allreduce_stats(s_enc + s_dec, self.group)
        return stats_dict(vec.cpu().numpy())


def _dict_to_vec(d: dict) -> np.ndarray:
    v = np.zeros(STAT_LEN, dtype=np.int64)
    v[0:3] = d["sse"]
    v[3] = d["nsamples"]
    v[4] = d["nonzero"]
    v[5] = d["sumabs"]
    v[6] = d["nindices"]
    v[7] = d.get("sumdiff", 0)
    v[8:8 + 768] = d["hist"].ravel()
    return v
