"""Host-side multi-GPU logic on CPU: frame sharding and the statistics all-reduce over
a world_size-2 gloo group (the NCCL path is the same calls on CUDA tensors)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_ranges_tile_the_sequence():
    from vcf_b200.frames import frame_range
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            r = [frame_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert frame_range(1024, 3, 8) == (384, 512)       # 128 frames per GPU
    with pytest.raises(ValueError):
        frame_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from oracle import vcf_oracle as O
    from vcf_b200 import _lib
    from vcf_b200.frames import FrameParallel, allreduce_stats, frame_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1) plain all-reduce of a statistics vector
        v = np.arange(_lib.STAT_LEN, dtype=np.int64) * (rank + 1)
        tot = allreduce_stats(v.copy())
        assert np.array_equal(tot, np.arange(_lib.STAT_LEN, dtype=np.int64) * sum(range(1, world + 1)))
        # 1b) the asynchronous form used by batch loops: several reductions in flight, waited for later
        import torch
        pend = []
        for b in range(3):
            t = torch.arange(_lib.STAT_LEN, dtype=torch.int64) * (rank + 1 + b)
            pend.append((b, *allreduce_stats(t, async_op=True)))
        for b, t, work in pend:
            assert work is not None
            work.wait()
            assert torch.equal(t, torch.arange(_lib.STAT_LEN, dtype=torch.int64) * sum(r + 1 + b for r in range(world)))
        # 2) sharded round-trip statistics with an oracle-backed stand-in for the GPU
        #    codec (test double only: the product class has no CPU path)
        n, H, W, B, qq = 5, 32, 48, 8, 16
        frames = np.stack([O.synthetic_frame(H, W, 40 + i, "natural") for i in range(n)])

        class OracleCodec:
            def encode(self, x, stats=False):
                idx = np.stack([O.encode_array(f, B, qq) for f in x])
                nz, sabs, hist = O.index_stats(idx)
                return idx, dict(sse=np.zeros(3, np.int64), nsamples=0, nonzero=nz, sumabs=sabs,
                                 nindices=idx.size, hist=hist)

            def decode(self, idx, shape, original=None, stats=False):
                y = np.stack([O.decode_array(k, (shape[0], shape[1], 3), B, qq) for k in idx])
                sse = np.array([O.sse_int(original[..., c], y[..., c]) for c in range(3)], np.int64)
                return y, dict(sse=sse, nsamples=y.size, nonzero=0, sumabs=0, nindices=0,
                               hist=np.zeros((3, 256), np.int64))

        fp = FrameParallel(OracleCodec())
        lo, hi = fp.my_range(n)
        assert (lo, hi) == frame_range(n, rank, world)
        st = fp.round_trip_stats(frames[lo:hi])
        # reference: the whole sequence on one rank
        idx = np.stack([O.encode_array(f, B, qq) for f in frames])
        y = np.stack([O.decode_array(k, (H, W, 3), B, qq) for k in idx])
        assert int(st["sse"].sum()) == O.sse_int(frames, y)
        assert st["nsamples"] == frames.size and st["nindices"] == idx.size
        nz, sabs, hist = O.index_stats(idx)
        assert st["nonzero"] == nz and st["sumabs"] == sabs and np.array_equal(st["hist"], hist)
        assert abs(st["rmse"] - float(O.rmse(frames, y))) < 1e-4
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
