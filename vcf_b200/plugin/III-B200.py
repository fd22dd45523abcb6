'''Intra-only sequence coding: every frame goes through the 2D codec, in pipelined GPU chunks.'''

# Batched counterpart of the reference's src/III.py.  Same flags (-T transform, -N
# number_of_frames, src/III.py:23-32), same file names (/tmp/original_%04d.png,
# /tmp/encoded_%04d<ext> + _shape.bin, /tmp/decoded_%04d.png, :85-86, :133-134) and the same
# per-frame results as calling the transform's encode_fn / decode_fn once per frame (what
# :96-104 intends and :132-144 does) -- but the frames travel in chunks through a fixed-depth
# ring of pinned buffers (vcf_b200/pipeline.py): while one chunk is on the GPU the previous one
# is entropy-coded / written and the next one is read by a pool of host threads.  Under
# torchrun the sequence is sharded by contiguous frame ranges, one rank per GPU (vcf_b200.frames).
# The entropy coder and the file formats are the chain's own (compress / decompress /
# encode_write_fn / decode_write_fn of the -c codec), so either side can be the reference.

import importlib
import logging
import os
import struct
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(_here))
for _p in (os.getcwd(), _here, _repo):
    if _p not in sys.path:
        sys.path.append(_p)

import main  # noqa: E402
with open("/tmp/description.txt", 'w') as f:
    f.write(__doc__)
import parser  # noqa: E402

from vcf_b200.frames import frame_range  # noqa: E402
from vcf_b200.pipeline import ChunkPipeline  # noqa: E402

DEFAULT_TRANSFORM = "2D-DCT-B200"
N_FRAMES = 20                      # src/video_coding.py:29
ORIGINAL_PREFIX = "/tmp/original"  # frames extracted by src/III.py:85
ENCODE_OUTPUT_PREFIX = "/tmp/encoded"
DECODE_OUTPUT_PREFIX = "/tmp/decoded"

for _p in (parser.parser_encode, parser.parser_decode):
    _p.add_argument("-T", "--transform", type=str, help=f"module of the 2D codec (default {DEFAULT_TRANSFORM})", default=DEFAULT_TRANSFORM)
    _p.add_argument("-N", "--number_of_frames", type=parser.int_or_str, help=f"how many frames of the sequence to code (default {N_FRAMES})", default=N_FRAMES)
    _p.add_argument("--io_threads", type=int, default=8, help="host threads for entropy coding and file IO")
    _p.add_argument("--chunk_frames", type=int, default=8, help="frames per GPU chunk of the pipeline")
    _p.add_argument("--pipeline_depth", type=int, default=3, help="chunks in flight: one being read, one on the GPU, one being written")

args = parser.parser.parse_known_args()[0]
transform = importlib.import_module(args.transform)


def _rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _local_device():
    """One process per GPU under torchrun: rank r of a node works on GPU LOCAL_RANK."""
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        local %= max(1, torch.cuda.device_count())
        torch.cuda.set_device(local)
    return local


def _wait_for(paths, timeout_s=600.0, poll_s=0.05):
    """Ranks other than 0 wait until rank 0 has written the extracted frames (a marker file
    written after the last frame closes the race with a half-written PNG)."""
    import time
    t0 = time.time()
    while not all(os.path.exists(p) for p in paths):
        if time.time() - t0 > timeout_s:
            raise TimeoutError(f"frames not extracted after {timeout_s:.0f} s: {paths[-1]}")
        time.sleep(poll_s)


class CoDec:

    def __init__(self, args):
        logging.debug("trace")
        self.args = args
        self.device = _local_device()
        self.transform_codec = transform.CoDec(args)
        if not hasattr(self.transform_codec, "_codec"):
            raise TypeError("III-B200 batches through the GPU transform; use -T 2D-DCT-B200")
        self.transform_codec.device = self.device       # Codec objects of the transform stage use this GPU
        logging.info(f"Using {args.transform} codec")

    def bye(self):
        pass

    def _extract_frames(self, n):
        """src/III.py:73-115: demux args.original with PyAV into /tmp/original_%04d.png --
        only when the frames are not there yet and PyAV is installed."""
        if all(os.path.exists(f"{ORIGINAL_PREFIX}_%04d.png" % i) for i in range(n)):
            return
        import av  # not a dependency of the GPU path
        import cv2
        container = av.open(self.args.original)
        i = 0
        for frame in container.decode(video=0):
            img = np.array(frame.to_image().convert("RGB"))
            cv2.imwrite(f"{ORIGINAL_PREFIX}_%04d.png" % i, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
            i += 1
            if i >= n:
                break

    def encode(self):
        tc = self.transform_codec
        n = int(self.args.number_of_frames)
        rank, world = _rank_world()
        if world > 1:
            # only rank 0 extracts; the others wait for a marker that is unique to this launch
            run = os.environ.get("TORCHELASTIC_RUN_ID", "") + "." + os.environ.get("MASTER_PORT", "")
            marker = f"{ORIGINAL_PREFIX}.extracted.{run}.{n}"
            if rank == 0:
                self._extract_frames(n)
                open(marker, "w").close()
            else:
                _wait_for([marker])
        else:
            self._extract_frames(n)
        lo, hi = frame_range(n, rank, world)
        if hi == lo:
            return 0
        codec = tc._codec()
        gpu_entropy = bool(getattr(tc, "accepts_device_arrays", False))
        shapes = {}

        def read(i):
            fr = tc.encode_read_fn(f"{ORIGINAL_PREFIX}_%04d.png" % i)
            tc._check_image(fr)
            shapes[i] = fr.shape
            return fr

        def finish(i, k, on_device):
            out_fn = f"{ENCODE_OUTPUT_PREFIX}_%04d" % i
            with open(f"{out_fn}_shape.bin", "wb") as file:           # src/2D-DCT.py:285-286
                file.write(struct.pack("iii", *shapes[i]))
            # the entropy stage of the chain (host zlib / TIFF), or the GPU one fed from HBM (-c z_lib-B200 / TIFF-B200)
            return tc.encode_write_fn(tc.compress(k if on_device else np.array(k)), out_fn)

        pipe = ChunkPipeline(device=self.device, depth=self.args.pipeline_depth, chunk=self.args.chunk_frames,
                             io_threads=self.args.io_threads, keep_on_device=gpu_entropy)
        sizes = pipe.run(hi - lo, read, lambda x, m: codec.encode(x), finish, first=lo)
        logging.info(f"rank {rank}: frames [{lo},{hi}) -> {sum(sizes)} bytes")
        return sum(sizes)

    def decode(self):
        tc = self.transform_codec
        n = int(self.args.number_of_frames)
        rank, world = _rank_world()
        lo, hi = frame_range(n, rank, world)
        if hi == lo:
            return 0
        if getattr(self.args, "filter", "no_filter") != "no_filter":
            # a real post-filter needs the un-clipped float image of every frame (src/2D-DCT.py:454-466):
            # the transform stage's own decode_fn does exactly that
            return sum(tc.decode_fn(f"{ENCODE_OUTPUT_PREFIX}_%04d" % i, f"{DECODE_OUTPUT_PREFIX}_%04d.png" % i)
                       for i in range(lo, hi))
        codec = tc._codec(decode=True)
        shapes = {}

        def read(i):
            in_fn = f"{ENCODE_OUTPUT_PREFIX}_%04d" % i
            with open(f"{in_fn}_shape.bin", "rb") as file:
                shapes[i] = struct.unpack("iii", file.read(12))
            k = np.ascontiguousarray(tc.decompress(tc.decode_read_fn(in_fn)))
            if k.dtype != np.uint8:
                raise ValueError(f"code-stream holds {k.dtype}, expected uint8 indices")
            return k

        def gpu(k, m):
            shape = shapes[lo]
            if len({shapes[i] for i in shapes}) != 1:
                raise ValueError("all frames of a sequence must have the same shape")
            return codec.decode(k, shape[:2])

        def finish(i, y, on_device):
            return tc.decode_write_fn(np.array(y), f"{DECODE_OUTPUT_PREFIX}_%04d.png" % i)

        pipe = ChunkPipeline(device=self.device, depth=self.args.pipeline_depth, chunk=self.args.chunk_frames,
                             io_threads=self.args.io_threads)
        return sum(pipe.run(hi - lo, read, gpu, finish, first=lo))


if __name__ == "__main__":
    main.main(parser.parser, logging, CoDec)
