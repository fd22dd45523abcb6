'''Block-DCT spatial stage (fixed block size) with the colour / DCT / deadzone arithmetic on a B200 GPU.'''

# Drop-in replacement of the reference's src/2D-DCT.py: same flags and dests
# (src/2D-DCT.py:36-45), same class chain (``class CoDec(CT.CoDec)``, CT chosen by
# -t at import time, :47-56), same entry points (``encode_fn/decode_fn(in_fn, out_fn)``,
# ``encode()/decode()``, :268, :374, :377, :470) and the same side files
# (``<out>_shape.bin`` :285-286, the code-stream written by the entropy stage).  Only
# the arithmetic between ``encode_read_fn`` and ``compress`` (:276-361) and between
# ``decompress`` and ``decode_write_fn`` (:398-466) is replaced: it runs in
# libvcfb200.so on the GPU (vcf_b200.Codec) instead of the DCT2D / color_transforms /
# scalar_quantization packages.  Use it like the original:
#
#     python 2D-DCT-B200.py encode -B 8 -q 32        (from the reference's src/)
#     python III.py decode -T 2D-DCT-B200            (src/III.py:41 imports it by name)
#     python IPP_DCT.py encode --st 2D-DCT-B200      (src/IPP_DCT.py:52-78)
#
# See INTEGRATION.md for the search-path setup.

import importlib
import logging
import os
import struct
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(_here))
for _p in (os.getcwd(), _repo):            # the reference's src/ (main, parser, CT chain) and vcf_b200
    if _p not in sys.path:
        sys.path.append(_p)

import main  # noqa: E402  (reference src/main.py)
with open("/tmp/description.txt", 'w') as f:   # handshake read by src/parser.py:67
    f.write(__doc__)
import parser  # noqa: E402  (reference src/parser.py)

from vcf_b200 import Codec  # noqa: E402
from vcf_b200.codec import stats_dict  # noqa: E402
from vcf_b200.rd import rd_stats_fused  # noqa: E402

default_block_size = 8
default_CT = "YCoCg"
perceptual_quantization = False
disable_subbands = False
SUPPORTED_B = (2, 4, 8, 16, 32, 64, 128)     # src/2D-DCT.py:538: 2**i, i = 1..7

for _p in (parser.parser_encode, parser.parser_decode):
    _p.add_argument("-B", "--block_size_DCT", type=parser.int_or_str, help=f"side of the square DCT blocks: a power of two in [2, 128] (default {default_block_size})", default=default_block_size)
    _p.add_argument("-t", "--color_transform", type=parser.int_or_str, help=f"module providing the colour stage / base class (default {default_CT})", default=default_CT)
    _p.add_argument("-p", "--perceptual_quantization", action='store_true', help="weight the coefficients with the JPEG luma / chroma tables before quantising", default=perceptual_quantization)
    _p.add_argument("-x", "--disable_subbands", action='store_true', help="keep the coefficients in block order instead of grouping them by subband", default=disable_subbands)
parser.parser_encode.add_argument("-L", "--Lambda", type=parser.int_or_str, help="when given (float): pick the block size in {4, 8, 16, 32} minimising bytes + Lambda * RMSE")
parser.parser_decode.add_argument("--b200_fast_decode", action='store_true', help="float32 GPU decoder (pixels within +-1 of the reference, PSNR within 0.01 dB) instead of the bit-exact float64 one", default=False)

parser.parser_encode.add_argument("--b200_fast_encode", action='store_true', help="tensor-core GPU encoder (B = 8, q a power of two >= 8): fewer than 1e-6 of the indices differ from the reference, only at rounding boundaries", default=False)
parser.parser_decode.add_argument("--b200_synth_f32", action='store_true', help="float64 GPU decoder in the upstream variant that stores the synthesised image as float32 (see include/vcfb200.h VCFB_F_SYNTH_F32)", default=False)

args = parser.parser.parse_known_args()[0]
CT = importlib.import_module(args.color_transform)


class CoDec(CT.CoDec):

    def __init__(self, args):
        logging.debug("trace")
        super().__init__(args)
        self.block_size = args.block_size_DCT
        logging.debug(f"block_size = {self.block_size}")
        self.perceptual = bool(args.perceptual_quantization)
        self.disable_subbands = bool(args.disable_subbands)
        self._codecs = {}
        if args.quantizer != "deadzone":
            # src/2D-DCT.py:107-110 sets offset 0 for other quantizers; only the default
            # stack (deadzone) is on the GPU path.
            raise NotImplementedError("2D-DCT-B200 implements the default stack only (-a deadzone)")
        self.offset = 128
        if self.encoding and getattr(args, "Lambda", None) is not None:
            if not args.perceptual_quantization:
                self.Lambda = float(args.Lambda)
                logging.info("optimizing the block size")
                self.optimize_block_size()
                logging.info(f"optimal block_size={self.block_size}")
            else:
                logging.warning("sorry, perceptual quantization is only available for block_size=8")

    # -- GPU codec objects, one per (block size, direction) -------------------------
    def _codec(self, block_size=None, decode=False):
        B = int(block_size if block_size is not None else self.block_size)
        if B not in SUPPORTED_B:
            raise ValueError(f"block size {B} is not supported by the GPU path (supported: {SUPPORTED_B})")
        fp64 = decode and not getattr(self.args, "b200_fast_decode", False)
        synth32 = fp64 and bool(getattr(self.args, "b200_synth_f32", False))
        fast = (not decode) and bool(getattr(self.args, "b200_fast_encode", False))
        key = (B, fp64, synth32, fast)
        if key not in self._codecs:
            # ``-t`` only selects the base class in the reference; the arithmetic is
            # always YCoCg (src/2D-DCT.py:22-23, :298, :449).
            self._codecs[key] = Codec(block_size=B, q=self.QSS, color="YCoCg", perceptual=self.perceptual,
                                      disable_subbands=self.disable_subbands, fp64=fp64, synth_f32=synth32, fast=fast,
                                      device=getattr(self, "device", None))
        return self._codecs[key]

    @staticmethod
    def _check_image(img):
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise ValueError(f"the GPU path needs an 8-bit RGB image, got shape {img.shape} dtype {img.dtype}")

    def encode_fn(self, in_fn, out_fn):
        logging.debug("trace")
        logging.debug(f"in_fn = {in_fn}")
        logging.debug(f"out_fn = {out_fn}")
        img = self.encode_read_fn(in_fn)                       # uint8 HWC RGB
        self._check_image(img)
        self.original_shape = img.shape
        with open(f"{out_fn}_shape.bin", "wb") as file:        # src/2D-DCT.py:285-286
            file.write(struct.pack("iii", *self.original_shape))
        decom_k = self._codec().encode(np.ascontiguousarray(img))   # replaces :276-361
        decom_k = self.compress(decom_k)                        # :364
        output_size = self.encode_write_fn(decom_k, out_fn)     # :369
        return output_size

    def encode(self, in_fn="/tmp/original.png", out_fn="/tmp/encoded"):
        return self.encode_fn(in_fn, out_fn)

    def decode_fn(self, in_fn, out_fn):
        logging.debug("trace")
        logging.debug(f"in_fn = {in_fn}")
        logging.debug(f"out_fn = {out_fn}")
        decom_k = self.decode_read_fn(in_fn)                    # :385
        with open(f"{in_fn}_shape.bin", "rb") as file:          # :386-387
            self.original_shape = struct.unpack("iii", file.read(12))
        decom_k = np.ascontiguousarray(self.decompress(decom_k))     # :392
        if decom_k.dtype != np.uint8:
            raise ValueError(f"code-stream holds {decom_k.dtype}, expected uint8 indices")
        codec = self._codec(decode=True)
        if getattr(self.args, "filter", "no_filter") == "no_filter":
            y = codec.decode(decom_k, self.original_shape[:2])  # replaces :398-466
            y = CT.CoDec.filter(self, y)                        # :461 (identity)
        else:
            # a real post-filter receives the un-clipped float image (:454-461)
            _, y = codec.decode(decom_k, self.original_shape[:2], return_float=True)
            y = CT.CoDec.filter(self, y)
            y = np.clip(y, 0, 255).astype(np.uint8)             # :466
        output_size = self.decode_write_fn(y, out_fn)           # :467
        return output_size

    def decode(self, in_fn="/tmp/encoded", out_fn="/tmp/decoded.png"):
        return self.decode_fn(in_fn, out_fn)

    def encode_decode_array(self, img, out_fn):
        '''In-memory counterpart of the hybrid codec's ``encode_decode_proxy``
        (src/IPP_DCT.py:595-626): same payload files (``<out_fn><ext>``, ``<out_fn>_shape.bin``)
        and the same ``(reconstruction, size)`` as ``encode_fn(tmp_png, out_fn)`` followed by
        ``decode_fn(out_fn, tmp_png)`` -- without the two temporary PNGs and without re-reading
        the code-stream: the frame goes to the GPU once, the indices come back for the entropy
        stage, and the reconstruction is decoded from the indices still on the device.  (PNG and
        the entropy coders are lossless, so the decoder sees the very same indices.)'''
        import torch
        img = np.ascontiguousarray(img)
        self._check_image(img)
        self.original_shape = img.shape
        with open(f"{out_fn}_shape.bin", "wb") as file:
            file.write(struct.pack("iii", *self.original_shape))
        x = torch.from_numpy(img).cuda()
        k_dev = self._codec().encode(x)
        dec = self._codec(decode=True)
        if getattr(self.args, "filter", "no_filter") == "no_filter":
            y_dev = dec.decode(k_dev, img.shape[:2])             # overlaps with the entropy stage below
            decom_k = self.compress(k_dev if getattr(self, "accepts_device_arrays", False) else k_dev.cpu().numpy())
            output_size = self.encode_write_fn(decom_k, out_fn)
            y = CT.CoDec.filter(self, y_dev.cpu().numpy())
        else:
            _, yf = dec.decode(k_dev, img.shape[:2], return_float=True)
            decom_k = self.compress(k_dev if getattr(self, "accepts_device_arrays", False) else k_dev.cpu().numpy())
            output_size = self.encode_write_fn(decom_k, out_fn)
            y = np.clip(CT.CoDec.filter(self, yf.cpu().numpy()), 0, 255).astype(np.uint8)
        return y, output_size

    def optimize_block_size(self):
        '''src/2D-DCT.py:533-579 on the GPU: J = rate + Lambda*RMSE per block size, rate = bytes of the
        entropy-coded indices.  The loop body is reproduced as the reference RUNS it, not as it reads: it is
        called from __init__ (:99-103) before ``self.offset = 128`` is assigned (:107-110), so ``self.offset`` still
        is the [0, 0, 0] of the colour stage (src/YCoCg.py:28-29) -- no -128 on the pixels, no +128 on the
        indices (VCFB_F_NO_OFFSET); the dequantiser gets the quantiser's own indices, never narrowed to uint8
        (VCFB_F_NOWRAP, :565-568); neither -p nor -x applies (:540-556).  tests/golden/ref_flow_L_*.npz hold the
        J values the unmodified reference logs; tests/test_plugin.py compares.'''
        logging.debug("trace")
        best = 1000000
        img = self.encode_read()
        self._check_image(img)
        img = np.ascontiguousarray(img)
        for block_size in [2**i for i in range(1, 8)]:
            if img.shape[0] % block_size or img.shape[1] % block_size:
                logging.warning(f"block_size={block_size} skipped (the reference applies no padding here)")
                continue
            decom_k = Codec(block_size=block_size, q=self.QSS, no_offset=True).encode(img)       # :540-559
            decom_k_bytes = self.compress(decom_k)
            decom_k_bytes.seek(0)
            rate = len(decom_k_bytes.read())
            st = stats_dict(rd_stats_fused(img, block_size, [self.QSS], nowrap=True, no_offset=True,
                                           hist=False)[0].cpu().numpy())                        # :565-573
            RMSE = float(np.sqrt(float(st["sse"].sum()) / st["nsamples"]))                       # :574
            J = rate + self.Lambda * RMSE
            logging.debug(f"J={J} for block_size={block_size}")
            if J < best:
                best = J
                self.block_size = block_size


if __name__ == "__main__":
    main.main(parser.parser, logging, CoDec)
