'''Entropy coding of images with deflate on a B200 GPU, in the TIFF container of the TIFF codec.'''

# Drop-in replacement of the reference's src/TIFF.py (the chain's DEFAULT entropy stage, ``-c TIFF``,
# imported by name in src/no_filter.py:12-21): same base class (``class CoDec(EIC.CoDec)``,
# src/TIFF.py:16), same ``file_extension`` (".tif", :21), same ``compress(img) -> BytesIO`` positioned
# at 0 (:23-31) and ``decompress(bytes) -> ndarray`` (:33-39).  ``tifffile.imwrite(..., compression=
# 'zlib')`` is replaced by vcf_b200.entropy.tiff_zlib: one strip whose zlib stream (deflate +
# Adler-32) is produced by libvcfb200.so on the GPU.  tifffile -- i.e. the stock ``-c TIFF`` decoder
# -- libtiff and Pillow read the file; it is valid but not byte-identical with tifffile's (another
# deflate parse, one strip; sizes in DESIGN.md section 4.4).  Use it like the original:
#
#     python 2D-DCT-B200.py encode -c TIFF-B200        (from the reference's src/)

import io as pyio
import logging
import os
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(_here))
for _p in (os.getcwd(), _repo):            # the reference's src/ (main, parser, EIC) and vcf_b200
    if _p not in sys.path:
        sys.path.append(_p)

import main  # noqa: E402  (reference src/main.py)
with open("/tmp/description.txt", 'w') as f:   # handshake read by src/parser.py:67
    f.write(__doc__)
import parser  # noqa: E402  (reference src/parser.py)
import entropy_image_coding as EIC  # noqa: E402  (reference src/entropy_image_coding.py)

from vcf_b200 import entropy  # noqa: E402


class CoDec(EIC.CoDec):

    accepts_device_arrays = True      # compress() also takes CUDA tensors (see z_lib-B200.py)

    def __init__(self, args):
        logging.debug("trace")
        super().__init__(args)
        self.file_extension = ".tif"

    def compress(self, img):
        '''src/TIFF.py:23-31 with the strip compressed on the GPU.'''
        logging.debug("trace")
        logging.debug(f"img.dtype={img.dtype}")
        compressed_img = pyio.BytesIO(entropy.tiff_zlib(img))     # uint8 / uint16 only, like :26
        compressed_img.seek(0)
        return compressed_img

    def decompress(self, compressed_img):
        '''src/TIFF.py:33-39: tifffile reads the file.  Where tifffile is not installed the same
        file goes through libtiff (OpenCV) -- a container reader on the host either way.'''
        logging.debug("trace")
        try:
            import tifffile
        except ImportError:
            import cv2
            img = cv2.imdecode(np.frombuffer(compressed_img, np.uint8), cv2.IMREAD_UNCHANGED)
            if img is None:
                raise ValueError("not a TIFF file libtiff can read")
            return cv2.cvtColor(img, cv2.COLOR_BGR2RGB) if img.ndim == 3 else img
        img = tifffile.imread(pyio.BytesIO(compressed_img))
        logging.debug(f"img.dtype={img.dtype}")
        return img


if __name__ == "__main__":
    main.main(parser.parser, logging, CoDec)
