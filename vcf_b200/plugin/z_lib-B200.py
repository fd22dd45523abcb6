'''Entropy coding of images with deflate on a B200 GPU, in the .npz container of z_lib.'''

# Drop-in replacement of the reference's src/z_lib.py (the entropy stage selected with
# ``-c z_lib``, imported by name in src/no_filter.py:21): same base class
# (``class CoDec(EIC.CoDec)``, src/z_lib.py:12), same ``file_extension`` (".npz", :17), same
# ``compress(img) -> io.BytesIO`` (:19-23) and ``decompress(bytes) -> ndarray`` (:25-29).  Only
# the deflate call inside ``np.savez_compressed`` is replaced: the array's bytes are compressed
# by libvcfb200.so on the GPU (vcf_b200.entropy.savez_compressed -> vcfb_deflate_dev) and wrapped
# in the same zip member layout, so ``decompress`` -- and the reference's own z_lib.decompress --
# read it with ``np.load``.  The streams are valid deflate but not byte-identical with zlib's
# (a run-length parse; sizes in DESIGN.md section 4.4).  Use it like the original:
#
#     python 2D-DCT-B200.py encode -c z_lib-B200       (from the reference's src/)
#     python z_lib-B200.py encode                      (stand-alone, like src/z_lib.py:31-32)

import io
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

    # compress() also takes CUDA tensors: the GPU transform stages above (2D-DCT-B200, III-B200)
    # check this and keep the indices in HBM instead of handing over a host copy
    accepts_device_arrays = True

    def __init__(self, args):
        logging.debug(f"trace args={args}")
        super().__init__(args)
        self.file_extension = ".npz"

    def compress(self, img):
        '''src/z_lib.py:19-23 with the deflate stream produced on the GPU.  ``img``: numpy array
        or a CUDA tensor (the batched driver hands the indices over without a host round trip
        of its own; the checksum of the zip member still needs the host copy).'''
        logging.debug(f"trace img={img}")
        compressed_img = io.BytesIO()
        entropy.savez_compressed(compressed_img, a=img)
        return compressed_img

    def decompress(self, compressed_img):
        '''src/z_lib.py:25-29, unchanged: the host's inflate (np.load) reads the stream.'''
        logging.debug(f"trace compressed_img={compressed_img}")
        compressed_img = io.BytesIO(compressed_img)
        img = np.load(compressed_img)['a']
        return img


if __name__ == "__main__":
    main.main(parser.parser, logging, CoDec)
