'''Entropy coding with zlib in an .npz container (interface of the reference's src/z_lib.py).'''
import io

import numpy as np
with open("/tmp/description.txt", 'w') as f:
    f.write(__doc__)
import parser  # noqa: F401
import entropy_image_coding as EIC


class CoDec(EIC.CoDec):
    def __init__(self, args):
        super().__init__(args)
        self.file_extension = ".npz"

    def compress(self, img):
        b = io.BytesIO()
        np.savez_compressed(file=b, a=img)
        return b

    def decompress(self, data):
        return np.load(io.BytesIO(data))['a']
