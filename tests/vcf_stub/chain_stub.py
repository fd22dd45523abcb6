'''Stand-in for the stages below the spatial transform: colour stage, deadzone stage,
decoding filter, entropy codec (zlib in an .npz container) and file IO in ONE module.

Selected with `-t chain_stub`.  It offers the attributes the drop-in plugin touches:
flags -a/--quantizer, -q/--QSS, -f/--filter, -c/--entropy_image_codec, -o/-e/-d;
`encoding`, `QSS`, `file_extension`; `encode_read_fn`, `encode_read`, `compress`,
`encode_write_fn`, `decode_read_fn`, `decompress`, `filter`, `decode_write_fn`, `bye`.'''
import io
import os

import cv2
import numpy as np

with open("/tmp/description.txt", "w") as fh:
    fh.write(__doc__)
import parser as cli

for side in (cli.parser_encode, cli.parser_decode):
    side.add_argument("-a", "--quantizer", default="deadzone")
    side.add_argument("-q", "--QSS", type=cli.int_or_str, default=32)
    side.add_argument("-c", "--entropy_image_codec", default="npz_zlib")
    side.add_argument("-e", "--encoded", type=cli.int_or_str, default="/tmp/encoded")
cli.parser_encode.add_argument("-o", "--original", type=cli.int_or_str, default="/tmp/original.png")
cli.parser_decode.add_argument("-d", "--decoded", type=cli.int_or_str, default="/tmp/decoded.png")
cli.parser_decode.add_argument("-f", "--filter", type=cli.int_or_str, default="no_filter")

FILTERS = {
    "no_filter": lambda img: img,
    # a non-identity filter: must receive the un-clipped FLOAT image
    "half_filter": lambda img: _require_float(img) * 0.5 + 300.25,
}


def _require_float(img):
    assert img.dtype.kind == "f", "post-filters get the un-clipped float image"
    return img


class CoDec:
    file_extension = ".npz"

    def __init__(self, args):
        self.args = args
        self.encoding = args.subparser_name == "encode"
        self.QSS = args.QSS
        self.total_input_size = self.total_output_size = 0
        # -c <module>: another entropy codec (the reference makes it the base class, src/no_filter.py:21;
        # the stand-in delegates to an instance)
        self.entropy = None
        if getattr(args, "entropy_image_codec", "npz_zlib") != "npz_zlib":
            import importlib
            self.entropy = importlib.import_module(args.entropy_image_codec).CoDec(args)
            self.file_extension = self.entropy.file_extension
            self.accepts_device_arrays = getattr(self.entropy, "accepts_device_arrays", False)

    def bye(self):
        pass

    # ---- file IO ------------------------------------------------------------------
    def encode_read_fn(self, fn):
        self.total_input_size += os.path.getsize(fn)
        return cv2.cvtColor(cv2.imread(fn, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB)

    def encode_read(self, fn="/tmp/original.png"):
        return self.encode_read_fn(fn)

    def encode_write_fn(self, codestream, fn):
        path = fn + self.file_extension
        with open(path, "wb") as out:
            out.write(codestream.getvalue())
        self.total_output_size += os.path.getsize(path)
        return os.path.getsize(path)

    def decode_read_fn(self, fn):
        path = fn + self.file_extension
        self.total_input_size += os.path.getsize(path)
        with open(path, "rb") as src:
            return src.read()

    def decode_write_fn(self, img, fn):
        cv2.imwrite(fn, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        self.total_output_size += os.path.getsize(fn)
        return os.path.getsize(fn)

    # ---- entropy codec ----------------------------------------------------------------
    def compress(self, img):
        if self.entropy is not None:
            return self.entropy.compress(img)
        buf = io.BytesIO()
        np.savez_compressed(buf, a=img)
        return buf

    def decompress(self, data):
        if self.entropy is not None:
            return self.entropy.decompress(data)
        return np.load(io.BytesIO(data))["a"]

    # ---- decoding filter ----------------------------------------------------------------
    def filter(self, img):
        return FILTERS[getattr(self.args, "filter", "no_filter")](img)
