'''File IO and byte accounting (interface of the reference's src/entropy_image_coding.py).'''
import os

import cv2 as cv
import parser

parser.parser_encode.add_argument("-o", "--original", type=parser.int_or_str, default="/tmp/original.png")
parser.parser_encode.add_argument("-e", "--encoded", type=parser.int_or_str, default="/tmp/encoded")
parser.parser_decode.add_argument("-e", "--encoded", type=parser.int_or_str, default="/tmp/encoded")
parser.parser_decode.add_argument("-d", "--decoded", type=parser.int_or_str, default="/tmp/decoded.png")


class CoDec:
    def __init__(self, args):
        self.args = args
        self.encoding = args.subparser_name == "encode"
        self.total_input_size = 0
        self.total_output_size = 0

    def bye(self):
        pass

    def encode_read_fn(self, fn):
        self.total_input_size += os.path.getsize(fn)
        return cv.cvtColor(cv.imread(fn, cv.IMREAD_UNCHANGED), cv.COLOR_BGR2RGB)

    def encode_read(self, fn="/tmp/original.png"):
        return self.encode_read_fn(fn)

    def encode_write_fn(self, codestream, fn):
        codestream.seek(0)
        with open(fn + self.file_extension, "wb") as f:
            f.write(codestream.read())
        n = os.path.getsize(fn + self.file_extension)
        self.total_output_size += n
        return n

    def decode_read_fn(self, fn):
        self.total_input_size += os.path.getsize(fn + self.file_extension)
        return open(fn + self.file_extension, "rb").read()

    def decode_write_fn(self, img, fn):
        cv.imwrite(fn, cv.cvtColor(img, cv.COLOR_RGB2BGR))
        n = os.path.getsize(fn)
        self.total_output_size += n
        return n
