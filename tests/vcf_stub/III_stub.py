'''Per-frame loop around a 2D codec chosen by -T (interface of the reference's src/III.py:23-59,:120-144).'''
import importlib
import logging
import main
with open("/tmp/description.txt", 'w') as f:
    f.write(__doc__)
import parser

for p in (parser.parser_encode, parser.parser_decode):
    p.add_argument("-T", "--transform", type=str, default="2D-DCT-B200")
    p.add_argument("-N", "--number_of_frames", type=parser.int_or_str, default=3)
args = parser.parser.parse_known_args()[0]
transform = importlib.import_module(args.transform)  # registers -t etc.


class CoDec:
    def __init__(self, args):
        self.args = args
        self.transform_codec = transform.CoDec(args)

    def bye(self):
        pass

    def encode(self):       # the behaviour src/III.py:96-104 intends (call commented out there)
        for i in range(self.args.number_of_frames):
            self.transform_codec.encode_fn("/tmp/original_%04d.png" % i, "/tmp/encoded_%04d" % i)

    def decode(self):       # src/III.py:132-144
        for i in range(self.args.number_of_frames):
            self.transform_codec.decode_fn("/tmp/encoded_%04d" % i, "/tmp/decoded_%04d.png" % i)


if __name__ == "__main__":
    main.main(parser.parser, logging, CoDec)
