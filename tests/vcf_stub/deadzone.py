'''Deadzone quantizer stage: -q / QSS, -f filter (decode only).'''
import importlib
with open("/tmp/description.txt", 'w') as f:
    f.write(__doc__)
import parser

parser.parser_encode.add_argument("-q", "--QSS", type=parser.int_or_str, default=32)
parser.parser_decode.add_argument("-q", "--QSS", type=parser.int_or_str, default=32)
parser.parser_decode.add_argument("-f", "--filter", type=parser.int_or_str, default="no_filter")
args = parser.parser.parse_known_args()[0]
try:
    denoiser = importlib.import_module(args.filter)
except AttributeError:
    denoiser = importlib.import_module("no_filter")


class CoDec(denoiser.CoDec):
    def __init__(self, args, min_index_val=0, max_index_val=255):
        super().__init__(args)
        self.QSS = args.QSS
