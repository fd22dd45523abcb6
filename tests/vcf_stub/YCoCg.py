'''Colour stage: adds -a / --quantizer and derives from the quantizer stage.'''
import importlib
with open("/tmp/description.txt", 'w') as f:
    f.write(__doc__)
import parser

parser.parser_encode.add_argument("-a", "--quantizer", default="deadzone")
parser.parser_decode.add_argument("-a", "--quantizer", default="deadzone")
args = parser.parser.parse_known_args()[0]
Q = importlib.import_module(args.quantizer)


class CoDec(Q.CoDec):
    def __init__(self, args):
        super().__init__(args)
