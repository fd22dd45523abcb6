'''Identity decoding filter; picks the entropy codec (-c).'''
import importlib
with open("/tmp/description.txt", 'w') as f:
    f.write(__doc__)
import parser

parser.parser_encode.add_argument("-c", "--entropy_image_codec", default="z_lib")
parser.parser_decode.add_argument("-c", "--entropy_image_codec", default="z_lib")
args = parser.parser.parse_known_args()[0]
EC = importlib.import_module(args.entropy_image_codec)


class CoDec(EC.CoDec):
    def filter(self, img):
        return img
