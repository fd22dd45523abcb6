'''Stand-in for the reference's src/entropy_image_coding.py (base class of the entropy codecs):
only what an entropy plugin touches -- the constructor taking the parsed flags and the
`file_extension` attribute.  The file IO of the chain lives in chain_stub.py.'''


class CoDec:
    file_extension = ".bin"

    def __init__(self, args):
        self.args = args
        self.encoding = getattr(args, "subparser_name", "encode") == "encode"
        self.total_input_size = self.total_output_size = 0
