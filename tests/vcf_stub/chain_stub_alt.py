'''A second colour-stage module (`-t chain_stub_alt`): like the reference's `-t YCrCb`, it
only changes the base class of the spatial codec, not its arithmetic.'''
import chain_stub


class CoDec(chain_stub.CoDec):
    colour_stage = "alt"
