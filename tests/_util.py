"""Helpers shared by the test modules."""


def parse_flags(flags):
    flags = [str(f) for f in flags]
    kw = dict(B=8, q=32, perceptual=False, disable_subbands=False)
    i = 0
    while i < len(flags):
        f = flags[i]
        if f == "-B":
            kw["B"] = int(flags[i + 1]); i += 2
        elif f == "-q":
            kw["q"] = int(flags[i + 1]); i += 2
        elif f == "-p":
            kw["perceptual"] = True; i += 1
        elif f == "-x":
            kw["disable_subbands"] = True; i += 1
        elif f == "-t":
            i += 2   # the reference keeps the YCoCg arithmetic (src/2D-DCT.py:22-23)
        else:
            raise AssertionError(f)
    return kw
