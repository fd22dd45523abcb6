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
        elif f == "-L":
            i += 2   # block-size search: the chosen size is recorded with the vector (L_chosen)
        elif f == "-g":
            i += 1
        elif f == "-f":
            i += 2   # decode-side denoising filter: applied by the tests that know it (golden_filter)
        else:
            raise AssertionError(f)
    return kw


def golden_kw(g):
    """Codec keywords of a golden vector; a -L vector was encoded with the block size the search chose."""
    kw = parse_flags(g["flags"])
    if "L_chosen" in g.files:
        kw["B"] = int(g["L_chosen"])
    return kw


def golden_filter(g):
    """Name of the decode-side filter of a golden vector (src/deadzone.py:33), or None."""
    flags = [str(f) for f in g["flags"]]
    return flags[flags.index("-f") + 1] if "-f" in flags else None
