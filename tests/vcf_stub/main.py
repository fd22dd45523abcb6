'''Entry point of the stand-in chain: build the codec from the parsed flags, run the verb.'''


def main(parser, logging, CoDec):
    ns, _unknown = parser.parse_known_args()
    logging.basicConfig(level=logging.DEBUG if ns.debug else logging.INFO)
    codec = CoDec(ns)
    try:
        ns.func(codec)
    finally:
        codec.bye()
