def main(parser, logging, CoDec):
    args = parser.parse_known_args()[0]
    logging.basicConfig(level=logging.DEBUG if args.debug else logging.INFO)
    codec = CoDec(args)
    args.func(codec)
    codec.bye()
