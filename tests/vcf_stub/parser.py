'''Command line shared by every stage of the stand-in chain.

Stages extend `parser_encode` / `parser_decode` while they are being imported; the first
line of /tmp/description.txt (written by whichever stage is imported first) becomes the
program description -- the handshake the drop-in plugin performs before importing this
module.'''
import argparse as _ap


def int_or_str(text):
    return int(text) if text.lstrip("+-").isdigit() else text


def _make():
    with open("/tmp/description.txt") as fh:
        top = _ap.ArgumentParser(description=fh.readline(), exit_on_error=False)
    top.add_argument("-g", "--debug", action="store_true")
    sub = top.add_subparsers(dest="subparser_name")
    made = {}
    for verb in ("encode", "decode"):
        made[verb] = sub.add_parser(verb)
        made[verb].set_defaults(func=lambda codec, _v=verb: getattr(codec, _v)())
    return top, made["encode"], made["decode"]


parser, parser_encode, parser_decode = _make()
