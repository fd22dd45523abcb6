'''Argument parser shared by all stages (interface of the reference's src/parser.py).'''
import argparse


def int_or_str(text):
    try:
        return int(text)
    except ValueError:
        return text


def encode(codec):
    return codec.encode()


def decode(codec):
    return codec.decode()


with open("/tmp/description.txt") as f:
    description = f.readline()

parser = argparse.ArgumentParser(description=description, exit_on_error=False)
parser.add_argument("-g", "--debug", action="store_true")
subparser = parser.add_subparsers(dest="subparser_name")
parser_encode = subparser.add_parser("encode")
parser_decode = subparser.add_parser("decode")
parser_encode.set_defaults(func=encode)
parser_decode.set_defaults(func=decode)
