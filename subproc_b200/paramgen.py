"""Parameter wire format of the reference's paramgen.py: the file an external engine reads.

``write_data`` (paramgen.py:12-19): the values ``(header, w0..w35)`` as one byte each, negatives as
two's complement (``256 + n``, paramgen.py:5-9), followed by a trailing 0x00 -- 38 bytes for the
ProgressPositionMoves parameter set.
"""


def conv_num(num):                                      # paramgen.py:5-9
    return 256 + num if num < 0 else num


def encode(parameters):
    return bytes(bytearray([conv_num(int(p)) for p in parameters]) + bytearray([0]))


def decode(data):
    """inverse of encode: (header, signed weights...)"""
    vals = list(bytearray(data))[:-1]
    return tuple([vals[0]] + [v - 256 if v > 127 else v for v in vals[1:]])


def write_data(file_full_path, parameters):             # paramgen.py:12-19
    with open(file_full_path, "wb") as fout:
        fout.write(encode(parameters))


def read_data(file_full_path):
    with open(file_full_path, "rb") as fin:
        return decode(fin.read())
