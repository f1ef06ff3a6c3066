"""Bitstream container of the reference (src/utils/stream_helper.py:19-99): big-endian uint32 headers.
  I-frame file: [height, width, len(y), len(z)] + y-string + z-string        (encode_i / decode_i, :61-82)
  P-frame file: [len(string)] + string                                      (encode_p / decode_p, :85-99)"""
import struct
from pathlib import Path


def get_downsampled_shape(height, width, p, resample_times=1):
    pad_d = p * resample_times
    new_h = (height + pad_d - 1) // pad_d * pad_d
    new_w = (width + pad_d - 1) // pad_d * pad_d
    return int(new_h / p + 0.5), int(new_w / p + 0.5)


def filesize(filepath):
    if not Path(filepath).is_file():
        raise ValueError(f'Invalid file "{filepath}".')
    return Path(filepath).stat().st_size


def encode_i(height, width, y_string, z_string, output):
    with Path(output).open("wb") as f:
        f.write(struct.pack(">4I", height, width, len(y_string), len(z_string)))
        f.write(y_string)
        f.write(z_string)


def decode_i(inputpath):
    with Path(inputpath).open("rb") as f:
        height, width, ny, nz = struct.unpack(">4I", f.read(16))
        y_string = f.read(ny)
        z_string = f.read(nz)
    return height, width, y_string, z_string


def encode_p(string, output):
    with Path(output).open("wb") as f:
        f.write(struct.pack(">I", len(string)))
        f.write(string)


def decode_p(inputpath):
    with Path(inputpath).open("rb") as f:
        (n,) = struct.unpack(">I", f.read(4))
        return f.read(n)
