#!/usr/bin/env python3
"""Pack the bundled tokenizer model (vocab.txt + bpe.codes) into one compressed asset.

The two files are the tokenizer's *model data* (the equivalent of weights), not code; they are
read from the reference checkout (genz_tokenize/data/, loaded at tokenize.py:19,23) in the build
container and stored as genz_tokenize_b200/data/bundled.pack so that `Tokenize()` works on a box
that has no /root/reference.  Usage: python tools/pack_bundled.py [/root/reference]
"""
import os
import struct
import sys
import zlib

MAGIC = b"GZTPACK1"


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(ref, "genz_tokenize", "data")
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "genz_tokenize_b200", "data", "bundled.pack")
    out = [MAGIC, struct.pack("<I", 2)]
    for name in ("vocab.txt", "bpe.codes"):
        raw = open(os.path.join(src, name), "rb").read()
        comp = zlib.compress(raw, 9)
        nb = name.encode()
        out += [struct.pack("<I", len(nb)), nb, struct.pack("<QQI", len(raw), len(comp), zlib.crc32(raw)), comp]
    with open(dst, "wb") as f:
        f.write(b"".join(out))
    print("wrote", os.path.normpath(dst), os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
