#!/usr/bin/env python3
"""Generate tests/golden/decode_v2.json.gz by IMPORTING THE UNMODIFIED REFERENCE (/root/reference): `Tokenize.decode`
(tokenize.py:137-139) on id rows shaped like encoder output and unlike it (trailing / inner / leading runs of pad ids, a
non-pad id behind a run, rows of nothing but pads, ids outside the decoder), for the bundled files and for pad tokens whose
text has other lengths (1 ... 9 bytes, multi-byte, strings vocab.txt also holds).  These are the shapes the CUDA decode
treats specially (pad runs written as periodic text).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden_decode.py
"""
import gzip
import json
import os
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

from genz_tokenize.tokenize import Tokenize  # noqa: E402  (the reference)
from parity_util import pad_run_rows  # noqa: E402  (row generator shared with the tests)

EXOTIC = [-1, -2**31, 2**31 - 1, 48423, 48422, 10**6, 3, 4, 1, 2]
PAD_TOKENS = ["<pad>", "[P]", "\x7f@@", "Ω@@", "zq@@", "a~b", "ặ~", "abcdefg", "abcdefgh", "<padtok>", "ab", "p@@"]


def main():
    rng = np.random.default_rng(2024)
    blocks = []
    for pad_tok in PAD_TOKENS:
        tok = Tokenize(pad_token=pad_tok)
        pad = tok.encoder[pad_tok]
        widths = [1, 2, 3, 4, 8, 12, 16, 20, 33, 64, 128, 132, 256, 260] if pad_tok == "<pad>" else [5, 16, 36, 128]
        for w in widths:
            ids = pad_run_rows(rng, 24 if w > 64 else 40, w, pad, tok.vocab_size(), EXOTIC)
            blocks.append({"pad_token": pad_tok, "pad_id": int(pad), "width": w, "ids": ids.tolist(),
                           "out": [tok.decode(r) for r in ids.tolist()]})
    out = os.path.join(ROOT, "tests", "golden", "decode_v2.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps({"meta": {"generator": "oracle/gen_golden_decode.py", "reference": REF}, "blocks": blocks},
                           ensure_ascii=True, separators=(",", ":")).encode("ascii"))
    print(out, os.path.getsize(out), "bytes,", sum(len(b["ids"]) for b in blocks), "rows in", len(blocks), "blocks")


if __name__ == "__main__":
    main()
