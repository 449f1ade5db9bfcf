#!/usr/bin/env python3
"""Generate tests/golden/padtt_v1.json.gz by IMPORTING THE UNMODIFIED REFERENCE (/root/reference): sentence-pair calls of
`Tokenize.__call__` (tokenize.py:184-259) with a pad token that is a word of vocab.txt, so that the pad id is not 0.  The
reference pads `token_type_ids` through `__padding` (tokenize.py:256-258 calling :141-146), i.e. with the PAD ID, and the
attention mask compares ids with that id (tokenize.py:148-152): a word of the text that equals the pad token is masked out.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden_padtt.py
"""
import gzip
import json
import os
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from genz_tokenize.tokenize import Tokenize  # noqa: E402  (the reference)
from genz_tokenize_b200 import workload  # noqa: E402


def main():
    blocks = []
    for pad_tok in [",", "của", "sinh_viên", "hel@@"]:           # pad ids below and above 127 (int8 / GENZTOK_PAD_MARK)
        tok = Tokenize(pad_token=pad_tok)
        pad = tok.encoder[pad_tok]
        ta = workload.unpack(*workload.generate_hashed(11, 0, 60, 0, 1, 9, 0.05))
        tb = workload.unpack(*workload.generate_hashed(11, 0, 60, 1, 0, 7, 0.05))
        ta[3] = "xin " + pad_tok + " chào " + pad_tok                                   # the pad token inside the text
        tb[5] = pad_tok
        calls = []
        for i, (a, b) in enumerate(zip(ta, tb)):
            for kw in ({"max_len": 16}, {"max_len": 32}, {"max_len": 21}, {"max_len": 64, "truncation": False}, {"max_len": 8, "padding": False}, {}):
                if i % 3 and kw.get("max_len") in (21, 64):
                    continue
                try:
                    out = tok(a, b, **kw)
                except ValueError as e:
                    out = {"raises": str(e)}
                calls.append({"text": a, "pair": b, "kw": kw, "out": out})
            if i % 10 == 0:                                                               # single sentences too (mask by pad id)
                calls.append({"text": a, "pair": None, "kw": {"max_len": 16}, "out": tok(a, max_len=16)})
        blocks.append({"pad_token": pad_tok, "pad_id": int(pad), "calls": calls})
    out = os.path.join(ROOT, "tests", "golden", "padtt_v1.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps({"meta": {"generator": "oracle/gen_golden_padtt.py", "reference": REF}, "blocks": blocks},
                           ensure_ascii=True, separators=(",", ":")).encode("ascii"))
    print(out, os.path.getsize(out), "bytes,", sum(len(b["calls"]) for b in blocks), "calls in", len(blocks), "blocks; pad ids", [b["pad_id"] for b in blocks])


if __name__ == "__main__":
    main()
