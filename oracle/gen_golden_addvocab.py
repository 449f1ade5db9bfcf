#!/usr/bin/env python3
"""Generate tests/golden/addvocab_v1.json.gz by IMPORTING THE UNMODIFIED REFERENCE (/root/reference): what `add_vocab_file`
and `add_bpe_file` do when called AFTER construction (tokenize.py:44-57).  `encoder` grows (duplicates move to the current
size, tokenize.py:51), `decoder` stays as __init__ built it (tokenize.py:40) -- so the added ids decode to the unk token and a
moved word still decodes from its old id -- and `bpe_ranks` is replaced.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden_addvocab.py
"""
import gzip
import json
import os
import sys
import tempfile

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)

from genz_tokenize.tokenize import Tokenize  # noqa: E402  (the reference)

VOCAB2 = "zzqx 5\nsinh_viên 3\nqqzz@@ 2\nnospace\n"
CODES2 = "#version: 0.2\nh e\nl l\nhe ll\nhell o</w>\n"


def main():
    with tempfile.TemporaryDirectory() as td:
        v2, c2 = os.path.join(td, "v2.txt"), os.path.join(td, "c2.codes")
        open(v2, "w", encoding="utf-8").write(VOCAB2)
        open(c2, "w", encoding="utf-8").write(CODES2)
        tok = Tokenize()
        n0 = tok.vocab_size()
        tok.add_vocab_file(v2)
        texts = ["zzqx sinh_viên xin chào", "hello nospac qqzz zzqx", "sinh_viên công_nghệ"]
        steps = [{"after": "add_vocab_file", "vocab_size": tok.vocab_size(), "enc": {w: tok.encoder.get(w) for w in ["zzqx", "sinh_viên", "qqzz@@", "nospac", "xin"]},
                  "decoder_len": len(tok.decoder), "calls": [{"text": t, "out": tok(t, max_len=12)} for t in texts],
                  "pairs": [{"text": texts[0], "pair": texts[1], "out": tok(texts[0], texts[1], max_len=16)}],
                  "decode": [{"ids": ids, "out": tok.decode(ids)} for ids in ([1, 770, n0, n0 + 1, n0 + 2, 2], [n0 + 3, 5, 770], tok(texts[0])["input_ids"])]}]
        tok.add_bpe_file(c2)
        steps.append({"after": "add_bpe_file", "vocab_size": tok.vocab_size(), "n_ranks": len(tok.bpe_ranks),
                      "calls": [{"text": t, "out": tok(t, max_len=12)} for t in texts + ["hello hell he"]],
                      "bpe": [{"w": w, "out": tok.bpe(w)} for w in ["hello", "hell", "xin", "sinh_viên"]],
                      "decode": [{"ids": ids, "out": tok.decode(ids)} for ids in ([1, 770, n0, 2], tok("hello zzqx")["input_ids"])]})
    out = os.path.join(ROOT, "tests", "golden", "addvocab_v1.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps({"meta": {"generator": "oracle/gen_golden_addvocab.py", "reference": REF}, "vocab2": VOCAB2, "codes2": CODES2, "n0": n0, "steps": steps},
                           ensure_ascii=True, separators=(",", ":")).encode("ascii"))
    print(out, os.path.getsize(out), "bytes;", json.dumps(steps[0]["enc"], ensure_ascii=False), steps[0]["decode"][0]["out"])


if __name__ == "__main__":
    main()
