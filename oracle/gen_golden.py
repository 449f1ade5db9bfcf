#!/usr/bin/env python3
"""Generate tests/golden/*.json.gz by IMPORTING THE UNMODIFIED REFERENCE (/root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden.py
The fixtures pin oracle/genztok_oracle.c and the CUDA engine to the reference's actual behaviour
(the reference's own tests hold no tokenizer vectors -- SURVEY.md §4 -- except README.md:11-15,
which is case "readme" below).
"""
import gzip
import hashlib
import itertools
import json
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from genz_tokenize.tokenize import Tokenize  # noqa: E402  (the reference)
from genz_tokenize_b200 import workload  # noqa: E402  (data generator only)


def call(tok, text, pair=None, **kw):
    try:
        return tok(text, pair, **kw)
    except ValueError as e:
        return {"raises": "ValueError", "msg": str(e)}


def case(tok, text, pair=None, **kw):
    return {"text": text, "pair": pair, "kw": kw, "out": call(tok, text, pair, **kw)}


def main():
    tok = Tokenize()
    G = {"meta": {"generator": "oracle/gen_golden.py", "reference": REF,
                  "vocab_size": tok.vocab_size(), "n_ranks": len(tok.bpe_ranks)}}

    # ---- README known answer (README.md:11-15)
    G["readme"] = [case(tok, "sinh_viên công_nghệ", "hello", max_len=10, padding=True, truncation=True)]
    G["readme_decode"] = {"ids": [1, 770, 2], "out": tok.decode([1, 770, 2])}

    # ---- hand-written corner cases
    texts = [
        "", " ", "   \t\n ", "a", "hello", "hello\n", "a\n", "công_nghệ\n", "xin chào", "ab\ncd \n ef\r\ngh\n\nij kl​mn\x1fop\x85q",
        "a b c d e f g h i j k l m n o p q r　s",
        "x​y﻿z᠎w\x00v", "aaa", "aaaa", "aaaaa", "a" * 33, "ab" * 40, "nnnnnnngggggg", "hello</w>", "a</w>b", "</w>",
        "@@", "a@@ b", "hel@@ lo", "<s> </s> <pad> <unk> <mask>", "#version: 0.2", "#version:0.2", "😀", "中文 😀😀 𐍈", "\ud800", "a\udfffb",
        "sinh_viên công_nghệ thông_tin Việt_Nam", "Hà_Nội , ngày 1 tháng 1 năm 2020 .", "x" * 1200, "é" * 70 + " " + "ộ" * 45,
        "https://example.com/a/b?c=d&e=f#g", "a\nb\nc\n", "\na", "a \n", "a\n\n", "tôi\x0bbạn\x0ccậu\x1cmày\x1dnó\x1ehọ",
        "A B C D E F G H I J K L M N O P", "th ng nh", "t h", "ng", "1234567890123456", "12345678901234567", "abcdefghijklmnop qrstuvwxyzabcdefg",
    ]
    calls = []
    for t in texts:
        calls.append(case(tok, t))
        calls.append(case(tok, t, max_len=8))
        calls.append(case(tok, t, "hello", max_len=12))
    pairs = [("sinh_viên công_nghệ", "hello"), ("sinh_viên công_nghệ", ""), ("", ""), ("", "a"), ("a b c d e f", "g h i j"),
             ("</s>", "</s>"), ("<s>", "<s> </s>"), ("hello\n", "a\n"), ("x", "y" * 50)]
    for a, b in pairs:
        for ml in [None, -3, -2, -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 64]:
            for padding in (True, False):
                for trunc in (True, False):
                    kw = dict(padding=padding, truncation=trunc)
                    if ml is not None:
                        kw["max_len"] = ml
                    calls.append(case(tok, a, b, **kw))
                    if b in ("hello", ""):
                        calls.append(case(tok, a, None, **kw))
    for t, p in [("xin chào các bạn", None), ("xin chào\n các bạn", "tôi là sinh_viên"), ("", ""), ("a", None), ("hello\n", "hello")]:
        calls.append(case(tok, t, p, return_offset=True))
        calls.append(case(tok, t, p, max_len=6, return_offset=True))
    G["calls"] = calls

    # ---- randomised differential rows (noise-heavy), small max_len so the fixture stays small
    wl = workload.default_wordlist()
    rnd = []
    for seed, n, lo, hi, noise, ml, padding, trunc, paired in [
        (11, 400, 0, 6, 0.15, 12, True, True, False), (12, 400, 0, 6, 0.15, 16, True, True, True),
        (13, 200, 3, 13, 0.05, 32, True, True, True), (14, 200, 3, 13, 0.05, None, True, True, True),
        (15, 150, 0, 8, 0.2, 10, False, True, True), (16, 150, 0, 8, 0.2, 10, True, False, True),
        (17, 150, 0, 5, 0.3, 7, True, True, True), (18, 100, 3, 13, 0.0, 128, True, True, False),
        (19, 60, 20, 60, 0.1, 24, True, True, True),
    ]:
        tb, to = workload.generate(seed, n, lo, hi, noise, wl)
        ts = workload.unpack(tb, to)
        ps = workload.unpack(*workload.generate(seed + 1000, n, lo, hi, noise, wl)) if paired else [None] * n
        kw = dict(padding=padding, truncation=trunc)
        if ml is not None:
            kw["max_len"] = ml
        rows = [call(tok, t, p, **kw) for t, p in zip(ts, ps)]
        rnd.append({"gen": dict(seed=seed, n=n, lo=lo, hi=hi, noise=noise, paired=paired), "kw": kw,
                    "texts": ts, "pairs": ps if paired else None, "out": rows})
    G["random"] = rnd

    # ---- bpe() strings
    bw = ["hello", "hello\n", "a", "aaa", "aaaa", "nnnggg", "công_nghệ", "công_nghệ\n", "xinchào", "Việt_Nam", "#version:", "0.2",
          "hello</w>", "a</w>", "</w>", "😀", "中文", "x" * 100, "abcdefghijklmnopqrstuvwxyz" * 3, "ng", "n", "th", "nh", "ngh"]
    rng = np.random.default_rng(7)
    bw += [wl.words[i] for i in rng.integers(0, len(wl.words), size=600)]
    bw += [workload._noise_token(rng, wl).split()[0] for _ in range(600)]
    G["bpe"] = [{"w": w, "out": tok.bpe(w)} for w in bw if w]
    # digest over the whole vocab and all merge concatenations (pins 97k words without storing them)
    allw = [w[:-2] if w.endswith("@@") else w for w in tok.encoder.keys()]
    allw += ["".join(k).replace("</w>", "") for k in tok.bpe_ranks.keys()]
    allw = [w for w in allw if w and not any(c.isspace() for c in w)]
    hs = hashlib.sha256()
    for w in allw:
        hs.update(tok.bpe(w).encode("utf-8"))
        hs.update(b"\n")
    G["bpe_digest"] = {"n": len(allw), "sha256": hs.hexdigest(),
                       "how": "words = encoder keys minus a trailing '@@' + ''.join(merge pair) minus '</w>', whitespace-free, non-empty, in dict order"}

    # ---- get_sequence_id / get_token_type state machine: every id sequence of length <= 6 over {0,1,2,5}
    sq = []
    for L in range(1, 7):
        for s in itertools.product([0, 1, 2, 5], repeat=L):
            raw = tok.get_sequence_id(list(s))
            try:
                tt = tok.get_token_type(list(raw))
            except ValueError:
                tt = "ValueError"
            sq.append([list(s), raw, tt])
    G["seqid"] = sq

    # ---- decode
    dec = [[1, 770, 2], [], [0, 0], [15117], [15117, 3019, 4, 2], [15117, 3019, 30469], [-1, 99999999, 48422, 48423, 5], [3, 4, 4, 3]]
    for _ in range(300):
        n = int(rng.integers(0, 20))
        dec.append([int(v) for v in rng.integers(-3, tok.vocab_size() + 3, size=n)])
    cont = [v for k, v in tok.encoder.items() if k.endswith("@@")]
    for _ in range(100):
        dec.append([int(cont[int(i)]) for i in rng.integers(0, len(cont), size=int(rng.integers(1, 8)))] + [int(rng.integers(0, 40000))])
    G["decode"] = [{"ids": d, "out": tok.decode(d)} for d in dec]

    # ---- loader quirks through fromFile on custom files
    loaders = []
    specs = [
        ("ab 10\nc 5\nab 3\nd 2\n<unk> 7\ne 1\nnospace\n\n  x y 3  \nlo@@ 1\nhel@@ 2", "h e\nl o\nhe l\nx y z\n\nsingle\nl o\nhel lo</w>\ndropped last"),
        ("﻿a 1\r\nb 2\rc 3\n</s> 9\n<s> 1\nd 4\n", "#version: 0.2\r\na b\rb c</w>\na b\n"),
        ("a 1\nb 2\nc</w> 3\na@@ 4\nab@@ 5\nabc 6\nbc 7\n@@ 8\n", "a b\nab c</w>\nb c</w>\n"),
        ("", ""), ("x 1", "x y"), ("a 1\nb 1\nab 1\naa@@ 1\naa 1\na@@ 1\naaa 1\naaaa 1\n", "a a\naa a</w>\na a</w>\naa aa</w>\n"),
    ]
    probes = ["a", "ab", "abc", "aaa", "aaaa", "aaaaa", "a b c", "hello", "hel lo", "abc ab a", "x y", "e", "nospace nospac", "lo hel d c", "﻿a b c d", "a b"]
    with tempfile.TemporaryDirectory() as td:
        for i, (v, m) in enumerate(specs):
            vp, mp = os.path.join(td, "v%d.txt" % i), os.path.join(td, "m%d.codes" % i)
            with open(vp, "wb") as f:
                f.write(v.encode("utf-8"))
            with open(mp, "wb") as f:
                f.write(m.encode("utf-8"))
            t = Tokenize.fromFile(vp, mp)
            loaders.append({
                "vocab": v, "merges": m, "encoder": t.encoder, "decoder": {str(k): s for k, s in t.decoder.items()},
                "ranks2": [[k[0], k[1], r] for k, r in t.bpe_ranks.items() if len(k) == 2], "n_ranks": len(t.bpe_ranks),
                "vocab_size": t.vocab_size(),
                "special_ids": [t.encoder[x] for x in (t.pad_token, t.bos_token, t.eos_token, t.mask_token, t.unk_token)],
                "calls": [case(t, p) for p in probes] + [case(t, p, "a b", max_len=9) for p in probes],
                "decode": [{"ids": d, "out": t.decode(d)} for d in ([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14], [12, 13, 5], [13, 12], [])],
            })
    G["loaders"] = loaders

    # ---- custom special tokens (constructor kwargs, tokenize.py:7-12)
    t2 = Tokenize(pad_token="[PAD]", bos_token="[CLS]", eos_token="[SEP]", mask_token="[MASK]", unk_token="[UNK]")
    G["custom_specials"] = {"specials": ["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]"], "vocab_size": t2.vocab_size(),
                            "calls": [case(t2, "xin chào zzzqqq", "hello\n", max_len=12)],
                            "decode": [{"ids": [0, 1, 2, 3, 4, 99999999], "out": t2.decode([0, 1, 2, 3, 4, 99999999])}]}

    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(G, ensure_ascii=True, sort_keys=True).encode("ascii"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(calls), "calls,", sum(len(r["out"]) for r in rnd), "random rows,",
          len(sq), "seqid,", len(G["bpe"]), "bpe,", len(G["decode"]), "decode")


if __name__ == "__main__":
    main()
