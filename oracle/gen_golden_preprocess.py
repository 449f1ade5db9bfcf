#!/usr/bin/env python3
"""Generate tests/golden/preprocess_v1.json.gz by IMPORTING THE UNMODIFIED REFERENCE's preprocess.py
(/root/reference/genz_tokenize/preprocess.py).  Build container only."""
import gzip
import json
import os
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
from genz_tokenize import preprocess as P  # noqa: E402  (the reference)
from genz_tokenize_b200 import workload  # noqa: E402

FNS = [P.remove_html, P.convert_unicode, P.remove_punctuations, P.remove_emoji, P.remove_URL]
NAMES = ["remove_html", "convert_unicode", "remove_punctuations", "remove_emoji", "remove_URL"]


def main():
    rng = np.random.default_rng(2024)
    wl = workload.default_wordlist()
    hand = [
        "", " ", "plain text", "<b>xin</b> chào <a href='x'>bạn</a>", "1 < 2 > 3", "a <b c <d> e> f", "<<>>", "<>", "<unclosed", "closed> <x", "a<\n\n>b",
        "> > < < >", "<" * 40 + ">" * 40, "x" * 70 + "<tag" + "y" * 70 + ">" + "z" * 70, "tiếng <i>Việt</i> có dấu <br/> ở đây",
        "à á ả ã ạ", "à ế ợ Ấ ửx Ạ", "à", "̀a", "aà", "ằ ẵ ặ Ờ Ự", "ệ ệ", "ấ", "q̀ ẓ", "è" * 50, "ấễự",
        "a!b,c. d? (e) [f] ~g", "!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~", "giá: 1.000.000đ (VAT) — “quote” … ok", "no_punct_here", "___",
        "hi 😀 there　you ☺ ok ‍  end ", "😀", " 😀 ", "a😀b", "a 😀 b", "  nhiều   khoảng \t trắng \n\n ở đây  ", "中文 字符 và tiếng Việt", "a‍b ⏏ ⏩ ⌚ ️ 〰 c",
        "x y z　w", "Ⓛ Ⓜ Ⓝ", "emoji 👍🏽 flags 🇻🇳 done", "\ud800 lone", "a b c", "only   spaces", "\x1c\x1d a \x1e\x1f",
        "see http://x.y/z and https://q http httpx http\nz xhttp://k", "http", "http ", " http", "httpa", "http://", "hTTP://no HTTP://no", "ahttpb chttp", "httphttp://x y",
        "http x http　y http​z", "link:http://a.b/c?d=e&f=g#h.", "h t t p : / /", "http" * 30, "xx http://tiếng.việt/đường dẫn tiếp",
    ]
    rand = []
    alphabet = list("<>/ ab\n\t!?.,;:'\"()[]{}hHtTpP:/-_~") + ["😀", "☺", "ế", "ợ", "à", "é", "õ", "ủ", "ỵ", "ấ",
                                                               "ờ", "　", " ", "‍", " ", "http", "http://", "<a>", "</b>", "中", "Ⓜ", "⏏"]
    for _ in range(400):
        k = int(rng.integers(0, 60))
        parts = []
        for _ in range(k):
            if rng.random() < 0.4:
                parts.append(wl.words[int(rng.integers(0, len(wl.words)))])
                parts.append(" " if rng.random() < 0.7 else "")
            else:
                parts.append(alphabet[int(rng.integers(0, len(alphabet)))])
        rand.append("".join(parts))
    texts = hand + rand
    G = {"meta": {"generator": "oracle/gen_golden_preprocess.py", "reference": REF + "/genz_tokenize/preprocess.py", "ops": NAMES},
         "texts": texts, "out": [[f(t) for t in texts] for f in FNS]}
    path = os.path.join(ROOT, "tests", "golden", "preprocess_v1.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(G, ensure_ascii=True, sort_keys=True).encode("ascii"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(texts), "texts x", len(FNS), "ops")


if __name__ == "__main__":
    main()
