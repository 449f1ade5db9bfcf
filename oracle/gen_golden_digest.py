#!/usr/bin/env python3
"""Generate tests/golden/encode_digest_v1.json by IMPORTING THE UNMODIFIED REFERENCE (/root/reference): SHA-256 digests over
the reference's `Tokenize.__call__` outputs on 85,000 seeded synthetic rows (the BASELINE configs' shapes plus noisy,
heavily truncated, ragged and unpadded regimes).  The rows themselves are regenerated from the seeds by the tests
(genz_tokenize_b200.workload is deterministic), so only the digests are stored.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden_digest.py
"""
import hashlib
import json
import os
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

from genz_tokenize.tokenize import Tokenize  # noqa: E402  (the reference)
from genz_tokenize_b200 import workload  # noqa: E402  (data generator only)
from golden_util import row_bytes  # noqa: E402

CONFIGS = [
    dict(seed=9001, n=20000, lo=3, hi=13, noise=0.02, paired=False, kw=dict(max_len=128)),
    dict(seed=9002, n=20000, lo=3, hi=13, noise=0.02, paired=True, kw=dict(max_len=256)),
    dict(seed=9003, n=10000, lo=0, hi=8, noise=0.3, paired=True, kw=dict(max_len=12)),
    dict(seed=9004, n=10000, lo=0, hi=8, noise=0.2, paired=True, kw=dict()),
    dict(seed=9005, n=10000, lo=0, hi=8, noise=0.2, paired=True, kw=dict(max_len=10, padding=False)),
    dict(seed=9006, n=10000, lo=0, hi=8, noise=0.2, paired=True, kw=dict(max_len=10, truncation=False)),
    dict(seed=9007, n=5000, lo=20, hi=60, noise=0.1, paired=False, kw=dict(max_len=24)),
]


def main():
    tok = Tokenize()
    wl = workload.default_wordlist()
    out = []
    for c in CONFIGS:
        ts = workload.unpack(*workload.generate(c["seed"], c["n"], c["lo"], c["hi"], c["noise"], wl))
        ps = workload.unpack(*workload.generate(c["seed"] + 1000, c["n"], c["lo"], c["hi"], c["noise"], wl)) if c["paired"] else [None] * c["n"]
        h, hd = hashlib.sha256(), hashlib.sha256()             # hd: decode() of every row's input_ids (tokenize.py:137-139)
        errors = 0
        for t, p in zip(ts, ps):
            try:
                r = tok(t, p, **c["kw"])
            except ValueError:
                h.update(row_bytes(1))
                errors += 1
                continue
            h.update(row_bytes(0, r["input_ids"], r["attention_mask"], r.get("sequence_id"), r.get("token_type_ids")))
            hd.update(tok.decode(r["input_ids"]).encode("utf-8", "surrogatepass") + b"\n")
        out.append(dict(c, sha256=h.hexdigest(), value_errors=errors, decode_sha256=hd.hexdigest()))
        print(c["seed"], c["n"], c["kw"], h.hexdigest()[:16], "errors", errors, flush=True)
    path = os.path.join(ROOT, "tests", "golden", "encode_digest_v1.json")
    with open(path, "w") as f:
        json.dump({"meta": {"generator": "oracle/gen_golden_digest.py", "reference": REF,
                            "row_bytes": "tests/golden_util.py::row_bytes"}, "configs": out}, f, indent=1)
    print(path)


if __name__ == "__main__":
    main()
