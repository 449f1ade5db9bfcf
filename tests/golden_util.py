"""Shared checks: run any object with the reference's surface against tests/golden/golden_v1.json.gz.

Used twice: for the CPU oracle (not gpu) and for the CUDA `Tokenize` (gpu), so both are pinned to
the same reference-generated vectors.
"""
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with gzip.open(os.path.join(HERE, "golden", "golden_v1.json.gz"), "rb") as f:
        return json.loads(f.read().decode("ascii"))


def load_decode_golden():
    """Reference `decode` outputs on pad-run shaped rows (oracle/gen_golden_decode.py)."""
    with gzip.open(os.path.join(HERE, "golden", "decode_v2.json.gz"), "rb") as f:
        return json.loads(f.read().decode("ascii"))["blocks"]


def load_padtt_golden():
    """Reference pair calls with a pad token that is a vocab word -- pad id != 0 (oracle/gen_golden_padtt.py)."""
    with gzip.open(os.path.join(HERE, "golden", "padtt_v1.json.gz"), "rb") as f:
        return json.loads(f.read().decode("ascii"))["blocks"]


def _norm(out):
    if "offset" in out:
        out = dict(out)
        out["offset"] = [list(p) for p in out["offset"]]
    return out


def check_case(tok, c, where=""):
    kw = dict(c["kw"])
    exp = c["out"]
    if "raises" in exp:
        try:
            got = tok(c["text"], c["pair"], **kw)
        except ValueError:
            return
        raise AssertionError("%s expected ValueError for %r / %r %r, got %r" % (where, c["text"], c["pair"], kw, got))
    got = _norm(tok(c["text"], c["pair"], **kw))
    assert got == _norm(exp), "%s mismatch for text=%r pair=%r kw=%r\n got=%r\n exp=%r" % (where, c["text"], c["pair"], kw, got, exp)


def check_cases(tok, cases, where=""):
    for i, c in enumerate(cases):
        check_case(tok, c, "%s[%d]" % (where, i))


# ---- digests over many rows: pins tens of thousands of reference rows without storing them -------------------------
def row_bytes(status, ids=None, mask=None, seq=None, tt=None):
    """Canonical bytes of one encoded row: b"E" for a row on which the reference raises ValueError, else b"R" followed by,
    per field, an int32 length and the values (input_ids int32, attention_mask uint8, sequence_id / token_type_ids int8 with
    None as -1; the last two only for pairs)."""
    import numpy as np
    if status:
        return b"E"
    out = [b"R"]
    for v, dt in ((ids, np.int32), (mask, np.uint8), (seq, np.int8), (tt, np.int8)):
        if v is None:
            continue
        a = np.asarray([(-1 if x is None else x) for x in v] if isinstance(v, list) else v).astype(dt)
        out.append(np.int32(len(a)).tobytes())
        out.append(a.tobytes())
    return b"".join(out)


def load_encode_digests():
    with open(os.path.join(HERE, "golden", "encode_digest_v1.json"), "r") as f:
        return json.load(f)["configs"]
