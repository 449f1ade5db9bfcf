"""The step before the tokenizer: a streaming reader that turns a text file with one document per
line into the C ABI's packed batches -- (uint8 bytes, int64 offsets[n+1]) -- without building Python strings.

Every document keeps its line terminator as a trailing ASCII space (the `\\n` becomes `0x20`; a `\\r` before it is
whitespace anyway), so the documents of a batch are adjacent in one buffer and tokenise exactly like the stripped lines:
trailing whitespace produces no word (`tokenize.py:106`), whereas a kept `\\n` would attach to the last word and turn it
into `<unk>`.  An empty line is an empty document (`[<s>, </s>]`).
"""
import numpy as np


def iter_line_batches(path, docs_per_batch=1 << 20, read_bytes=64 << 20):
    """Yield (bytes uint8[...], offsets int64[n+1]) with at most `docs_per_batch` documents (lines of `path`) each."""
    docs_per_batch = int(docs_per_batch)
    if docs_per_batch < 1:
        raise ValueError("docs_per_batch must be positive")
    carry = np.empty(0, dtype=np.uint8)          # bytes behind the last complete line
    with open(path, "rb") as f:
        while True:
            raw = f.read(int(read_bytes))
            eof = not raw
            if eof:
                if len(carry) == 0:
                    return
                buf = np.concatenate([carry, np.array([10], dtype=np.uint8)])      # last line without a terminator
                carry = np.empty(0, dtype=np.uint8)
            else:
                buf = np.concatenate([carry, np.frombuffer(raw, dtype=np.uint8)]) if len(carry) else np.frombuffer(raw, dtype=np.uint8).copy()
            ends = np.flatnonzero(buf == 10) + 1                                   # one past every '\n'
            if len(ends) == 0:
                carry = buf
                if eof:
                    return
                continue
            carry = buf[ends[-1]:].copy()
            buf = buf[:ends[-1]]
            buf[ends - 1] = 32
            starts = np.concatenate([[0], ends]).astype(np.int64)
            for b in range(0, len(ends), docs_per_batch):
                e = min(b + docs_per_batch, len(ends))
                lo, hi = starts[b], starts[e]
                yield buf[lo:hi], starts[b:e + 1] - lo
            if eof:
                return
