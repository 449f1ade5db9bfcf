"""Timing harness for the UNMODIFIED Python reference in oracle/_ref (made by oracle/make_ref.py) -- TEST / BASELINE
INFRASTRUCTURE ONLY: the reference's `Tokenize.__call__` (tokenize.py:184-259) and `decode` (:137-139) in one
process, and in a `multiprocessing.Pool` with one `Tokenize()` per worker and ~500-document jobs.  Used by bench.py
(`cpu_baseline`, `--impl reference`) and by nothing in the product.

Workers are started with the 'spawn' method: bench.py has initialised CUDA by the time it times the CPU baseline, and a
forked child of a CUDA process is not safe.
"""
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
_tok = None


def _init():
    global _tok
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import make_ref
    _tok = make_ref.load_reference()()


def _encode_job(job):
    texts, pairs, max_len = job
    n = 0
    for i, t in enumerate(texts):
        try:
            out = _tok(t, pairs[i] if pairs is not None else None, max_len=max_len, padding=True, truncation=True)
        except ValueError:            # tokenize.py:157-159 on over-truncated pairs: the row has no result
            continue
        n += sum(out["attention_mask"])
    return n


def _decode_job(rows):
    return sum(len(_tok.decode(r)) for r in rows)


def available():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import make_ref
    return make_ref.available()


def encode_single(texts, pairs, max_len):
    """(real tokens, seconds) of the reference on this process."""
    if _tok is None:
        _init()
    t0 = time.perf_counter()
    n = _encode_job((texts, pairs, max_len))
    return n, time.perf_counter() - t0


def encode_rows_single(texts, pairs, max_len):
    """The reference's rows themselves (for direct comparisons): list of dicts, None where it raises ValueError."""
    if _tok is None:
        _init()
    out = []
    for i, t in enumerate(texts):
        try:
            out.append(_tok(t, pairs[i] if pairs is not None else None, max_len=max_len, padding=True, truncation=True))
        except ValueError:
            out.append(None)
    return out


def decode_single(rows):
    if _tok is None:
        _init()
    t0 = time.perf_counter()
    n = _decode_job(rows)
    return n, time.perf_counter() - t0


class RefPool:
    """multiprocessing.Pool(workers) with one reference Tokenize() per worker."""

    def __init__(self, workers=None, job_docs=500):
        self.workers = workers or os.cpu_count() or 1
        self.job_docs = job_docs
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init)
        self.pool.map(_encode_job, [(["xin chào"], None, 16)] * self.workers)          # every worker has built its Tokenize()

    def encode(self, texts, pairs, max_len):
        """(real tokens, seconds): the documents cut into jobs of job_docs, mapped over the pool."""
        J = self.job_docs
        jobs = [(texts[i:i + J], pairs[i:i + J] if pairs is not None else None, max_len) for i in range(0, len(texts), J)]
        t0 = time.perf_counter()
        n = sum(self.pool.map(_encode_job, jobs))
        return n, time.perf_counter() - t0

    def decode(self, rows):
        J = self.job_docs
        jobs = [rows[i:i + J] for i in range(0, len(rows), J)]
        t0 = time.perf_counter()
        n = sum(self.pool.map(_decode_job, jobs))
        return n, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
