"""Deterministic synthetic Vietnamese-like text; host-side, data only.

Word list W = the vocab.txt entries that do not end in '@@' (each round-trips to exactly one
token), weights = their count column.  ``generate(seed, n, lo, hi, noise)`` draws, per document,
k ~ U{lo..hi} words proportional to weight and joins them with one ASCII space; documents are made
in chunks of ``CHUNK`` with ``numpy.random.default_rng(seed + chunk_index)`` so that any shard of a
big batch can be regenerated independently (multi-GPU sharding by document).

``noise`` replaces that fraction of words by adversarial tokens (glued words, random ASCII /
diacritics / digits, vocab-only punctuation, a word directly followed by '\\n', exotic Unicode
whitespace, literal '</w>' and '@@', a few very long tokens) so that multi-piece words, <unk> and
the pre-split corner cases are exercised.

The result is the C-ABI's packed form: (uint8 bytes, int64 offsets[n+1]).
"""
import numpy as np

from .data import bundled_paths

CHUNK = 1 << 20

_EXOTIC_WS = ["\t", "\n", "\x0b", "\x0c", "\r", "\x1c", "\x1d", "\x1e", "\x1f", "\x85", "\xa0", " ", " ",
              " ", " ", " ", " ", " ", " ", "　", "\r\n", "\n\n", " \n "]
_NOT_WS = ["​", "﻿", "\x00", "᠎"]
_PUNCT = list("$'()*;[\\]^`{|}~")
_DIACRITICS = "àáảãạăằắẳẵặâầấẩẫậèéẻẽẹêềếểễệìíỉĩịòóỏõọôồốổỗộơờớởỡợùúủũụưừứửữựỳýỷỹỵđ"


class WordList:
    """vocab.txt words (no '@@' suffix) + sampling weights, as packed UTF-8."""

    def __init__(self, vocab_file=None):
        if vocab_file is None:
            vocab_file = bundled_paths()[0]
        words, counts = [], []
        with open(vocab_file, "r", encoding="utf-8") as f:
            for line in f:
                line = line.strip()
                idx = line.rfind(" ")
                if idx <= 0:
                    continue
                w = line[:idx]
                if w.endswith("@@") or any(ch.isspace() for ch in w):
                    continue
                try:
                    c = int(line[idx + 1:])
                except ValueError:
                    continue
                words.append(w)
                counts.append(c)
        self.words = words
        enc = [w.encode("utf-8") for w in words]
        self.wlen = np.array([len(e) for e in enc], dtype=np.int64)
        self.wstart = np.zeros(len(enc), dtype=np.int64)
        np.cumsum(self.wlen[:-1], out=self.wstart[1:])
        self.blob = np.frombuffer(b"".join(enc), dtype=np.uint8)
        p = np.asarray(counts, dtype=np.float64)
        self.p = p / p.sum()
        self.cdf = np.cumsum(self.p)
        self.cdf[-1] = 1.0


_default_wl = None


def default_wordlist():
    global _default_wl
    if _default_wl is None:
        _default_wl = WordList()
    return _default_wl


def _noise_token(rng, wl):
    kind = int(rng.integers(0, 12))
    pick = lambda: wl.words[int(np.searchsorted(wl.cdf, rng.random()))]
    if kind == 0:   # glued words
        return pick() + pick()
    if kind == 1:   # random ASCII letters
        return "".join(chr(int(c)) for c in rng.integers(97, 123, size=int(rng.integers(1, 15))))
    if kind == 2:   # random diacritics
        return "".join(_DIACRITICS[int(i)] for i in rng.integers(0, len(_DIACRITICS), size=int(rng.integers(1, 9))))
    if kind == 3:   # digits
        return "".join(chr(int(c)) for c in rng.integers(48, 58, size=int(rng.integers(1, 12))))
    if kind == 4:   # vocab-only punctuation glued to a word
        return pick() + _PUNCT[int(rng.integers(0, len(_PUNCT)))]
    if kind == 5:   # word directly followed by '\n'
        return pick() + "\n"
    if kind == 6:   # exotic whitespace between two words
        return pick() + _EXOTIC_WS[int(rng.integers(0, len(_EXOTIC_WS)))] + pick()
    if kind == 7:   # look-alike non-whitespace
        return pick() + _NOT_WS[int(rng.integers(0, len(_NOT_WS)))] + pick()
    if kind == 8:   # literal markers
        return pick() + ["</w>", "@@", "@@ ", "</s>", "<s>", "<pad>", "<unk>"][int(rng.integers(0, 7))]
    if kind == 9:   # non-BMP / CJK / emoji code points
        return "".join(chr(int(c)) for c in rng.choice([0x4E2D, 0x6587, 0x1F600, 0x1F4A9, 0x10348, 0x0416, 0x05D0], size=int(rng.integers(1, 5))))
    if kind == 10:  # repeated symbol runs (a a a -> aa a)
        return pick()[:1] * int(rng.integers(2, 40))
    # long token
    n = int(rng.integers(33, 1500))
    return "".join(chr(int(c)) for c in rng.integers(33, 127, size=n))


def _gen_chunk(rng, wl, n, lo, hi, noise):
    k = rng.integers(lo, hi + 1, size=n)
    K = int(k.sum())
    idx = np.searchsorted(wl.cdf, rng.random(K), side="right").astype(np.int64)
    np.minimum(idx, len(wl.words) - 1, out=idx)
    blob, wstart, wlen = wl.blob, wl.wstart, wl.wlen
    if noise > 0 and K > 0:
        sel = np.nonzero(rng.random(K) < noise)[0]
        if len(sel):
            toks = [_noise_token(rng, wl).encode("utf-8") for _ in range(len(sel))]
            extra = np.frombuffer(b"".join(toks), dtype=np.uint8)
            elen = np.array([len(t) for t in toks], dtype=np.int64)
            estart = np.zeros(len(toks), dtype=np.int64)
            np.cumsum(elen[:-1], out=estart[1:])
            wstart = np.concatenate([wstart, estart + len(blob)])
            wlen = np.concatenate([wlen, elen])
            blob = np.concatenate([blob, extra])
            idx[sel] = len(wl.words) + np.arange(len(sel))
    # piece lengths incl. one separator after every word except the last of its document
    last = np.zeros(K, dtype=bool)
    ends = np.cumsum(k)
    last[ends[k > 0] - 1] = True
    L = wlen[idx] + (~last)
    out_start = np.zeros(K + 1, dtype=np.int64)
    np.cumsum(L, out=out_start[1:])
    total = int(out_start[-1])
    out = np.full(total, 0x20, dtype=np.uint8)
    # gather word bytes: out[out_start[j] + t] = blob[wstart[idx[j]] + t], t < wlen
    wl_j = wlen[idx]
    rep = np.repeat(np.arange(K, dtype=np.int64), wl_j)
    t = np.arange(int(wl_j.sum()), dtype=np.int64) - np.repeat(np.cumsum(wl_j) - wl_j, wl_j)
    out[out_start[rep] + t] = blob[wstart[idx[rep]] + t]
    doc_off = np.zeros(n + 1, dtype=np.int64)
    word_first = np.concatenate([[0], ends])
    doc_off[:] = out_start[word_first]
    return out, doc_off


def generate(seed, n, lo=3, hi=13, noise=0.0, wordlist=None, first_chunk=0):
    """n documents -> (uint8 bytes, int64 offsets[n+1]).  Chunk c uses default_rng(seed + first_chunk + c)."""
    wl = wordlist or default_wordlist()
    parts, offs, base, done, c = [], [np.zeros(1, dtype=np.int64)], 0, 0, 0
    while done < n:
        m = min(CHUNK, n - done)
        rng = np.random.default_rng(seed + first_chunk + c)
        b, o = _gen_chunk(rng, wl, m, lo, hi, noise)
        parts.append(b)
        offs.append(o[1:] + base)
        base += len(b)
        done += m
        c += 1
    if not parts:
        return np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    return np.concatenate(parts), np.concatenate(offs)


def unpack(b, off):
    """Packed form -> list[str] (for feeding the Python-level APIs / the reference)."""
    raw = b.tobytes()
    return [raw[off[i]:off[i + 1]].decode("utf-8", "surrogatepass") for i in range(len(off) - 1)]


def shard_range(n, rank, world):
    """Contiguous document range [lo, hi) of shard `rank`: no collective, shards concatenate."""
    return (n * rank) // world, (n * (rank + 1)) // world
