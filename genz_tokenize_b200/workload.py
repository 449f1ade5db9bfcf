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


# ---------------------------------------------------------------------------------------------------------------------
# Counter-based generator (the 100 M-document workloads of BASELINE configs[2]/[3]).
#
# generate() above draws from numpy's PCG64 stream: 3 s per million documents on one core -- ten minutes for the
# 200 M sentences of configs[2].  generate_hashed() produces the same kind of text (k ~ U{lo..hi} words per document
# sampled in proportion to the vocab counts, one ASCII space between words, a `noise` fraction of adversarial words)
# from a hash of (seed, side, document index, word index), so that ANY range of documents can be produced
# independently -- by any rank, on the host (here, numpy, vectorised) or on the device (csrc/synth.cuh, the same
# arithmetic: tests/test_gpu_parity.py::test_synth_device_matches_host).  The sharded bench uses the device form;
# the CPU arms and the parity tests use this one.
# ---------------------------------------------------------------------------------------------------------------------
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1, _M2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)
_K_LEN, _K_NOISE = np.uint64(0xA5A5A5A5A5A5A5A5), np.uint64(0xC3C3C3C3C3C3C3C3)
_ALPHABET = b"abcdefghijklmnopqrstuvwxyz0123456789"


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * _M1
        x = x ^ (x >> np.uint64(27))
        x = x * _M2
        x = x ^ (x >> np.uint64(31))
    return x


def synth_extras():
    """The glue pieces of the noise words: punctuation that only the vocab knows, '\\n', exotic whitespace and its
    look-alikes, literal markers and special tokens, non-BMP text, symbol runs, long tokens."""
    ex = list(_PUNCT) + ["\n"] + _EXOTIC_WS + _NOT_WS + ["</w>", "@@", "@@ ", "</s>", "<s>", "<pad>", "<unk>", " </s> ", " <s> ", " <pad> ",
                                                          "中文", "\U0001F600", "\U0001F4A9\U00010348", "Жא"]
    ex += ["a" * 7, "n" * 39, "à" * 5, "".join(chr(33 + (i * 7) % 94) for i in range(40)), "".join(chr(33 + (i * 11) % 94) for i in range(1200))]
    return [e.encode("utf-8") for e in ex]


class SynthTables:
    """Word list + integer CDF + extras in the packed arrays both implementations read."""

    def __init__(self, wordlist=None):
        wl = wordlist or default_wordlist()
        self.nw = len(wl.words)
        self.wblob = np.ascontiguousarray(wl.blob)
        self.wstart = wl.wstart.astype(np.uint32)
        self.wlen = wl.wlen.astype(np.uint32)
        c = np.floor(wl.cdf * 4294967296.0)
        c = np.minimum(c, 4294967295.0).astype(np.uint32)
        c[-1] = np.uint32(0xFFFFFFFF)
        self.cdf32 = c
        ex = synth_extras()
        self.ne = len(ex)
        self.eblob = np.frombuffer(b"".join(ex), dtype=np.uint8)
        self.elen = np.array([len(e) for e in ex], dtype=np.uint32)
        self.estart = np.zeros(len(ex), dtype=np.uint32)
        np.cumsum(self.elen[:-1], out=self.estart[1:])
        self.alphabet = np.frombuffer(_ALPHABET, dtype=np.uint8)


_default_st = None


def default_synth_tables():
    global _default_st
    if _default_st is None:
        _default_st = SynthTables()
    return _default_st


def noise_threshold(noise):
    return int(min(max(float(noise), 0.0), 1.0) * 4294967296.0) if noise < 1.0 else 0xFFFFFFFF


def generate_hashed(seed, doc0, n, side=0, lo=3, hi=13, noise=0.0, tables=None):
    """Documents [doc0, doc0 + n) of side `side` (0 = text, 1 = pair_text) -> (uint8 bytes, int64 offsets[n+1]).

    Per document d:  D = mix64(seed * GOLD + 2 d + side);  k = lo + ((mix64(D ^ K_LEN) >> 32) * (hi - lo + 1) >> 32) words.
    Word j:  R = mix64(D + (j + 1) * GOLD);  vocab word = upper_bound(cdf32, R >> 32);  noisy iff (R & 0xFFFFFFFF) < noise * 2^32.
    Noisy word:  Q = mix64(R ^ K_NOISE), second word b = upper_bound(cdf32, Q >> 32), extra e = ((Q >> 8) & 0xFFFFFF) % n_extras,
    kind = Q & 3:  0 word+word(b) | 1 word+extra+word(b) | 2 random [a-z0-9]{1..16} | 3 word+extra.
    """
    T = tables or default_synth_tables()
    n = int(n)
    if n <= 0:
        return np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    with np.errstate(over="ignore"):
        d = np.arange(n, dtype=np.uint64) + np.uint64(doc0)
        D = _mix64(np.uint64(seed) * _GOLD + np.uint64(2) * d + np.uint64(side))
        k = (np.uint64(lo) + (((_mix64(D ^ _K_LEN) >> np.uint64(32)) * np.uint64(hi - lo + 1)) >> np.uint64(32))).astype(np.int64)
        K = int(k.sum())
        doc_of = np.repeat(np.arange(n, dtype=np.int64), k)
        first = np.cumsum(k) - k
        j = np.arange(K, dtype=np.int64) - np.repeat(first, k)
        R = _mix64(D[doc_of] + (j.astype(np.uint64) + np.uint64(1)) * _GOLD)
    idx = np.minimum(np.searchsorted(T.cdf32, (R >> np.uint64(32)).astype(np.uint32), side="right"), T.nw - 1).astype(np.int64)
    # every word is up to three pieces out of one blob: vocab words, extras, random strings
    blob_parts = [T.wblob, T.eblob]
    e_base = len(T.wblob)
    p_start = np.zeros((K, 3), dtype=np.int64)
    p_len = np.zeros((K, 3), dtype=np.int64)
    p_start[:, 0] = T.wstart[idx]
    p_len[:, 0] = T.wlen[idx]
    thr = noise_threshold(noise)
    if thr > 0 and K > 0:
        noisy = np.nonzero((R & np.uint64(0xFFFFFFFF)) < np.uint64(thr))[0]
        if len(noisy):
            Q = _mix64(R[noisy] ^ _K_NOISE)
            kind = (Q & np.uint64(3)).astype(np.int64)
            b = np.minimum(np.searchsorted(T.cdf32, (Q >> np.uint64(32)).astype(np.uint32), side="right"), T.nw - 1).astype(np.int64)
            e = (((Q >> np.uint64(8)) & np.uint64(0xFFFFFF)) % np.uint64(T.ne)).astype(np.int64)
            k0, k1, k2, k3 = kind == 0, kind == 1, kind == 2, kind == 3
            # kind 0: word + word(b)
            p_start[noisy[k0], 1] = T.wstart[b[k0]]; p_len[noisy[k0], 1] = T.wlen[b[k0]]
            # kind 1: word + extra + word(b)
            p_start[noisy[k1], 1] = e_base + T.estart[e[k1]].astype(np.int64); p_len[noisy[k1], 1] = T.elen[e[k1]]
            p_start[noisy[k1], 2] = T.wstart[b[k1]]; p_len[noisy[k1], 2] = T.wlen[b[k1]]
            # kind 3: word + extra
            p_start[noisy[k3], 1] = e_base + T.estart[e[k3]].astype(np.int64); p_len[noisy[k3], 1] = T.elen[e[k3]]
            # kind 2: a random string replaces the word: 1..16 characters, character i = alphabet[byte i of mix64(Q + 1 + i / 8) % 36]
            if k2.any():
                with np.errstate(over="ignore"):
                    q2 = Q[k2]
                    ln = (np.uint64(1) + ((q2 >> np.uint64(2)) & np.uint64(15))).astype(np.int64)
                    h0, h1 = _mix64(q2 + np.uint64(1)), _mix64(q2 + np.uint64(2))
                sh = (np.arange(8, dtype=np.uint64) * np.uint64(8))[None, :]
                by = np.concatenate([(h0[:, None] >> sh) & np.uint64(0xFF), (h1[:, None] >> sh) & np.uint64(0xFF)], axis=1).astype(np.int64)
                chars = T.alphabet[by % 36]                               # [m, 16]
                r_base = e_base + len(T.eblob)
                blob_parts.append(np.ascontiguousarray(chars).reshape(-1))
                rows = noisy[k2]
                p_start[rows, 0] = r_base + 16 * np.arange(len(rows), dtype=np.int64); p_len[rows, 0] = ln
    blob = np.concatenate(blob_parts) if len(blob_parts) > 2 else np.concatenate(blob_parts[:2])
    last = np.zeros(K, dtype=bool)
    ends = np.cumsum(k)
    last[ends[k > 0] - 1] = True
    wl_total = p_len.sum(axis=1)
    L = wl_total + (~last)
    out_start = np.zeros(K + 1, dtype=np.int64)
    np.cumsum(L, out=out_start[1:])
    out = np.full(int(out_start[-1]), 0x20, dtype=np.uint8)
    pos = out_start[:-1].copy()
    for c in range(3):
        pl = p_len[:, c]
        sel = np.nonzero(pl)[0]
        if len(sel):
            pls = pl[sel]
            rep = np.repeat(sel, pls)
            t = np.arange(int(pls.sum()), dtype=np.int64) - np.repeat(np.cumsum(pls) - pls, pls)
            out[pos[rep] + t] = blob[p_start[rep, c] + t]
        pos += pl
    doc_off = out_start[np.concatenate([[0], ends])]
    return out, doc_off.astype(np.int64)


def plane_digest(ids, mask, tt=None, row0=0):
    """Host restatement of csrc/synth.cuh::k_plane_digest (numpy): order-independent 64-bit digest of [n, W] planes."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    n, W = ids.shape
    with np.errstate(over="ignore"):
        rk = _mix64((np.arange(n, dtype=np.uint64) + np.uint64(row0)) * _GOLD)[:, None]
        i = np.arange(W, dtype=np.uint64)[None, :]
        total = _mix64((rk + i * _GOLD) ^ ids.view(np.uint32).astype(np.uint64)).sum(dtype=np.uint64)
        i4 = np.arange(W // 4, dtype=np.uint64)[None, :]
        for p, plane in ((1, mask), (2, tt)):
            if plane is None:
                continue
            w = np.ascontiguousarray(plane).view(np.uint8).reshape(n, W).view(np.uint32).astype(np.uint64)
            total = total + _mix64((rk + ((np.uint64(p) << np.uint64(32)) + i4) * _GOLD) ^ w).sum(dtype=np.uint64)
    return int(total)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: a large custom vocab / merge table and long documents with low word reuse (SURVEY.md 8 d2 "C5":
# none is bundled with the reference, so one is synthesised from the bundled word list).
# ---------------------------------------------------------------------------------------------------------------------
def build_custom_model(directory, n_words=120000, seed=77):
    """Writes vocab.txt / bpe.codes for ~n_words sampled words plus their '_'-joined pairs, with a left-to-right merge chain
    per word (so every word is one token after len-1 merges).  Returns (vocab_path, codes_path, words)."""
    import os
    rng = np.random.default_rng(seed)
    wl = default_wordlist()
    base = [wl.words[int(i)] for i in rng.integers(0, len(wl.words), size=n_words)]
    words = list(dict.fromkeys(base + [a + "_" + b for a, b in zip(base[::2], base[1::2])]))
    merges, seen, vocab = [], set(), {}
    for w in words:
        syms = list(w[:-1]) + [w[-1] + "</w>"]
        cur = syms[0]
        for s in syms[1:]:
            if (cur, s) not in seen:
                seen.add((cur, s))
                merges.append("%s %s" % (cur, s))
            cur = cur + s
            vocab.setdefault(cur.replace("</w>", "") + ("" if cur.endswith("</w>") else "@@"), 1)
        vocab[w] = 1
    vp, mp = os.path.join(directory, "vocab.txt"), os.path.join(directory, "bpe.codes")
    with open(vp, "w", encoding="utf-8") as f:
        f.write("".join("%s %d\n" % (k, v) for k, v in vocab.items()))
    with open(mp, "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n" + "\n".join(merges) + "\n")
    return vp, mp, words


def long_documents(words, n_docs, lo=6000, hi=9000, seed=5):
    """n_docs documents of lo..hi words sampled UNIFORMLY from `words` (low reuse) -> packed (bytes, offsets)."""
    rng = np.random.default_rng(seed)
    enc = [w.encode("utf-8") for w in words]
    wlen = np.array([len(e) for e in enc], dtype=np.int64)
    wstart = np.zeros(len(enc), dtype=np.int64)
    np.cumsum(wlen[:-1], out=wstart[1:])
    blob = np.frombuffer(b"".join(enc), dtype=np.uint8)
    k = rng.integers(lo, hi, size=n_docs)
    K = int(k.sum())
    idx = rng.integers(0, len(enc), size=K)
    last = np.zeros(K, dtype=bool)
    ends = np.cumsum(k)
    last[ends - 1] = True
    L = wlen[idx] + (~last)
    out_start = np.zeros(K + 1, dtype=np.int64)
    np.cumsum(L, out=out_start[1:])
    out = np.full(int(out_start[-1]), 0x20, dtype=np.uint8)
    wl_j = wlen[idx]
    rep = np.repeat(np.arange(K, dtype=np.int64), wl_j)
    t = np.arange(int(wl_j.sum()), dtype=np.int64) - np.repeat(np.cumsum(wl_j) - wl_j, wl_j)
    out[out_start[rep] + t] = blob[wstart[idx[rep]] + t]
    return out, out_start[np.concatenate([[0], ends])].astype(np.int64)
