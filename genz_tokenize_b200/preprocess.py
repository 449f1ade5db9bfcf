"""Device versions of the reference's text normalisers (genz_tokenize/preprocess.py) -- the optional step a user
runs before `Tokenize`.  Same names and results; each call goes through `genztok_preprocess` (CUDA byte-filter
kernels, csrc/prep.cuh).  `*_batch` variants take a list of strings (or the packed form) and return a list.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .tokenizer import Tokenize, pack_strings

REMOVE_HTML, CONVERT_UNICODE, REMOVE_PUNCTUATIONS, REMOVE_EMOJI, REMOVE_URL = range(5)

_tok = None


def _ctx():
    global _tok
    if _tok is None:
        _tok = Tokenize()
    return _tok


def run_batch(op, texts, tok=None):
    """One normaliser over many documents; returns list[str]."""
    tok = tok or _ctx()
    tb, to = texts if isinstance(texts, tuple) else pack_strings(texts)
    tb = np.ascontiguousarray(tb, dtype=np.uint8)
    to = np.ascontiguousarray(to, dtype=np.int64)
    n = len(to) - 1
    out = L.Text()
    rc = tok._lib.genztok_preprocess(tok._h, int(op), tb.ctypes.data, to.ctypes.data, n, C.byref(out))
    if rc:
        tok._err(rc, "genztok_preprocess")
    try:
        off = np.ctypeslib.as_array(out.off, shape=(n + 1,)).copy()
        raw = C.string_at(out.bytes, int(out.total)) if out.total else b""
    finally:
        tok._lib.genztok_free_text(tok._h, C.byref(out))
    return [raw[off[i]:off[i + 1]].decode("utf-8", "surrogatepass") for i in range(n)]


def run_device(op, d_text, d_text_off, tok=None):
    """One normaliser over documents that are already on the GPU (torch uint8 bytes + int64 offsets[n+1]); returns
    (uint8 bytes padded for the encoder's 16-byte loads, int64 offsets, byte count) on the same device, written on torch's
    current stream -- ready for `Tokenize.encode_device` without a trip through host memory."""
    import torch
    tok = tok or _ctx()
    n = d_text_off.numel() - 1
    dev = d_text.device
    st = tok._torch_stream(dev)
    off = torch.empty((n + 1,), dtype=torch.int64, device=dev)
    total = C.c_int64()
    rc = tok._lib.genztok_preprocess_device(tok._h, 0, int(op), d_text.data_ptr(), d_text_off.data_ptr(), n, off.data_ptr(), None, C.byref(total), st)
    if rc:
        tok._err(rc, "genztok_preprocess_device")
    nb = int(total.value)
    out = torch.zeros((nb + (-nb) % 16 + 32,), dtype=torch.uint8, device=dev)
    rc = tok._lib.genztok_preprocess_device(tok._h, 0, int(op), d_text.data_ptr(), d_text_off.data_ptr(), n, off.data_ptr(), out.data_ptr(), None, st)
    if rc:
        tok._err(rc, "genztok_preprocess_device")
    return out, off, nb


def _one(op, txt):
    if not isinstance(txt, str):
        raise TypeError("expected string or bytes-like object, got %r" % type(txt).__name__)
    return run_batch(op, [txt])[0]


def remove_html(txt: str):
    '''Remove html tags (preprocess.py:5-9).'''
    return _one(REMOVE_HTML, txt)


def convert_unicode(txt: str):
    '''Composed (base letter + combining tone mark) -> precomposed Vietnamese letters (preprocess.py:30-36).'''
    return _one(CONVERT_UNICODE, txt)


def remove_punctuations(txt: str):
    '''Drop every character of string.punctuation (preprocess.py:39-44).'''
    return _one(REMOVE_PUNCTUATIONS, txt)


def remove_emoji(txt: str):
    '''Remove emoji and collapse whitespace (preprocess.py:47-72).'''
    return _one(REMOVE_EMOJI, txt)


def remove_URL(txt: str):
    '''Remove http... up to the next whitespace (preprocess.py:75-80).'''
    return _one(REMOVE_URL, txt)


def remove_html_batch(texts): return run_batch(REMOVE_HTML, texts)
def convert_unicode_batch(texts): return run_batch(CONVERT_UNICODE, texts)
def remove_punctuations_batch(texts): return run_batch(REMOVE_PUNCTUATIONS, texts)
def remove_emoji_batch(texts): return run_batch(REMOVE_EMOJI, texts)
def remove_URL_batch(texts): return run_batch(REMOVE_URL, texts)


def vncore_tokenize(text, vncore):
    '''Word-segment with an external VnCoreNLP object (preprocess.py:83-89): host-side glue around `vncore.tokenize`.'''
    sentences = vncore.tokenize(text)
    return ' '.join(' '.join(' '.join(words) for words in sentences).split())
