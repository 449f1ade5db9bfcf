"""ctypes binding of oracle/libgenztok_oracle.so -- TEST INFRASTRUCTURE ONLY.

`Oracle` offers the reference's surface (tokenize.py:6-267) on top of the C restatement in
genztok_oracle.c: `__call__`, `decode`, `bpe`, `encoder_get`, ... plus batch forms that return
ragged numpy arrays so the CUDA path can be compared row by row.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgenztok_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "genztok_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgenztok_oracle.so"])
    return _SO


class _Encoded(C.Structure):
    _fields_ = [("n", C.c_int64),
                ("ids_off", C.POINTER(C.c_int64)), ("ids", C.POINTER(C.c_int32)), ("mask", C.POINTER(C.c_uint8)),
                ("seq_off", C.POINTER(C.c_int64)), ("seq", C.POINTER(C.c_int32)),
                ("tt_off", C.POINTER(C.c_int64)), ("tt", C.POINTER(C.c_int32)),
                ("span_off", C.POINTER(C.c_int64)), ("span", C.POINTER(C.c_int32)),
                ("status", C.POINTER(C.c_uint8))]


class _Text(C.Structure):
    _fields_ = [("n", C.c_int64), ("off", C.POINTER(C.c_int64)), ("bytes", C.POINTER(C.c_uint8))]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.gzo_create.restype = C.c_void_p
        L.gzo_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p)]
        L.gzo_destroy.argtypes = [C.c_void_p]
        L.gzo_last_error.restype = C.c_char_p
        L.gzo_vocab_size.restype = C.c_int64
        L.gzo_vocab_size.argtypes = [C.c_void_p]
        L.gzo_num_merges.restype = C.c_int64
        L.gzo_num_merges.argtypes = [C.c_void_p]
        L.gzo_encoder_get.restype = C.c_int32
        L.gzo_encoder_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.gzo_decoder_get.restype = C.c_int64
        L.gzo_decoder_get.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
        L.gzo_rank_get.restype = C.c_int32
        L.gzo_rank_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
        L.gzo_special_ids.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.gzo_bpe.restype = C.c_int64
        L.gzo_bpe.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
        L.gzo_encode.restype = C.c_int
        L.gzo_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                 C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Encoded)]
        L.gzo_free_encoded.argtypes = [C.POINTER(_Encoded)]
        L.gzo_sequence_id.restype = C.c_int64
        L.gzo_sequence_id.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.gzo_decode.restype = C.c_int
        L.gzo_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(_Text)]
        L.gzo_free_text.argtypes = [C.POINTER(_Text)]
        L.gzo_max_threads.restype = C.c_int
        L.gzo_preprocess.restype = C.c_int
        L.gzo_preprocess.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Text)]
        _lib = L
    return _lib


def pack_strings(strs):
    """list[str] -> (uint8 bytes, int64 offsets[n+1]) with the ABI's UTF-8 ('surrogatepass') form."""
    enc = [s.encode("utf-8", "surrogatepass") for s in strs]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum([len(e) for e in enc], out=off[1:])
    return np.frombuffer(b"".join(enc), dtype=np.uint8).copy(), off


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _opt(x):
    return None if x == -1 else int(x)


def preprocess(op, texts):
    """preprocess.py normalisers (0 remove_html, 1 convert_unicode, 2 remove_punctuations, 3 remove_emoji, 4 remove_URL)."""
    L = _load()
    tb, to = texts if isinstance(texts, tuple) else pack_strings(texts)
    n = len(to) - 1
    out = _Text()
    L.gzo_preprocess(int(op), tb.ctypes.data, to.ctypes.data, n, C.byref(out))
    try:
        o = _arr(out.off, n + 1, np.int64)
        b = _arr(out.bytes, int(o[-1]), np.uint8).tobytes()
    finally:
        L.gzo_free_text(C.byref(out))
    return [b[o[i]:o[i + 1]].decode("utf-8", "surrogatepass") for i in range(n)]


class Oracle:
    def __init__(self, vocab_file=None, bpe_file=None, specials=None):
        L = _load()
        if vocab_file is None or bpe_file is None:
            from genz_tokenize_b200.data import bundled_paths   # data files only, no compute
            v, b = bundled_paths()
            vocab_file, bpe_file = vocab_file or v, bpe_file or b
        sp = (C.c_char_p * 5)(*[(s.encode() if s is not None else None) for s in (specials or [None] * 5)])
        self._h = L.gzo_create(os.fsencode(vocab_file), os.fsencode(bpe_file), sp)
        if not self._h:
            msg = L.gzo_last_error().decode()
            if msg.startswith("FileNotFoundError"):
                raise FileNotFoundError(msg)
            raise ValueError(msg)
        self.unk_token = (specials[4] if specials and specials[4] else "<unk>")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.gzo_destroy(self._h)
            self._h = None

    # ---- introspection
    def vocab_size(self):
        return int(_lib.gzo_vocab_size(self._h))

    def num_merges(self):
        return int(_lib.gzo_num_merges(self._h))

    def special_ids(self):
        out = (C.c_int32 * 5)()
        _lib.gzo_special_ids(self._h, out)
        return list(out)

    def encoder_get(self, key):
        k = key.encode("utf-8", "surrogatepass")
        return _opt(_lib.gzo_encoder_get(self._h, k, len(k)))

    def decoder_get(self, i):
        buf = C.create_string_buffer(4096)
        n = _lib.gzo_decoder_get(self._h, int(i), buf, 4096)
        return None if n < 0 else buf.raw[:n].decode("utf-8", "surrogatepass")

    def rank_get(self, a, b):
        a, b = a.encode(), b.encode()
        return _opt(_lib.gzo_rank_get(self._h, a, len(a), b, len(b)))

    def bpe(self, token):
        w = token.encode("utf-8", "surrogatepass")
        cap = 8 * len(w) + 64
        buf = C.create_string_buffer(cap)
        n = _lib.gzo_bpe(self._h, w, len(w), buf, cap)
        return buf.raw[:n].decode("utf-8", "surrogatepass")

    # ---- batch forms (ragged numpy)
    def encode_batch(self, texts, pairs=None, max_len=None, padding=True, truncation=True, return_offset=False, threads=1):
        tb, to = texts if isinstance(texts, tuple) else pack_strings(texts)
        n = len(to) - 1
        if pairs is not None:
            pb, po = pairs if isinstance(pairs, tuple) else pack_strings(pairs)
            pbp, pop = pb.ctypes.data, po.ctypes.data
        else:
            pbp = pop = None
        out = _Encoded()
        rc = _lib.gzo_encode(self._h, tb.ctypes.data, to.ctypes.data, pbp, pop, n, int(max_len is not None),
                             int(max_len or 0), int(bool(padding)), int(bool(truncation)), int(bool(return_offset)),
                             int(threads), C.byref(out))
        if rc:
            raise ValueError("oracle: invalid UTF-8 input")
        try:
            ids_off = _arr(out.ids_off, n + 1, np.int64)
            seq_off = _arr(out.seq_off, n + 1, np.int64)
            tt_off = _arr(out.tt_off, n + 1, np.int64)
            span_off = _arr(out.span_off, n + 1, np.int64)
            res = dict(n=n, ids_off=ids_off, ids=_arr(out.ids, int(ids_off[-1]), np.int32),
                       mask=_arr(out.mask, int(ids_off[-1]), np.uint8), seq_off=seq_off,
                       seq=_arr(out.seq, int(seq_off[-1]), np.int32), tt_off=tt_off,
                       tt=_arr(out.tt, int(tt_off[-1]), np.int32), span_off=span_off,
                       span=_arr(out.span, 2 * int(span_off[-1]), np.int32).reshape(-1, 2),
                       status=_arr(out.status, n, np.uint8))
        finally:
            _lib.gzo_free_encoded(C.byref(out))
        return res

    def decode_batch(self, ids, off, threads=1):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = len(off) - 1
        out = _Text()
        _lib.gzo_decode(self._h, ids.ctypes.data, off.ctypes.data, n, int(threads), C.byref(out))
        try:
            o = _arr(out.off, n + 1, np.int64)
            b = _arr(out.bytes, int(o[-1]), np.uint8).tobytes()
        finally:
            _lib.gzo_free_text(C.byref(out))
        return [b[o[i]:o[i + 1]].decode("utf-8", "surrogatepass") for i in range(n)]

    # ---- the reference's single-call surface (tokenize.py:184-259, :137-139)
    def __call__(self, text, pair_text=None, max_len=None, padding=True, truncation=True, return_offset=False):
        r = self.encode_batch([text], None if pair_text is None else [pair_text], max_len, padding, truncation, return_offset)
        if r["status"][0]:
            raise ValueError("None is not in list")
        res = {}
        if return_offset:
            res["offset"] = [tuple(int(v) for v in p) for p in r["span"]]
        res["input_ids"] = r["ids"].tolist()
        res["attention_mask"] = r["mask"].tolist()
        if pair_text is not None:
            res["sequence_id"] = [None if v == -1 else int(v) for v in r["seq"]]
            res["token_type_ids"] = [None if v == -1 else int(v) for v in r["tt"]]
        return res

    def decode(self, ids):
        ids = np.asarray(list(ids), dtype=np.int64)
        ids = np.where((ids < -1) | (ids > 2**31 - 1), -1, ids).astype(np.int32)
        return self.decode_batch(ids, np.array([0, len(ids)], dtype=np.int64))[0]

    def sequence_id(self, ids, apply_token_type=False):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.zeros(max(len(a), 1), dtype=np.int32)
        ok = C.c_int(0)
        m = _lib.gzo_sequence_id(self._h, a.ctypes.data, len(a), int(apply_token_type), out.ctypes.data, C.byref(ok))
        if not ok.value:
            raise ValueError("None is not in list")
        return [None if v == -1 else int(v) for v in out[:m]]

    @staticmethod
    def max_threads():
        return int(_load().gzo_max_threads())
