"""Small end-to-end case for compute-sanitizer (memcheck): fixed + ragged + pairs + spans + decode + helpers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genz_tokenize_b200 import Tokenize, workload
tok = Tokenize()
tok.set_option("max_chunk_bytes", 1 << 16)
tok.set_option("chunk_rows", 300)
t = workload.generate(31, 1500, 0, 14, 0.15)
p = workload.generate(32, 1500, 0, 14, 0.15)
for kw in (dict(max_len=32), dict(max_len=7), dict(), dict(max_len=20, padding=False)):
    a = tok.encode_batch(t, p, **kw)
    b = tok.encode_batch(t, None, **kw)
c = tok.encode_batch(t, p, return_offset=True)
d = tok.decode_batch(a["input_ids"], a["row_off"])
e = tok.decode_batch(tok.encode_batch(t, p, max_len=48)["input_ids"])
print(tok("sinh_viên công_nghệ", "hello", max_len=10), tok.bpe("hello\n"), tok.get_sequence_id([1, 5, 2, 2, 6, 2]), tok.decode([1, 770, 2]))
long_word = "x" * 70000 + " " + "é" * 3000
tok2 = Tokenize()
tok2.set_option("max_chunk_bytes", 1 << 20)
print(len(tok2(long_word)["input_ids"]), len(tok2(long_word, long_word)["input_ids"]))
try:
    tok2(long_word, long_word, max_len=64)
except ValueError as exc:
    print("ValueError as in the reference:", exc)
print("sanitize case ok", a["real_tokens"], len(d), len(e))
