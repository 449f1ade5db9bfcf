"""Sentence pairs at max_len 256 (BASELINE configs[2], one 1,048,576-pair chunk) -- or, with SINGLES=1 in the environment, the single
sentences of configs[1] at max_len 128: step and kernel times for a few option sets.
    python tools/pairs_variants.py "" opt=val,opt=val ..."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genz_tokenize_b200 import Tokenize, workload

SINGLES = os.environ.get("SINGLES", "") not in ("", "0")
n, W = 1 << 20, (128 if SINGLES else 256)
dev = torch.device("cuda:0")
tb, to = workload.generate(1234, n, 3, 13, 0.0)
pb, po = workload.generate(6234, n, 3, 13, 0.0)
pad16 = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 16, dtype=np.uint8)])).to(dev)
d_t, d_to, d_p, d_po = pad16(tb), torch.from_numpy(to).to(dev), pad16(pb), torch.from_numpy(po).to(dev)
out = {"input_ids": torch.empty((n, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n, W), dtype=torch.uint8, device=dev),
       "token_type_ids": torch.empty((n, W), dtype=torch.int8, device=dev), "row_len": torch.empty((n,), dtype=torch.int32, device=dev),
       "seq_len": torch.empty((n,), dtype=torch.int32, device=dev), "row_status": torch.empty((n,), dtype=torch.uint8, device=dev)}
if SINGLES:
    out = {k: out[k] for k in ("input_ids", "attention_mask", "row_len")}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for opts in sys.argv[1:] or [""]:
    tok = Tokenize(devices=[0])
    tok.set_option("max_chunk_bytes", 1 << 27)
    for kv in opts.split(","):
        if "=" in kv:
            tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    step = (lambda: tok.encode_device(d_t, d_to, max_len=W, out=out, text_bytes=len(tb))) if SINGLES else \
           (lambda: tok.encode_device(d_t, d_to, d_p, d_po, max_len=W, out=out, text_bytes=len(tb), pair_bytes=len(pb)))
    for _ in range(3):
        flush.zero_(); step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        flush.zero_(); a.record(); step(); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / 10
    tok.set_profiling(True); tok.profile_report(reset=True)
    for _ in range(5):
        flush.zero_(); step()
    torch.cuda.synchronize()
    prof = tok.profile_report(reset=True)
    print(opts or "default", "ms/step %.4f" % ms, {k: round(v["ms"] / 5, 4) for k, v in prof.items() if v["ms"] / 5 > 0.002})
    del tok
