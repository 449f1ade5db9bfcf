"""1,048,576 sentence pairs, max_len=256, device-resident (BASELINE configs[2] per-GPU share): for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genz_tokenize_b200 import Tokenize, workload
n, W = 1 << 20, 256
tb, to = workload.generate(1234, n, 3, 13, 0.0)
pb, po = workload.generate(6234, n, 3, 13, 0.0)
dev = torch.device("cuda:0")
pad = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 16, dtype=np.uint8)])).to(dev)
d_t, d_to, d_p, d_po = pad(tb), torch.from_numpy(to).to(dev), pad(pb), torch.from_numpy(po).to(dev)
tok = Tokenize(devices=[0]); tok.set_option("max_chunk_bytes", 1 << 27)
out = {"input_ids": torch.empty((n, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n, W), dtype=torch.uint8, device=dev),
       "token_type_ids": torch.empty((n, W), dtype=torch.int8, device=dev), "row_len": torch.empty((n,), dtype=torch.int32, device=dev),
       "seq_len": torch.empty((n,), dtype=torch.int32, device=dev), "row_status": torch.empty((n,), dtype=torch.uint8, device=dev)}
for i in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tok.encode_device(d_t, d_to, d_p, d_po, max_len=W, out=out, text_bytes=len(tb), pair_bytes=len(pb)); b.record()
    torch.cuda.synchronize()
    print("pairs step %d: %.3f ms" % (i, a.elapsed_time(b)), flush=True)
