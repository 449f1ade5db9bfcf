"""Cold word cache and noisy text on one 1,048,576-pair chunk (max_len 256): step time and per-kernel times.
    python tools/cold_case.py [opt=val,...]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genz_tokenize_b200 import Tokenize

n, W = 1 << 20, 256
dev = torch.device("cuda:0")
tok = Tokenize(devices=[0])
tok.set_option("max_chunk_bytes", 1 << 27)
for kv in (sys.argv[1] if len(sys.argv) > 1 else "").split(","):
    if "=" in kv:
        tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))
out = {"input_ids": torch.empty((n, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n, W), dtype=torch.uint8, device=dev),
       "token_type_ids": torch.empty((n, W), dtype=torch.int8, device=dev), "row_len": torch.empty((n,), dtype=torch.int32, device=dev),
       "seq_len": torch.empty((n,), dtype=torch.int32, device=dev), "row_status": torch.empty((n,), dtype=torch.uint8, device=dev)}


def chunk(doc0, noise):
    a = tok.synth_device(1234, doc0, n, 0, 3, 13, noise, device=dev)
    b = tok.synth_device(1234, doc0, n, 1, 3, 13, noise, device=dev)
    return a, b


def run(c, label, profile=True):
    (ta, oa, na), (tb, ob, nb) = c
    if profile:
        tok.set_profiling(True); tok.profile_report(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tok.encode_device(ta, oa, tb, ob, max_len=W, out=out, text_bytes=na, pair_bytes=nb)
    e1.record()
    torch.cuda.synchronize()
    prof = tok.profile_report(reset=True) if profile else {}
    tok.set_profiling(False)
    print(label, "ms %.3f" % e0.elapsed_time(e1), {k: round(v["ms"], 3) for k, v in prof.items() if v["ms"] > 0.02}, "tokens", int(out["row_len"].sum()), flush=True)


clean0, clean1 = chunk(0, 0.0), chunk(n, 0.0)
noisy0, noisy1, noisy2 = chunk(2 * n, 0.01), chunk(3 * n, 0.01), chunk(4 * n, 0.01)
run(clean0, "first call (allocations)   ")
for rep in range(2):
    tok.cache_reset()
    run(clean0, "cold, clean text           ")
    run(clean1, "warm vocabulary, new chunk ")
    run(clean1, "same chunk again           ")
tok.cache_reset()
run(noisy0, "cold, 1% noise             ")
run(noisy1, "warm vocab, fresh 1% noise ")
run(noisy2, "warm vocab, fresh 1% noise ")
run(noisy2, "same noisy chunk again     ")
tok.check_errors(dev)
