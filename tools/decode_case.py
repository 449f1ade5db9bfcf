"""Decode of padded rows on one GPU: per-kernel times (library-side CUDA events)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genz_tokenize_b200 import Tokenize, workload
n = 1 << 20
dev = torch.device("cuda:0")
tok = Tokenize(devices=[0])
tok.set_option("max_chunk_bytes", 1 << 28)
for kv in os.environ.get("GENZTOK_OPTIONS", "").split(","):
    if "=" in kv:
        tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))
tb, to = workload.generate(1234, n, 3, 13, 0.0)
pad16 = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 16, dtype=np.uint8)])).to(dev)
d_t, d_to = pad16(tb), torch.from_numpy(to).to(dev)
ring = torch.empty((1 << 31,), dtype=torch.uint8, device=dev)     # the caller's text buffer, reused: sizes + text without a host read in between
for W in (128, 256):
    out = {"input_ids": torch.empty((n, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n, W), dtype=torch.uint8, device=dev),
           "row_len": torch.empty((n,), dtype=torch.int32, device=dev)}
    tok.encode_device(d_t, d_to, max_len=W, out=out, text_bytes=len(tb))
    torch.cuda.synchronize()
    for _ in range(2):
        txt, off = tok.decode_device(out["input_ids"], out=ring)
    torch.cuda.synchronize()
    tok.set_profiling(True); tok.profile_report(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); txt, off = tok.decode_device(out["input_ids"], out=ring); b.record()
    torch.cuda.synchronize()
    prof = tok.profile_report(reset=True)
    tok.set_profiling(False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:                                     # the figure: without the per-kernel events
        a.record(); txt, off = tok.decode_device(out["input_ids"], out=ring); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        tok.decode_device(out["input_ids"], out=ring, sync=False)
    b.record(); torch.cuda.synchronize()
    print("W=%d streamed without host reads: %.3f ms per batch" % (W, a.elapsed_time(b) / 8))
    alg = 4 * n * W + int(txt.numel()) + 16 * n
    print("W=%d ms %.3f alg GB/s %.0f text MB %.0f" % (W, ms, alg / ms / 1e6, txt.numel() / 1e6), {k: round(v["ms"] / 5, 4) for k, v in prof.items()})
    import hashlib
    print("  sha", hashlib.sha256(txt[:1 << 24].cpu().numpy().tobytes()).hexdigest()[:16], int(off[-1].item()))

# sentence pairs of BASELINE configs[2] at max_len 256 (~40 ids in front of the pad run): the warp-per-row write kernel against the
# lane-per-junction kernel with 512-byte junctions
ta, oa, na = tok.synth_device(1234, 0, n, 0, device=dev)
tb2, ob2, nb2 = tok.synth_device(1234, 0, n, 1, device=dev)
outp = tok.encode_device(ta, oa, tb2, ob2, max_len=256, text_bytes=na, pair_bytes=nb2)
torch.cuda.synchronize()
for mode in (1, 3, 0):
    tok.set_option("decode_write", mode)
    for _ in range(2):
        txt, off = tok.decode_device(outp["input_ids"], out=ring)
    torch.cuda.synchronize()
    tok.set_profiling(True); tok.profile_report(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); txt, off = tok.decode_device(outp["input_ids"], out=ring); b.record()
    torch.cuda.synchronize()
    prof = tok.profile_report(reset=True)
    tok.set_profiling(False)
    ms = sum(a.elapsed_time(b) for a, b in ev) / 5
    alg = 4 * n * 256 + int(txt.numel()) + 16 * n
    print("pairs W=256 decode_write=%d ms %.3f alg GB/s %.0f text MB %.0f" % (mode, ms, alg / ms / 1e6, txt.numel() / 1e6), {k: round(v["ms"] / 5, 4) for k, v in prof.items()})
    print("  sha", hashlib.sha256(txt[:1 << 26].cpu().numpy().tobytes()).hexdigest()[:16], int(off[-1].item()))
tok.set_option("decode_write", 0)
del outp, ta, tb2, txt, off

# mask-trimmed ragged rows of the same sentences (BASELINE configs[3], the other way to hand the rows over): a thread per id,
# and the warp-per-row kernels for comparison
x = torch.zeros((1 << 28,), dtype=torch.float32, device=dev)    # what a read-only pass over 1 GiB reaches on this GPU (torch's reduction)
for _ in range(3):
    x.sum()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for a, b in ev:
    a.record(); x.sum(); b.record()
torch.cuda.synchronize()
print("read-only probe: torch sum over 1 GiB: %.0f GB/s" % (x.numel() * 4 / min(a.elapsed_time(b) for a, b in ev) / 1e6))
del x
rag = tok.encode_batch((tb, to))
r_ids, r_off = torch.from_numpy(rag["input_ids"]).to(dev), torch.from_numpy(rag["row_off"]).to(dev)
for mode in (0, 1):
    tok.set_option("no_token_decode", mode)
    for _ in range(2):
        txt, off = tok.decode_device(r_ids, r_off)
    torch.cuda.synchronize()
    tok.set_profiling(True); tok.profile_report(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); txt, off = tok.decode_device(r_ids, r_off); b.record()
    torch.cuda.synchronize()
    prof = tok.profile_report(reset=True)
    tok.set_profiling(False)
    ms = sum(a.elapsed_time(b) for a, b in ev) / 5
    alg = 4 * int(r_ids.numel()) + int(txt.numel()) + 16 * n
    print("ragged no_token_decode=%d ms %.3f alg GB/s %.0f ids %d text MB %.0f" % (mode, ms, alg / ms / 1e6, r_ids.numel(), txt.numel() / 1e6), {k: round(v["ms"] / 5, 4) for k, v in prof.items()})
    print("  sha", hashlib.sha256(txt.cpu().numpy().tobytes()).hexdigest()[:16], int(off[-1].item()))
