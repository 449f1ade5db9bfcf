"""Floor of the output write pattern: how fast can the [n,256] planes be written in the pieces the pipeline uses
(staged columns 0..KR-1 by one kernel, pad columns KR..255 by another) compared with whole rows?  torch fill kernels, CUDA events."""
import json, sys, torch
n, W = 1 << 20, 256
dev = torch.device("cuda:0")
ids = torch.empty((n, W), dtype=torch.int32, device=dev)
mask = torch.empty((n, W), dtype=torch.uint8, device=dev)
tt = torch.empty((n, W), dtype=torch.int8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=10):
    for _ in range(3):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / reps
res = {}
for kr in (32, 64, 128):
    def real(kr=kr):
        ids[:, :kr].fill_(7); mask[:, :kr].fill_(1); tt[:, :kr].fill_(1)
    def pads(kr=kr):
        ids[:, kr:].fill_(0); mask[:, kr:].fill_(0); tt[:, kr:].fill_(0)
    r, p = timed(real), timed(pads)
    res["kr%d" % kr] = {"real_ms": r, "real_gbs": n * kr * 6 / r / 1e6, "pads_ms": p, "pads_gbs": n * (W - kr) * 6 / p / 1e6}
def full():
    ids.fill_(0); mask.fill_(0); tt.fill_(0)
f = timed(full)
res["full"] = {"ms": f, "gbs": n * W * 6 / f / 1e6}
def ids_only():
    ids.fill_(0)
f = timed(ids_only)
res["ids_full"] = {"ms": f, "gbs": n * W * 4 / f / 1e6}
print(json.dumps(res))
