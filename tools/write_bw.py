"""Pure-write bandwidth of the GPU (memset-style fill), for the roofline of a write-dominated kernel."""
import torch, json
dev = torch.device("cuda:0")
res = {}
for mb in (671, 1342, 4096):
    x = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for name, fn in (("zero_", lambda: x.zero_()), ("fill_int32", lambda: x.view(torch.int32).fill_(7))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        t = min(a.elapsed_time(b) for a, b in ev)
        res["%s_%dMiB" % (name, mb)] = {"ms": t, "GB/s": (mb << 20) / t / 1e6}
    del x
print(json.dumps(res, indent=1))
