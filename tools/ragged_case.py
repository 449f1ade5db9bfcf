"""The reference's default call shape (max_len=None: no padding, no truncation) over 1,048,576 single sentences through the host API:
wall clock of the call with the text in pageable and in pinned memory, and the kernels' share.
    python tools/ragged_case.py [opt=val,...]"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genz_tokenize_b200 import Tokenize, workload, _lib as L
n = 1 << 20
tok = Tokenize(devices=[0])
for kv in (sys.argv[1] if len(sys.argv) > 1 else "").split(","):
    if "=" in kv:
        tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))
tb, to = workload.generate(1234, n, 3, 13, 0.0)
lib = L.load()
hp = lib.genztok_host_alloc(len(tb) + 64)
pinned = np.frombuffer((C.c_uint8 * len(tb)).from_address(hp), dtype=np.uint8)
pinned[:] = tb
for label, text in (("pageable text", tb), ("pinned text", pinned)):
    for kw in (dict(), dict(return_offset=True)):
        ts = []
        for i in range(5):
            t0 = time.perf_counter()
            be = tok.encode_batch((text, to), **kw)
            dt = time.perf_counter() - t0
            ts.append(dt * 1e3); toks = int(be["real_tokens"]); del be
        tok.set_profiling(True); tok.profile_report(reset=True)
        be = tok.encode_batch((text, to), **kw); del be
        prof = tok.profile_report(reset=True); tok.set_profiling(False)
        print("%s %r: ms per call %s -> %.0f M tokens/s; kernels %.2f ms" % (label, kw, [round(t, 1) for t in ts], toks / min(ts[1:]) / 1e3, sum(v["ms"] for v in prof.values())), flush=True)
