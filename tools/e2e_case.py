"""End to end (host buffers) on 2,097,152 pairs at max_len 256: ms per call for a few option sets.
    python tools/e2e_case.py [opts ...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genz_tokenize_b200 import Tokenize, workload
n, W = 1 << 21, 256
dev = torch.device("cuda:0")
gen = Tokenize(devices=[0])
ta, oa, na = gen.synth_device(1234, 0, n, 0, device=dev)
tb, ob, nb = gen.synth_device(1234, 0, n, 1, device=dev)
import ctypes as C
from genz_tokenize_b200 import _lib as L
lib = L.load()
def pinned(t, nb):                                   # the caller's text in pinned memory, as bench.py's e2e leg has it
    hp = lib.genztok_host_alloc(nb + 64)
    buf = np.frombuffer((C.c_uint8 * nb).from_address(hp), dtype=np.uint8)
    buf[:] = t[:nb].cpu().numpy()
    return buf
A = (pinned(ta, na), oa.cpu().numpy()); B = (pinned(tb, nb), ob.cpu().numpy())
del gen
for opts in sys.argv[1:] or [""]:
    tok = Tokenize(devices=[0])
    tok.set_option("max_chunk_bytes", 1 << 27)
    for kv in opts.split(","):
        if "=" in kv:
            tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    ts = []
    for i in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        be = tok.encode_batch(A, B, max_len=W, sequence_id=False)
        x = int(be["row_len"][-1]); dt = time.perf_counter() - t0
        ts.append(dt * 1e3); d2h = be.d2h_bytes; toks = be["real_tokens"]; del be
    tok.set_profiling(True); tok.profile_report(reset=True)
    be = tok.encode_batch(A, B, max_len=W, sequence_id=False); del be
    prof = tok.profile_report(reset=True); tok.set_profiling(False)
    print("   kernels (ms per call):", {k: round(v["ms"], 2) for k, v in prof.items() if v["ms"] > 0.05}, flush=True)
    print(opts or "default", "ms per call:", [round(t, 1) for t in ts], "d2h MB %.0f" % (d2h / 1e6), "-> %.2f G tokens/s, d2h %.1f GB/s" % (toks / min(ts[2:]) / 1e6, d2h / min(ts[2:]) / 1e6), flush=True)
    del tok
