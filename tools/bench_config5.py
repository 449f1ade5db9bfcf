"""BASELINE configs[4]: Tokenize.fromFile with a large custom vocab / merge table on long documents (max_len=4096,
heavy truncation, low word reuse): cold and warm step times with the per-kernel split.
    python tools/bench_config5.py [opt=val,...]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from genz_tokenize_b200 import Tokenize, workload

with tempfile.TemporaryDirectory() as td:
    t0 = time.time()
    vp, mp, words = workload.build_custom_model(td)
    tok = Tokenize.fromFile(vp, mp, devices=[0])
    tok.set_option("max_chunk_bytes", 1 << 28)
    for kv in (sys.argv[1] if len(sys.argv) > 1 else "").split(","):
        if "=" in kv:
            tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    print("custom model: %d words, built+loaded in %.1fs; word bytes: mean %.1f, >16: %.2f, >24: %.2f" % (
        len(words), time.time() - t0, np.mean([len(w.encode()) for w in words]), np.mean([len(w.encode()) > 16 for w in words]),
        np.mean([len(w.encode()) > 24 for w in words])), flush=True)
    n_docs, W = 1024, 4096
    blob, off = workload.long_documents(words, n_docs)
    dev = torch.device("cuda:0")
    d_text = torch.from_numpy(np.concatenate([blob, np.zeros(64, dtype=np.uint8)])).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    out = {"input_ids": torch.empty((n_docs, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n_docs, W), dtype=torch.uint8, device=dev),
           "row_len": torch.empty((n_docs,), dtype=torch.int32, device=dev)}

    def run(label):
        tok.set_profiling(True); tok.profile_report(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tok.encode_device(d_text, d_off, max_len=W, out=out, text_bytes=int(off[-1])); b.record()
        torch.cuda.synchronize()
        prof = tok.profile_report(reset=True); tok.set_profiling(False)
        ms = a.elapsed_time(b)
        print("%s %.3f ms, %.0f docs/s" % (label, ms, n_docs / ms * 1e3), {k: round(v["ms"], 3) for k, v in prof.items() if v["ms"] > 0.02}, flush=True)

    run("first call (allocations)")
    for _ in range(2):
        tok.cache_reset(); run("cold                    ")
        run("warm                    ")
    from oracle.oracle import Oracle
    be = tok.encode_batch((blob[:off[64]], off[:65]), max_len=W)
    r = Oracle(vp, mp).encode_batch((blob[:off[64]], off[:65]), None, max_len=W, threads=Oracle.max_threads())
    print("parity on 64 documents:", bool(np.array_equal(r["ids"], be["input_ids"].reshape(-1))), flush=True)
