"""BASELINE configs[4]: Tokenize.fromFile with a large custom vocab / merge table on long documents (max_len=4096,
heavy truncation, low word reuse).  Prints device-resident and host-API timings."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from genz_tokenize_b200 import Tokenize, workload

def custom_model(td, n_words=120000, seed=77):
    rng = np.random.default_rng(seed)
    wl = workload.default_wordlist()
    base = [wl.words[int(i)] for i in rng.integers(0, len(wl.words), size=n_words)]
    words = list(dict.fromkeys(base + [a + "_" + b for a, b in zip(base[::2], base[1::2])]))
    merges, seen, vocab = [], set(), {}
    for w in words:
        syms = list(w[:-1]) + [w[-1] + "</w>"]
        cur = syms[0]
        for s in syms[1:]:
            if (cur, s) not in seen:
                seen.add((cur, s)); merges.append("%s %s" % (cur, s))
            cur = cur + s
            vocab.setdefault(cur.replace("</w>", "") + ("" if cur.endswith("</w>") else "@@"), 1)
        vocab[w] = 1
    vp, mp = os.path.join(td, "vocab.txt"), os.path.join(td, "bpe.codes")
    open(vp, "w", encoding="utf-8").write("".join("%s %d\n" % (k, v) for k, v in vocab.items()))
    open(mp, "w", encoding="utf-8").write("#version: 0.2\n" + "\n".join(merges) + "\n")
    return vp, mp, words, len(vocab), len(merges)

with tempfile.TemporaryDirectory() as td:
    t0 = time.time()
    vp, mp, words, nv, nm = custom_model(td)
    tok = Tokenize.fromFile(vp, mp)
    tok.set_option("max_chunk_bytes", 1 << 28)
    print("custom model: %d vocab entries, %d merges, built+loaded in %.1fs" % (nv, nm, time.time() - t0), flush=True)
    rng = np.random.default_rng(5)
    wb = [w.encode() for w in words]
    n_docs = 2000
    docs = []
    for i in range(n_docs):
        k = int(rng.integers(6000, 9000))
        docs.append(b" ".join(wb[int(j)] for j in rng.integers(0, len(wb), size=k)))     # uniform sampling: low reuse
    off = np.zeros(n_docs + 1, dtype=np.int64); np.cumsum([len(d) for d in docs], out=off[1:])
    blob = np.frombuffer(b"".join(docs), dtype=np.uint8)
    print("docs: %d, %.1f MB, %.0f B/doc" % (n_docs, off[-1] / 1e6, off[-1] / n_docs), flush=True)
    dev = torch.device("cuda:0")
    d_text = torch.from_numpy(np.concatenate([blob, np.zeros(32, dtype=np.uint8)])).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    W = 4096
    out = {"input_ids": torch.empty((n_docs, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n_docs, W), dtype=torch.uint8, device=dev),
           "row_len": torch.empty((n_docs,), dtype=torch.int32, device=dev)}
    for rep in range(4):
        if rep == 2:
            tok.set_profiling(True); tok.profile_report(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tok.encode_device(d_text, d_off, max_len=W, out=out, text_bytes=int(off[-1])); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        alg = int(off[-1]) + 8 * (n_docs + 1) + n_docs * W * 5
        print("rep %d: %.3f ms, %.1f docs/ms, input %.1f GB/s, alg %.0f GB/s, real tokens %d" % (rep, ms, n_docs / ms, off[-1] / ms / 1e6, alg / ms / 1e6, int(out["row_len"].sum())), flush=True)
    print(tok.profile_report(), flush=True)
    t0 = time.time(); be = tok.encode_batch((blob, off), max_len=W); print("host API: %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
    from oracle.oracle import Oracle
    orc = Oracle(vp, mp)
    t0 = time.time(); r = orc.encode_batch((blob[:off[200]], off[:201]), None, max_len=W, threads=Oracle.max_threads()); dt = time.time() - t0
    print("oracle (%d threads) on 200 docs: %.2f s -> %.1f docs/s; parity on those: %s" % (Oracle.max_threads(), dt, 200 / dt, bool(np.array_equal(r["ids"], be["input_ids"][:200].reshape(-1)))), flush=True)
