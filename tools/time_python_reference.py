#!/usr/bin/env python3
"""Times the UNMODIFIED pure-Python reference (genz_tokenize.Tokenize imported from /root/reference) on this host's cores:
single process and multiprocessing.Pool(all cores), on the synthetic workloads of BASELINE.json.

The reference cannot travel to the GPU box (its sources are not copied into this repository), so this runs in the build
container only; bench.py's `cpu_baseline` / `--impl reference` legs time the C restatement (oracle/) on the GPU box instead.

    python tools/time_python_reference.py [--ref /root/reference] [--out profiles/python_reference_container.json]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_tok = None


def _init(ref):
    global _tok
    sys.path.insert(0, ref)
    from genz_tokenize import Tokenize
    _tok = Tokenize()


def _encode_chunk(job):
    texts, pairs, max_len = job
    n = 0
    for i, t in enumerate(texts):
        out = _tok(t, pairs[i] if pairs else None, max_len=max_len, padding=True, truncation=True)
        n += sum(out["attention_mask"])
    return n


def _decode_chunk(rows):
    return sum(len(_tok.decode(r)) for r in rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "python_reference_container.json"))
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(args.ref, "genz_tokenize")):
        print(json.dumps({"unavailable": "no reference checkout at %s" % args.ref}))
        return
    from genz_tokenize_b200 import workload
    cores = os.cpu_count()

    def texts_of(seed, n):
        b, o = workload.generate(seed, n, 3, 13, 0.0)
        raw = b.tobytes()
        return [raw[o[i]:o[i + 1]].decode("utf-8") for i in range(n)]

    res = {"host": "build container", "cores": cores, "python": sys.version.split()[0], "reference": args.ref}
    _init(args.ref)
    # configs[0]: 10 k pairs, max_len 256, one process
    ta, tb = texts_of(1234, 10000), texts_of(6234, 10000)
    t0 = time.perf_counter(); tokens = _encode_chunk((ta, tb, 256)); dt = time.perf_counter() - t0
    res["pairs_256_single_process"] = {"pairs": 10000, "seconds": dt, "pairs_per_s": 10000 / dt, "tokens_per_s": tokens / dt}
    # configs[1]-like: 20 k single sentences, max_len 128, one process
    ts = texts_of(1234, 20000)
    t0 = time.perf_counter(); tokens = _encode_chunk((ts, None, 128)); dt = time.perf_counter() - t0
    res["singles_128_single_process"] = {"sentences": 20000, "seconds": dt, "sentences_per_s": 20000 / dt, "tokens_per_s": tokens / dt}
    # decode of 10 k padded 256-id rows, one process
    rows = [_tok(a, b, max_len=256)["input_ids"] for a, b in zip(ta[:10000], tb[:10000])]
    t0 = time.perf_counter(); nbytes = _decode_chunk(rows); dt = time.perf_counter() - t0
    res["decode_256_single_process"] = {"rows": 10000, "seconds": dt, "rows_per_s": 10000 / dt, "chars_per_s": nbytes / dt}
    # pools over all cores: one Tokenize() per worker, 500-document chunks
    with mp.Pool(cores, initializer=_init, initargs=(args.ref,)) as pool:
        n = 40000
        pa, pb = texts_of(1234, n), texts_of(6234, n)
        jobs = [(pa[i:i + 500], pb[i:i + 500], 256) for i in range(0, n, 500)]
        pool.map(_encode_chunk, jobs[:cores])                                      # warm the workers
        t0 = time.perf_counter(); tokens = sum(pool.map(_encode_chunk, jobs)); dt = time.perf_counter() - t0
        res["pairs_256_pool"] = {"pairs": n, "workers": cores, "seconds": dt, "pairs_per_s": n / dt, "tokens_per_s": tokens / dt}
        n = 80000
        ps = texts_of(1234, n)
        jobs = [(ps[i:i + 500], None, 128) for i in range(0, n, 500)]
        t0 = time.perf_counter(); tokens = sum(pool.map(_encode_chunk, jobs)); dt = time.perf_counter() - t0
        res["singles_128_pool"] = {"sentences": n, "workers": cores, "seconds": dt, "sentences_per_s": n / dt, "tokens_per_s": tokens / dt}
        jobs = [rows[i:i + 500] for i in range(0, len(rows), 500)]
        t0 = time.perf_counter(); pool.map(_decode_chunk, jobs); dt = time.perf_counter() - t0
        res["decode_256_pool"] = {"rows": len(rows), "workers": cores, "seconds": dt, "rows_per_s": len(rows) / dt}
    line = json.dumps(res)
    print(line)
    with open(args.out, "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
