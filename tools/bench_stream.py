#!/usr/bin/env python3
"""BASELINE.json configs[2] / configs[3] at one GPU's share of the 100 M-pair batch: stream the pairs through the device
API in chunks with a reused output ring (165 GB of planes do not fit beside the input), then decode every chunk.

    python tools/bench_stream.py [--pairs 12500000] [--chunk 1048576] [--max-len 256]
    torchrun --nproc-per-node N tools/bench_stream.py ...      (rank r encodes chunks r, r+N, ... of the global batch)

Per chunk: the two sides are generated on the host (seeded per global chunk, so any shard can be regenerated), copied to the
GPU (untimed), encoded into the same [chunk, max_len] planes (CUDA events), and the planes are decoded (CUDA events).  Prints
one JSON line per rank 0: pairs, real tokens, encode / decode time, tokens/s, algorithmic GB/s against the measured HBM peak,
and a checksum of the row lengths that is independent of the number of GPUs (sum over chunks)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=12_500_000, help="pairs per GPU (100 M / 8)")
    ap.add_argument("--chunk", type=int, default=1 << 20)
    ap.add_argument("--max-len", type=int, default=256)
    ap.add_argument("--no-decode", action="store_true")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import torch
    import torch.distributed as dist
    from genz_tokenize_b200 import Tokenize, workload
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W, n = args.max_len, args.chunk
    tok = Tokenize(devices=[local])
    tok.set_option("max_chunk_bytes", 1 << 28)
    out = {"input_ids": torch.empty((n, W), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n, W), dtype=torch.uint8, device=dev),
           "token_type_ids": torch.empty((n, W), dtype=torch.int8, device=dev), "row_len": torch.empty((n,), dtype=torch.int32, device=dev),
           "seq_len": torch.empty((n,), dtype=torch.int32, device=dev), "row_status": torch.empty((n,), dtype=torch.uint8, device=dev)}
    pad16 = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 16, dtype=np.uint8)])).to(dev)
    n_chunks = (args.pairs + n - 1) // n
    enc_ms = dec_ms = 0.0
    pairs = tokens = in_bytes = dec_bytes = 0
    checksum = 0
    t_host = time.perf_counter()
    txt_ring = None
    for k in range(n_chunks):
        g = k * world + rank                                             # global chunk index: shard by document, no collective
        m = min(n, args.pairs - k * n)
        tb, to = workload.generate(1234, m, 3, 13, 0.0, first_chunk=2 * g)
        pb, po = workload.generate(1234, m, 3, 13, 0.0, first_chunk=2 * g + 1)
        d_t, d_to, d_p, d_po = pad16(tb), torch.from_numpy(to).to(dev), pad16(pb), torch.from_numpy(po).to(dev)
        o = out if m == n else {kk: v[:m] for kk, v in out.items()}
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        tok.encode_device(d_t, d_to, d_p, d_po, max_len=W, out=o, text_bytes=len(tb), pair_bytes=len(pb))
        b.record()
        if not args.no_decode:
            txt, toff = tok.decode_device(o["input_ids"], out=txt_ring)
            if txt_ring is None or txt_ring.numel() < txt.numel():       # the text ring: sized by the first chunk, with headroom
                txt_ring = torch.empty((int(txt.numel() * 1.02) + (1 << 20),), dtype=torch.uint8, device=dev)
        c.record()
        torch.cuda.synchronize()
        if k > 0 or n_chunks == 1:                                       # the first chunk warms the word cache (cold BPE)
            enc_ms += a.elapsed_time(b)
            dec_ms += b.elapsed_time(c)
            pairs += m
            rl = o["row_len"]
            tokens += int(rl.sum().item())
            in_bytes += len(tb) + len(pb)
            if not args.no_decode:
                dec_bytes += int(txt.numel())
        checksum = (checksum + int(o["row_len"].to(torch.int64).sum().item()) * (g + 1)) % (1 << 40)
        assert int(o["row_status"].sum().item()) == 0
    host_s = time.perf_counter() - t_host
    t = torch.tensor([enc_ms, dec_ms], dtype=torch.float64, device=dev)
    s = torch.tensor([pairs, tokens, in_bytes, dec_bytes, checksum], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    if rank == 0:
        enc_ms, dec_ms = t.tolist()
        pairs, tokens, in_bytes, dec_bytes, checksum = s.tolist()
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
        enc_alg = in_bytes + 16 * pairs + pairs * W * 6
        dec_alg = 4 * pairs * W + dec_bytes + 16 * pairs
        line = {"workload": "bundled vocab, synthetic sentence pairs (3-13 words a side), max_len=%d, planes ids+mask+token types, streamed in chunks of %d "
                            "through a reused output ring (BASELINE configs[2]/[3] at this GPU count's share)" % (W, n),
                "n_gpus": world, "pairs_timed": pairs, "chunks_per_gpu": n_chunks, "real_tokens": tokens, "input_bytes": in_bytes,
                "encode_ms": enc_ms, "encode_pairs_per_s": pairs / (enc_ms * 1e-3), "encode_tokens_per_s": tokens / (enc_ms * 1e-3),
                "encode_input_gb_per_s": in_bytes / (enc_ms * 1e-3) / 1e9, "encode_alg_gb_per_s": enc_alg / (enc_ms * 1e-3) / 1e9,
                "encode_hbm_frac": enc_alg / (enc_ms * 1e-3) / 1e9 / (peak * world),
                "decode_ms": dec_ms, "decode_rows_per_s": pairs / (dec_ms * 1e-3) if dec_ms else None, "decode_text_bytes": dec_bytes,
                "decode_alg_gb_per_s": dec_alg / (dec_ms * 1e-3) / 1e9 if dec_ms else None,
                "decode_hbm_frac": dec_alg / (dec_ms * 1e-3) / 1e9 / (peak * world) if dec_ms else None,
                "row_len_checksum": checksum, "host_seconds_total": host_s,
                "note": "times are CUDA events around the device API calls, max over ranks; the first chunk of every rank (cold word cache) is run but not timed; "
                        "decode writes into a reused text ring (sized by the first chunk) and includes one device->host read of the text size per chunk"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
