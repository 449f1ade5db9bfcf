#!/usr/bin/env python3
"""bench.py -- the encode hot path on N B200s (one process per GPU), next to the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): bundled vocab, 1,048,576 synthetic single sentences per GPU
(3-13 words sampled from vocab.txt in proportion to their counts), max_len=128,
padding + truncation.  One step = one pass of the whole batch through the CUDA pipeline.  With N GPUs
every rank encodes its own shard of documents (no collective; shards concatenate), so scaling is weak.

Prints ONE JSON line (rank 0).  `value` = real tokens/s with the text already resident in HBM and the
[n,128] planes written to HBM; `e2e` = the same through Tokenize.encode_batch with host buffers (pinned
host text in, pinned host planes out, copies inside the timed region); `roofline` = algorithmic bytes of
the dominant kernel (k_flat_rows: offsets -> planes) over its CUDA-event duration against the measured HBM copy
bandwidth, with `roofline.whole_path` = the same for all kernels of a step;
`cpu_baseline` = oracle/ (the C restatement of tokenize.py) on this host's cores, a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DOCS = 1 << 20
MAX_LEN = 128
SEED = 1234
L2_FLUSH_BYTES = 256 << 20


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.busy = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        alt = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
               "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self.stop_flag:
            try:
                if self.busy:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k in names:
                        bit = getattr(nv, names[k], None) or getattr(nv, alt[k], 0)
                        if r & bit:
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.nv and self.t.is_alive():
            self.t.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def oracle_rate(tb, to, n_sample, threads, repeats=1):
    """tokens/s of the CPU restatement on the first n_sample documents."""
    from oracle.oracle import Oracle
    o = Oracle()
    sub = (tb[:to[n_sample]], to[:n_sample + 1])
    best, toks = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = o.encode_batch(sub, None, max_len=MAX_LEN, threads=threads)
        dt = time.perf_counter() - t0
        toks = int(r["mask"].sum())
        best = dt if best is None or dt < best else best
    return toks / best, toks, best


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host CPU (oracle/ port, all host threads)."""
    if rank != 0:
        return
    from genz_tokenize_b200 import workload
    from oracle.oracle import Oracle
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm uses every core of the host regardless
    threads = max(Oracle.max_threads(), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    n_sample = N_DOCS
    tb, to = workload.generate(SEED, n_sample, 3, 13, 0.0)
    for _ in range(max(args.warmup, 0)):
        oracle_rate(tb, to, 1 << 13, threads)
    t_tot, tok_tot = 0.0, 0
    for _ in range(args.steps):
        rate, toks, dt = oracle_rate(tb, to, n_sample, threads)
        t_tot += dt
        tok_tot += toks
    value = tok_tot / t_tot
    in_bytes = int(to[n_sample])
    line = {
        "impl": "reference", "metric": "encode_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": "bundled vocab, synthetic single sentences (3-13 words), max_len=128, padding+truncation (BASELINE configs[1])",
                   "docs_per_step": n_sample, "note": "the same 1,048,576-document batch per step; host CPU only"},
        "input_gb_per_s": in_bytes * args.steps / t_tot / 1e9,
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port",
                         "sample": "all %d documents per step, oracle/genztok_oracle.c with OpenMP over documents" % n_sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=N_DOCS, help="documents per GPU per step (default: the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the pair-encode and decode side measurements")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from genz_tokenize_b200 import Tokenize, workload
    from genz_tokenize_b200 import _lib as L

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.docs
    # shard by document: rank r owns chunk r of the global batch (generator chunks are 1M documents)
    tb, to = workload.generate(SEED, n, 3, 13, 0.0, first_chunk=rank)
    in_bytes = int(to[-1])
    tok = Tokenize(devices=[local_rank])
    for kv in os.environ.get("GENZTOK_OPTIONS", "").split(","):      # e.g. GENZTOK_OPTIONS=grid_mult=4,group=7 (experiments)
        if "=" in kv:
            tok.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    # ---- device-resident leg ---------------------------------------------------------------------
    pad = (-in_bytes) % 16 + 16
    d_text = torch.from_numpy(np.concatenate([tb, np.zeros(pad, dtype=np.uint8)])).to(dev)
    d_off = torch.from_numpy(to).to(dev)
    out = {"input_ids": torch.empty((n, MAX_LEN), dtype=torch.int32, device=dev),
           "attention_mask": torch.empty((n, MAX_LEN), dtype=torch.uint8, device=dev),
           "row_len": torch.empty((n,), dtype=torch.int32, device=dev)}
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def step():
        tok.encode_device(d_text, d_off, max_len=MAX_LEN, out=out, text_bytes=in_bytes)

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.busy = True        # clocks are sampled while the GPU works: warm-up, timed steps, profiling and e2e legs
    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    tokens_per_step = int(out["attention_mask"].sum().item())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = tok.launch_count()
    for a, b in ev:
        flush.zero_()          # evict the batch and the tables from L2 (untimed)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = tok.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- dominant-kernel duration (library-side CUDA events on the launching stream), outside the timed region
    tok.set_profiling(True)
    tok.profile_report(reset=True)
    for _ in range(5):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    prof = tok.profile_report(reset=True)
    tok.set_profiling(False)
    # the dominant kernel is whichever took the most time: k_flat_rows (byte-parallel pipeline) or k_rows_fixed* (fused kernel)
    dom_name = max(prof, key=lambda kk: prof[kk]["ms"]) if prof else "k_flat_rows"
    k = prof.get(dom_name, {"launches": 1, "ms": float("nan")})
    k_ms = k["ms"] / max(k["launches"], 1)
    step_kernel_ms = sum(v["ms"] for v in prof.values()) / 5.0

    # ---- end-to-end leg: public API, host buffers, copies inside the timed region ---------------------
    lib = L.load()
    hp = lib.genztok_host_alloc(in_bytes + 64)
    import ctypes as C
    pin_text = np.frombuffer((C.c_uint8 * in_bytes).from_address(hp), dtype=np.uint8)
    pin_text[:] = tb
    e2e_ms, h2d, d2h, e2e_tokens = 0.0, 0, 0, 0
    for i in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        be = tok.encode_batch((pin_text, to), max_len=MAX_LEN)
        checksum = int(be["row_len"][-1])      # the result is in host memory when the call returns
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_ms += dt * 1e3
        h2d = in_bytes + to.nbytes
        d2h = be["input_ids"].nbytes + be["attention_mask"].nbytes + be["row_len"].nbytes
        e2e_tokens = int(be["real_tokens"])
        del be
    lib.genztok_host_free(hp)

    # ---- side measurements (N=1 only; reported under "extra", not the headline) ---------------------------
    extra = None
    if world == 1 and not args.no_extras:
        extra = {}
        peak0, _ = measured_hbm_peak()

        def timed(fn, reps=5):
            for _ in range(2):
                flush.zero_(); fn()
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for a, b in evs:
                flush.zero_(); a.record(); fn(); b.record()
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs) / reps

        # decode of the [n,128] rows just produced (BASELINE configs[3] at one GPU's share)
        holder = {}
        ms = timed(lambda: holder.__setitem__("d", tok.decode_device(out["input_ids"])))
        dbytes = int(holder["d"][0].numel())
        dalg = 4 * n * MAX_LEN + dbytes + 8 * (n + 1)
        extra["decode_padded_rows"] = {"rows_per_s": n / (ms * 1e-3), "ids_per_s": n * MAX_LEN / (ms * 1e-3), "ms": ms, "text_bytes": dbytes,
                                       "alg_gb_per_s": dalg / (ms * 1e-3) / 1e9, "hbm_frac": dalg / (ms * 1e-3) / 1e9 / peak0,
                                       "note": "includes the allocation of the text tensor and one device->host read of the total size"}
        del holder
        # sentence pairs, max_len=256, token types (BASELINE configs[2] at one GPU's share of 1,048,576 pairs)
        pb_, po_ = workload.generate(SEED + 5000, n, 3, 13, 0.0)
        pbytes = int(po_[-1])
        d_pair = torch.from_numpy(np.concatenate([pb_, np.zeros((-pbytes) % 16 + 16, dtype=np.uint8)])).to(dev)
        d_poff = torch.from_numpy(po_).to(dev)
        W2 = 256
        pout = {"input_ids": torch.empty((n, W2), dtype=torch.int32, device=dev), "attention_mask": torch.empty((n, W2), dtype=torch.uint8, device=dev),
                "token_type_ids": torch.empty((n, W2), dtype=torch.int8, device=dev), "row_len": torch.empty((n,), dtype=torch.int32, device=dev),
                "seq_len": torch.empty((n,), dtype=torch.int32, device=dev), "row_status": torch.empty((n,), dtype=torch.uint8, device=dev)}
        tok2 = Tokenize(devices=[local_rank])
        tok2.set_option("max_chunk_bytes", 1 << 27)       # both sides of 1M pairs (~106 MB) in one chunk
        ms = timed(lambda: tok2.encode_device(d_text, d_off, d_pair, d_poff, max_len=W2, out=pout, text_bytes=in_bytes, pair_bytes=pbytes))
        ptok = int(pout["row_len"].sum().item())
        palg = in_bytes + pbytes + 16 * (n + 1) + n * W2 * 6
        extra["encode_pairs_256"] = {"tokens_per_s": ptok / (ms * 1e-3), "pairs_per_s": n / (ms * 1e-3), "ms": ms, "input_gb_per_s": (in_bytes + pbytes) / (ms * 1e-3) / 1e9,
                                     "alg_gb_per_s": palg / (ms * 1e-3) / 1e9, "hbm_frac": palg / (ms * 1e-3) / 1e9 / peak0,
                                     "planes": "input_ids int32 + attention_mask uint8 + token_type_ids int8 [n,256]"}
        # decode of the [n,256] pair rows (BASELINE configs[3] at one GPU's share of 1,048,576 rows)
        holder = {}
        ms = timed(lambda: holder.__setitem__("d", tok2.decode_device(pout["input_ids"])))
        dbytes = int(holder["d"][0].numel())
        dalg = 4 * n * W2 + dbytes + 8 * (n + 1)
        extra["decode_padded_rows_256"] = {"rows_per_s": n / (ms * 1e-3), "ids_per_s": n * W2 / (ms * 1e-3), "ms": ms, "text_bytes": dbytes,
                                           "alg_gb_per_s": dalg / (ms * 1e-3) / 1e9, "hbm_frac": dalg / (ms * 1e-3) / 1e9 / peak0,
                                           "note": "includes the allocation of the text tensor and one device->host read of the total size"}
        del holder, pout, d_pair, d_poff, tok2

    clocks = sampler.stop()
    # ---- max over ranks --------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_ms, k_ms], dtype=torch.float64, device=dev)
    s = torch.tensor([tokens_per_step, in_bytes, launches, e2e_tokens], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, k_ms = [float(x) for x in t.tolist()]
    tot_tokens, tot_in_bytes, tot_launches, tot_e2e_tokens = [float(x) for x in s.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = dev_ms / args.steps
    value = tot_tokens / (ms_per_step * 1e-3)
    peak, peak_src = measured_hbm_peak()
    alg_bytes = in_bytes + 8 * (n + 1) + n * MAX_LEN * (4 + 1)          # the whole path, per GPU per step
    # algorithmic bytes of the dominant kernel alone: k_flat_rows turns offsets into planes (it never reads the text);
    # k_flat_words reads the text; the fused k_rows_fixed* kernels do the whole path in one launch
    # (the pipeline stages KR columns per row -- genztok.cu setup_tma: bytes/rows/3 + 12 rounded up to 16, at least 32 -- and
    #  k_flat_words writes the other max_len - KR pad columns of both planes on the side)
    kr = min(MAX_LEN, max(32, (in_bytes // n // 3 + 12 + 15) // 16 * 16))
    kernel_alg = {"k_flat_rows": 8 * (n + 1) + n * kr * (4 + 1), "k_flat_words": in_bytes + n * (MAX_LEN - kr) * (4 + 1)}
    dom_alg = kernel_alg.get(dom_name, alg_bytes)
    achieved = dom_alg / (k_ms * 1e-3) / 1e9
    traffic, traffic_all = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic_all = json.load(f)
            traffic = traffic_all.get(dom_name + "_dram_bytes_per_launch")
    except Exception:
        pass
    per_kernel = {}
    for name, v in prof.items():
        ms1 = v["ms"] / max(v["launches"], 1)
        ent = {"ms": ms1, "launches_per_step": v["launches"] / 5.0}
        if name in kernel_alg:
            ent["alg_bytes"] = kernel_alg[name]
            ent["alg_gb_per_s"] = kernel_alg[name] / (ms1 * 1e-3) / 1e9
            ent["frac_of_peak"] = ent["alg_gb_per_s"] / peak
        per_kernel[name] = ent
    line = {
        "metric": "encode_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": "bundled vocab, %d synthetic single sentences per GPU (3-13 words), max_len=%d, padding+truncation (BASELINE configs[1])" % (n, MAX_LEN),
                   "docs_per_gpu": n, "max_len": MAX_LEN, "sharding": "by document, no collective", "l2": "flushed between timed steps (256 MiB write)",
                   "outputs": "input_ids int32 + attention_mask uint8 [n,128] + row_len"},
        "input_gb_per_s": tot_in_bytes / (ms_per_step * 1e-3) / 1e9,
        "alg_gb_per_s": world * alg_bytes / (ms_per_step * 1e-3) / 1e9,
        "hbm_frac_of_step": world * alg_bytes / (ms_per_step * 1e-3) / 1e9 / (world * peak),
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": dom_alg, "kernel_ms": k_ms,
                     "kernels_ms_per_step": step_kernel_ms, "kernel_share_of_step": k_ms / step_kernel_ms if step_kernel_ms else None,
                     "whole_path": {"alg_bytes_per_step": alg_bytes, "achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                                    "frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                                    "note": "all kernels of a step against the same peak: the number to compare with the 50% target"}},
        "e2e": {"value": tot_e2e_tokens / (e2e_ms / max(args.e2e_steps, 1) * 1e-3) if e2e_ms else None, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / max(args.e2e_steps, 1), "api": "Tokenize.encode_batch(packed text in pinned host memory) -> pinned numpy planes"},
        "gpu_launches": int(tot_launches),
        "clocks": clocks,
        "kernels": per_kernel,
    }
    if extra is not None:
        line["extra"] = extra
    if not args.no_cpu and world == 1:            # the CPU baseline is reported at N=1 only
        threads = 0
        try:
            from oracle.oracle import Oracle
            threads = max(Oracle.max_threads(), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
            n1 = min(n, 1 << 19)
            r1, _, dt1 = oracle_rate(tb, to, n1, 1)
            reps = 12
            rN, _, dtN = oracle_rate(tb, to, n, threads, repeats=reps)
            line["cpu_baseline"] = {"value": rN, "unit": "tokens/s", "cores": threads, "kind": "port", "single_thread_value": r1,
                                    "sample": "oracle/genztok_oracle.c (C restatement of tokenize.py, no memoisation): all %d documents with %d OpenMP threads, "
                                              "best of %d passes (%.2f s each); single thread on the first %d documents (%.1f s)" % (n, threads, reps, dtN, n1, dt1)}
        except Exception as e:   # the baseline is reporting only; never fail the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": "failed: %r" % (e,)}
        # The pure-Python reference itself cannot run here (its sources are not in this repository); what it did in the build
        # container is on file (tools/time_python_reference.py) and quoted beside the port's numbers, marked as such.
        try:
            with open(os.path.join(ROOT, "profiles", "python_reference_container.json")) as f:
                pr = json.loads(f.readline())
            line["cpu_baseline"]["python_reference_recorded"] = {
                "where": "build container (%d vCPU), not this host: recorded by tools/time_python_reference.py" % pr["cores"],
                "single_process_tokens_per_s": pr["singles_128_single_process"]["tokens_per_s"],
                "pool_tokens_per_s": pr["singles_128_pool"]["tokens_per_s"], "pool_workers": pr["singles_128_pool"]["workers"]}
        except Exception:
            pass
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
