"""What the host side of the box gives N ranks at once: pinned-memory copies device->host, host->device and both, alone (rank 0) and with
every rank active.  The denominator of the end-to-end (host buffers) scaling of bench.py.
    torchrun --nproc-per-node N tools/pcie_probe.py"""
import json, os, time
import torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
MB = 512
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_in.zero_()
d_a = torch.zeros(MB << 20, dtype=torch.uint8, device="cuda")
d_b = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def leg(kind, active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if not active:
        if world > 1:
            dist.barrier()
        return None
    reps = 6
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s1):
                h_out.copy_(d_a, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d_b.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    return reps * (MB << 20) / dt / 1e9          # GB/s per direction

res = {}
for kind in ("d2h", "h2d", "both"):
    leg(kind, True)                               # warm-up
    alone = leg(kind, rank == 0)
    allr = leg(kind, True)
    if world > 1:
        t = torch.tensor([allr], device="cuda")
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        res[kind] = {"rank0_alone_gb_s": alone, "all_ranks_min_gb_s": float(lo), "all_ranks_sum_gb_s": float(sm)}
    else:
        res[kind] = {"rank0_alone_gb_s": alone}
if rank == 0:
    print(json.dumps({"n_gpus": world, "per_direction": res}))
if world > 1:
    dist.destroy_process_group()
