"""Multi-GPU path on CPU: documents shard by contiguous ranges across ranks with no data-path collective
(SURVEY.md §8e).  Two gloo processes each take their shard, encode it (the CPU oracle stands in for the GPU
engine here -- this test is about the host-side sharding and concatenation logic) and the concatenation of the
shards must equal the single-process result; the generator's chunks must be reproducible per rank."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from genz_tokenize_b200 import workload
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tb, to = workload.generate(99, n, 0, 9, 0.1)              # every rank can rebuild the global batch deterministically
    lo, hi = workload.shard_range(n, rank, world)
    sub = (tb[to[lo]:to[hi]], to[lo:hi + 1] - to[lo])
    r = Oracle().encode_batch(sub, None, max_len=24)
    digest = hashlib.sha256(r["ids"].tobytes()).digest()
    # the only cross-rank traffic: bookkeeping (row counts), never token data
    rows = torch.tensor([hi - lo], dtype=torch.int64)
    allrows = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allrows, rows)
    dist.barrier()
    q.put((rank, lo, hi, r["ids"], digest, [int(t) for t in allrows]))
    dist.destroy_process_group()


def test_two_rank_shards_concatenate_to_the_full_batch():
    import torch.multiprocessing as mp
    from genz_tokenize_b200 import workload
    from oracle.oracle import Oracle
    n, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    tb, to = workload.generate(99, n, 0, 9, 0.1)
    full = Oracle().encode_batch((tb, to), None, max_len=24)
    assert [o[1] for o in outs] == [0, 500] and [o[2] for o in outs] == [500, 1001]
    assert outs[0][5] == [500, 501]
    assert np.array_equal(np.concatenate([o[3] for o in outs]), full["ids"])


def test_shard_ranges_partition():
    from genz_tokenize_b200.workload import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1


def test_generator_chunks_are_rank_reproducible(monkeypatch):
    from genz_tokenize_b200 import workload
    monkeypatch.setattr(workload, "CHUNK", 300)
    tb, to = workload.generate(5, 900, 3, 6, 0.05)
    for c in range(3):
        cb, co = workload.generate(5, 300, 3, 6, 0.05, first_chunk=c)
        assert np.array_equal(cb, tb[to[300 * c]:to[300 * (c + 1)]])
        assert np.array_equal(co, to[300 * c:300 * (c + 1) + 1] - to[300 * c])
