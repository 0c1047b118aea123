"""Multi-GPU plumbing on CPU: streams shard across ranks with no data-path collective; only packet counts and
timings are reduced (gloo, world_size 2).  The per-rank 'engine' here is the CPU oracle standing in as the checker
of the host-side partitioning logic."""
import os
import socket
import sys

import numpy as np
import pytest

from wenet_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_everything():
    for n in (1, 7, 8, 4096, 4097):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                lo, hi = sharding.shard_range(n, r, w)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
                for s in (lo, hi - 1):
                    if lo < hi:
                        assert sharding.owner_of(s, n, w) == r
            assert seen == list(range(n))
            sizes = [sharding.shard_range(n, r, w)[1] - sharding.shard_range(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.weak_global_streams(4096, 8) == 32768
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _worker(rank, world, port, n_streams, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from wenet_b200 import siggen
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    port_o = O.Oracle("port")
    lo, hi = sharding.shard_range(n_streams, rank, world)
    npk, nbytes = 0, 0
    digest = np.zeros(n_streams, dtype=np.int64)
    for s in range(lo, hi):
        raw, _ = siggen.make_stream(s, n_packets=1, ebno_db=10.0, fmt="cu8")
        sd, _, _ = port_o.fsk(921416, 115177).run(raw, "cu8")
        pk = port_o.deframer("v1", 10).feed(sd)["packets"]
        npk += len(pk) // 256
        digest[s] = int(np.frombuffer(pk, np.uint8).astype(np.int64).sum())
    t = torch.tensor([npk], dtype=torch.int64)
    dist.all_reduce(t)                                   # only bookkeeping is reduced
    d = torch.from_numpy(digest)
    dist.all_reduce(d)
    tm = torch.tensor([float(rank + 1)])
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)            # the bench's max-over-ranks timing
    if rank == 0:
        q.put((int(t.item()), d.numpy().tolist(), float(tm.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(oracle_port):
    import torch.multiprocessing as mp
    from wenet_b200 import siggen
    n_streams = 5
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    [p.start() for p in procs]
    total, digest, tmax = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = []
    for st in range(n_streams):
        raw, _ = siggen.make_stream(st, n_packets=1, ebno_db=10.0, fmt="cu8")
        sd, _, _ = oracle_port.fsk(921416, 115177).run(raw, "cu8")
        pk = oracle_port.deframer("v1", 10).feed(sd)["packets"]
        want.append(int(np.frombuffer(pk, np.uint8).astype(np.int64).sum()))
    assert total == n_streams and digest == want and tmax == 2.0
