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


def test_weighted_ranges():
    """blocks in proportion to what each GPU can take; sizes add up exactly, within one stream of the exact share"""
    for n, w in ((32768, [23.3] * 4 + [35.3] * 4), (10, [1, 1, 1]), (7, [5, 0.5, 1.5]), (4096, [1.0])):
        r = sharding.weighted_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        for (lo, hi), x in zip(r, w):
            assert abs((hi - lo) - n * x / sum(w)) < 1.0
    r = sharding.weighted_ranges(32768, [23.3] * 4 + [35.3] * 4)
    assert [hi - lo for lo, hi in r[:4]] == [3257] * 4 or sum(hi - lo for lo, hi in r[:4]) in (13028, 13029)
    assert sharding.equal_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    with pytest.raises(ValueError):
        sharding.weighted_ranges(8, [0, 0])


class _FakeEngine:
    """stands in for one GPU: remembers what it was fed, 'decodes' one packet per stream whose payload names the stream"""
    made = []

    def __init__(self, n_streams, device=0, **kw):
        self.n_streams, self.device, self.kw = n_streams, device, kw
        self.fed, self.steps = None, 0
        _FakeEngine.made.append(self)

    def feed_strided(self, block):
        assert block.shape[0] == self.n_streams
        self.fed = block

    def process(self):
        self.steps += 1

    def sync(self):
        pass

    def drain_all_packets(self):
        dt = np.dtype([("stream", "<i4"), ("seq", "<u4"), ("payload", "u1", (256,))])
        out = np.zeros(self.n_streams, dtype=dt)
        out["stream"] = np.arange(self.n_streams)
        out["seq"] = self.steps - 1
        out["payload"][:, 0] = self.fed[:, 0]              # first byte of what this stream was fed
        return out

    def drain_soft(self, s):
        return np.full(3, float(self.fed[s, 0]), dtype=np.float32)

    def nin(self):
        return np.full(self.n_streams, 384, dtype=np.uint32)

    last_samples = 10
    last_codewords = 1
    launch_count = 5

    def close(self):
        pass


def test_multi_engine_places_feeds_and_gathers():
    """the multi-GPU host object with fake engines: contiguous placement (equal and weighted), each engine is fed its own
    rows of the global block from its own thread, packets come back under global stream numbers"""
    from wenet_b200.multi import MultiEngine
    _FakeEngine.made.clear()
    n = 11
    block = np.zeros((n, 8), dtype=np.uint8)
    block[:, 0] = np.arange(n) + 100
    me = MultiEngine(n, devices=[0, 1, 2], engine_cls=_FakeEngine, in_fmt="cu8", framing="v1")
    assert [g.n_streams for g in me.engines] == [4, 4, 3] and [g.device for g in me.engines] == [0, 1, 2]
    assert me.engines[0].kw == {"in_fmt": "cu8", "framing": "v1"}
    me.step(block)
    me.sync()
    pk = me.drain_all_packets()
    assert pk["stream"].tolist() == list(range(n)) and pk["payload"][:, 0].tolist() == (np.arange(n) + 100).tolist()
    assert [me.owner(s) for s in (0, 3, 4, 8, 10)] == [(0, 0), (0, 3), (1, 0), (2, 0), (2, 2)]
    assert me.drain_soft(9)[0] == 109.0 and me.nin().size == n
    assert me.last_samples == 30 and me.launch_count == 15
    # the pipelined form: packets of step k come back with step k + 1 (or with flush)
    first = me.stream_step(block)
    assert len(first) == 0
    second = me.stream_step(block)
    assert second["stream"].tolist() == list(range(n)) and set(second["seq"].tolist()) == {1}
    last = me.flush()
    assert last["stream"].tolist() == list(range(n)) and set(last["seq"].tolist()) == {2} and len(me.flush()) == 0
    assert me.each(lambda k, g: g.n_streams) == [4, 4, 3]
    me.close()
    # weighted, two engines per device
    me = MultiEngine(12, devices=[0, 1], weights=[1.0, 2.0], engines_per_device=2, engine_cls=_FakeEngine)
    assert [g.n_streams for g in me.engines] == [2, 2, 4, 4] and [g.device for g in me.engines] == [0, 0, 1, 1]
    me.close()
    with pytest.raises(ValueError):
        MultiEngine(2, devices=[0, 1, 2], engine_cls=_FakeEngine)
