"""MultiEngine -- the multi-GPU host object of the path: one Engine and one feeder thread per GPU, streams
placed in contiguous blocks (sharding.py), host buffers in, packets gathered on the host as (global stream, seq, payload).

Streams are independent (reference src/fsk.h:43-90: no state is shared between two `struct FSK`; src/drs232_ldpc.c:106-118:
nor between two deframers), so there is NO collective on this path and no torch.distributed here: the engines only
share the caller's host arrays.  Every libwenet_b200 call releases the GIL (ctypes), so the per-GPU threads really run
side by side: one GPU's host->device copy overlaps another's kernels and drains.

    me = MultiEngine(32768, devices=range(8), in_fmt="cu8", framing="v1", chunk_samples=1 << 19)
    me.feed_strided(block)          # block: [32768, 2 * nsamp] uint8, ideally from me.pinned_block(nsamp)
    me.process(); me.sync()
    pk = me.drain_all_packets()     # fields stream (global), seq, payload

`weights` sizes the blocks in proportion to what each GPU can take (e.g. its host->device copy rate when all GPUs copy
at once, tools/micro/pcie_multi.cu); the default is equal blocks.  `engines_per_device` > 1 splits a GPU's block over
several engines whose copies and kernels overlap each other on that GPU (what bench.py's e2e leg does with two).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import engine as E
from . import sharding


def probe_copy_rates(devices, seconds=0.5, nbytes=256 << 20):
    """Pinned-host -> device copy rate (GB/s) of every device in `devices` with ALL of them copying at once: what each
    GPU can take when the whole box is busy feeding.  On an HGX box the GPUs do not get equal shares (shared host bridges,
    profiles/r02_pcie_multi_8gpu.jsonl); MultiEngine(weights=probe_copy_rates(devs)) sizes the stream blocks accordingly."""
    import threading
    import time
    devices = list(devices)
    probes = [E.CopyProbe(d, nbytes) for d in devices]
    rates = [0.0] * len(devices)
    gate = threading.Barrier(len(devices))

    def work(i):
        p = probes[i]
        p.run(1)
        gate.wait()
        t0, nb = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds:
            p.run(1)
            nb += p.nbytes
        rates[i] = nb / (time.perf_counter() - t0) / 1e9

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    [t.start() for t in th]
    [t.join() for t in th]
    for p in probes:
        p.close()
    return rates


class MultiEngine:
    def __init__(self, n_streams, devices=(0,), weights=None, engines_per_device=1, engine_cls=None, **engine_kw):
        devices = list(devices)
        if not devices or n_streams < len(devices) * engines_per_device:
            raise ValueError("need at least one stream per engine")
        cls = engine_cls or E.Engine
        slots = [d for d in devices for _ in range(engines_per_device)]
        w = None if weights is None else [float(weights[i // engines_per_device]) / engines_per_device for i in range(len(slots))]
        self.ranges = sharding.equal_ranges(n_streams, len(slots)) if w is None else sharding.weighted_ranges(n_streams, w)
        if any(hi <= lo for lo, hi in self.ranges):
            raise ValueError("a weight leaves an engine without streams")
        self.n_streams = n_streams
        self.devices = slots
        self.fmt = engine_kw.get("in_fmt", "cf32")
        self.pool = ThreadPoolExecutor(max_workers=len(slots), thread_name_prefix="wb-gpu")
        self.engines = []
        try:
            # engines are created one after the other (cudaMalloc of the resident buffers), then driven in parallel
            for (lo, hi), d in zip(self.ranges, slots):
                self.engines.append(cls(hi - lo, device=d, **engine_kw))
        except Exception:
            self.close()
            raise
        self._los = np.array([lo for lo, _ in self.ranges], dtype=np.int64)

    # ---- placement ----
    def owner(self, stream):
        """-> (engine index, local stream) of a global stream"""
        if not (0 <= stream < self.n_streams):
            raise IndexError("stream %d out of range" % stream)
        k = int(np.searchsorted(self._los, stream, side="right")) - 1
        return k, stream - int(self._los[k])

    def _each(self, fn):
        futs = [self.pool.submit(fn, k, g) for k, g in enumerate(self.engines)]
        return [f.result() for f in futs]           # re-raises the first engine error

    def each(self, fn):
        """run fn(engine index, engine) on every engine, each in its own thread -> list of results"""
        return self._each(fn)

    def pinned_block(self, nsamp):
        """one pinned host array [n_streams, elems * nsamp] for feed_strided (each engine copies its own rows of it)"""
        return E.PinnedBuffer((self.n_streams, E.FMT_ELEMS[self.fmt] * nsamp), E.FMT_DTYPE[self.fmt])

    # ---- streaming path: the calls of Engine, over global streams ----
    def feed_strided(self, block):
        assert block.shape[0] == self.n_streams
        self._each(lambda k, g: g.feed_strided(block[self.ranges[k][0]:self.ranges[k][1]]))

    def feed(self, streams):
        assert len(streams) == self.n_streams
        self._each(lambda k, g: g.feed(streams[self.ranges[k][0]:self.ranges[k][1]]))

    def process(self):
        self._each(lambda k, g: g.process())

    def process_soft(self, streams):
        assert len(streams) == self.n_streams
        self._each(lambda k, g: g.process_soft(streams[self.ranges[k][0]:self.ranges[k][1]]))

    def sync(self):
        self._each(lambda k, g: g.sync())

    def step(self, block):
        """feed_strided + process on every engine, each in its own thread (one engine's copy overlaps another's kernels)"""
        def one(k, g):
            g.feed_strided(block[self.ranges[k][0]:self.ranges[k][1]])
            g.process()
        self._each(one)

    def stream_step(self, block):
        """The steady-state form of a receiver: on every engine (its own thread) collect the packets of the step queued
        before (sync + drain), then feed this block and start processing it -- so one engine's host work and copy overlap
        the kernels of the others, and with engines_per_device > 1 also those of its neighbour on the same GPU.
        -> packets of the PREVIOUS step (global stream numbers), empty on the first call; flush() returns the last ones."""
        def one(k):
            g = self.engines[k]
            out = None
            if self._pending[k]:
                g.sync()
                out = g.drain_all_packets()
                out["stream"] += self.ranges[k][0]
            g.feed_strided(block[self.ranges[k][0]:self.ranges[k][1]])
            g.process()
            self._pending[k] = True
            return out

        def device(ks):
            # the engines of ONE GPU take their turns in one thread: while engine i waits for its kernels and drains,
            # the copy engine i + 1 queued a moment ago is running, and the other way round
            return [one(k) for k in ks]
        if not hasattr(self, "_pending"):
            self._pending = [False] * len(self.engines)
        groups = {}
        for k, d in enumerate(self.devices):
            groups.setdefault(d, []).append(k)
        futs = [self.pool.submit(device, ks) for ks in groups.values()]
        parts = [p for f in futs for p in f.result() if p is not None]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=self._pk_dtype())

    def flush(self):
        """wait for what stream_step queued last and return its packets"""
        def one(k, g):
            if not getattr(self, "_pending", [False] * len(self.engines))[k]:
                return None
            g.sync()
            out = g.drain_all_packets()
            out["stream"] += self.ranges[k][0]
            self._pending[k] = False
            return out
        parts = [p for p in self._each(one) if p is not None]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=self._pk_dtype())

    @staticmethod
    def _pk_dtype():
        return np.dtype([("stream", "<i4"), ("seq", "<u4"), ("payload", "u1", (E.PACKET_BYTES,))])

    def drain_all_packets(self):
        """-> (stream, seq, payload[256]) of every engine, global stream numbers, sorted by (stream, seq)"""
        parts = self._each(lambda k, g: g.drain_all_packets())
        for (lo, _), p in zip(self.ranges, parts):
            p["stream"] += lo
        return np.concatenate(parts) if parts else parts

    def drain_packets(self, stream):
        k, s = self.owner(stream)
        return self.engines[k].drain_packets(s)

    def drain_soft(self, stream):
        k, s = self.owner(stream)
        return self.engines[k].drain_soft(s)

    def drain_codewords(self):
        parts = self._each(lambda k, g: g.drain_codewords())
        for (lo, _), p in zip(self.ranges, parts):
            p["stream"] += lo
        return np.concatenate(parts)

    def stats(self, stream):
        k, s = self.owner(stream)
        return self.engines[k].stats(s)

    def nin(self):
        return np.concatenate(self._each(lambda k, g: g.nin()))

    @property
    def last_samples(self):
        return sum(g.last_samples for g in self.engines)

    @property
    def last_codewords(self):
        return sum(g.last_codewords for g in self.engines)

    @property
    def launch_count(self):
        return sum(g.launch_count for g in self.engines)

    def close(self):
        for g in getattr(self, "engines", []):
            try:
                g.close()
            except Exception:
                pass
        self.engines = []
        if getattr(self, "pool", None):
            self.pool.shutdown(wait=True)
            self.pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
