#!/usr/bin/env python3
"""Experiment: the headline workload (4096 streams x 1 Mi samples resident, v1 end to end) on ONE engine against
the same streams split over TWO engines of 2048 whose steps are issued back to back on their own CUDA streams, so
that one engine's deframe + LDPC kernels can run beside the other's FSK kernel.  Wall clock around a full sync
(steps are tens of ms; exploration only, bench.py stays the graded measurement)."""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bench                                   # noqa: E402
from wenet_b200 import engine as E             # noqa: E402

N, CHUNK, NSRC, STEPS, WARM = 4096, 1 << 20, 16, 10, 3
sources = bench.make_sources(NSRC, CHUNK, seed_base=0, mode="v1")


def build(n):
    e = E.Engine(n, in_fmt="cf32", chunk_samples=CHUNK, framing="v1")
    e.feed(sources + [None] * (n - NSRC))
    e.sync()
    e.dev_replicate(NSRC, CHUNK, 4096 + 16 * 37)
    e.dev_set_fill(CHUNK)
    return e


def run(engs, stagger_ms):
    def step(e):
        e.dev_set_fill(CHUNK)
        e.process()
    for k in range(WARM):
        for e in engs:
            step(e)
    for e in engs:
        e.sync()
        e.drain_all_packets()
    t0 = time.perf_counter()
    # engine 0 starts first; everything after that is queued back to back on the two CUDA streams
    step(engs[0])
    if stagger_ms:
        time.sleep(stagger_ms * 1e-3)
    for k in range(STEPS):
        for i, e in enumerate(engs):
            if not (k == 0 and i == 0):
                step(e)
    for e in engs:
        e.sync()
    dt = time.perf_counter() - t0
    samples = sum(e.last_samples for e in engs)
    pk = sum(len(e.drain_all_packets()) for e in engs)
    return dt / STEPS * 1e3, samples / (dt / STEPS) / 1e6, pk, [np.round(e.last_kernel_ms(), 2).tolist() for e in engs]


one = build(N)
print("1 engine  x %d: %.2f ms/step  %.0f Msamples/s  packets %d  kernel ms %s" % ((N,) + run([one], 0)), flush=True)
one.close()
two = [build(N // 2), build(N // 2)]
for st in (0, 10, 20, 30, 40):
    print("2 engines x %d, engine 1 %2d ms behind: %.2f ms/step  %.0f Msamples/s  packets %d  kernel ms %s"
          % ((N // 2, st) + run(two, st)), flush=True)
for e in two:
    e.close()
