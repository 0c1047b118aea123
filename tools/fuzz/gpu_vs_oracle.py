"""Randomised parity sweep on a GPU box: batches of random streams (seed, Eb/N0 in [0, 14] dB, clock offset in +-5000 ppm, ragged
lengths, a few empty ones; one input format and framing per batch) through ONE engine launch each, every stream compared with the
CPU oracle -- soft decisions bit for bit, packets byte for byte.  The oracle runs in a process pool beside the GPU.

    python tools/fuzz/gpu_vs_oracle.py SECONDS [STREAMS_PER_BATCH]

Prints one line per batch and a summary; exit code 1 on any mismatch."""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from wenet_b200 import siggen                      # noqa: E402

BATCHES = [("cf32", "v1"), ("cu8", "v1"), ("cs16", "v2"), ("cs16", "v1"), ("cf32", "v2")]


def make(args):
    seed, n_samples, ebno, ppm, fmt, framing = args
    if n_samples == 0:
        return np.zeros(0, dtype=siggen.make_stream(1, n_samples=2000, fmt=fmt, framing=framing)[0].dtype)
    return siggen.make_stream(seed, n_samples=n_samples, ebno_db=ebno, fmt=fmt, framing=framing, clock_ppm=ppm)[0]


def oracle_run(args):
    raw, fmt, framing = args
    from oracle import oracle as O
    global _ORC
    try:
        _ORC
    except NameError:
        _ORC = O.Oracle("port")
    if raw.size == 0:
        return np.zeros(0, dtype=np.float32), b""
    cfg = siggen.V1 if framing == "v1" else siggen.V2
    sd, _, _ = _ORC.fsk(cfg["Fs"], cfg["Rs"], M=2, P=None).run(raw, fmt)
    res = _ORC.deframer(framing, 10).feed(sd)
    return sd, res["packets"]


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    from oracle import oracle as O
    O.build()
    from wenet_b200 import engine as E
    rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "20261017")))
    t0 = time.time()
    total = bad = packets = 0
    b = 0
    with Pool(min(32, os.cpu_count() or 4)) as pool:
        while time.time() - t0 < budget:
            fmt, framing = BATCHES[b % len(BATCHES)]
            b += 1
            cfg = siggen.V1 if framing == "v1" else siggen.V2
            specs = []
            for s in range(per):
                n = 0 if rng.random() < 0.03 else int(rng.integers(9000, 70000))
                specs.append((int(rng.integers(1, 1 << 30)), n, float(rng.uniform(0.0, 14.0)), float(rng.uniform(-5000.0, 5000.0)), fmt, framing))
            raws = pool.map(make, specs, chunksize=8)
            fut = pool.map_async(oracle_run, [(r, fmt, framing) for r in raws], chunksize=4)
            e = E.Engine(per, Fs=cfg["Fs"], Rs=cfg["Rs"], in_fmt=fmt, framing=framing, chunk_samples=72000)
            e.feed(raws)
            e.process()
            e.sync()
            got = [(e.drain_soft(s), e.drain_packets(s)) for s in range(per)]
            e.close()
            want = fut.get()
            nb = 0
            for s in range(per):
                sd_g, pk_g = got[s]
                sd_o, pk_o = want[s]
                ok = sd_g.size == sd_o.size and np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32)) and pk_g == pk_o
                packets += len(pk_o) // 256
                if not ok:
                    nb += 1
                    print("MISMATCH batch %d stream %d spec %r" % (b, s, specs[s]), flush=True)
            total += per
            bad += nb
            print("batch %d %s %s: %d streams, %d mismatches (%.0f s)" % (b, fmt, framing, per, nb, time.time() - t0), flush=True)
    print("streams %d, packets %d, mismatches %d" % (total, packets, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
