"""Randomised sweep of the oracle restatement against the compiled reference (oracle/_ref) on FSK streams: random seed, framing,
input format, Eb/N0 in [-2, 16] dB, clock offset in +-6000 ppm, P in the divisors of Ts.  Needs /root/reference-built oracle/_ref.
    python tools/fuzz/oracle_fsk.py SECONDS
Round 2: 19 463 streams in 420 s, 0 mismatches (soft decisions, nin, f_est, timing, ppm, Eb/N0, phi_c, fft_est bit for bit)."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from wenet_b200 import siggen
O.build()
port, ref = O.Oracle("port"), O.Oracle("reference")
rng = np.random.default_rng(20261017)
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < float(sys.argv[1]):
    seed = int(rng.integers(1, 1 << 30))
    cfg = siggen.V1 if rng.random() < 0.7 else siggen.V2
    framing = "v1" if cfg is siggen.V1 else "v2"
    fmt = str(rng.choice(["cf32", "cs16", "cu8"]))
    ebno = float(rng.uniform(-2.0, 16.0))
    ppm = float(rng.uniform(-6000.0, 6000.0))
    Ts = cfg["Fs"] // cfg["Rs"]
    P = int(rng.choice([Ts] + [d for d in (2, 4, 5) if Ts % d == 0]))
    raw, _ = siggen.make_stream(seed, n_packets=int(rng.integers(1, 4)), ebno_db=ebno, framing=framing, fmt=fmt, clock_ppm=ppm)
    a = port.fsk(cfg["Fs"], cfg["Rs"], M=2, P=None if P == Ts else P); b = ref.fsk(cfg["Fs"], cfg["Rs"], M=2, P=None if P == Ts else P)
    sa, la, ca = a.run(raw, fmt); sb, lb, cb = b.run(raw, fmt)
    ok = (ca == cb and np.array_equal(sa.view(np.uint32), sb.view(np.uint32)) and np.array_equal(la[:, :3].view(np.uint32), lb[:, :3].view(np.uint32))
          and np.array_equal(la[:, 5:].view(np.uint32), lb[:, 5:].view(np.uint32)) and np.array_equal(a.state()[:4], b.state()[:4])
          and np.array_equal(a.fft_est().view(np.uint32), b.fft_est().view(np.uint32)))
    n += 1
    if not ok:
        bad += 1; print("MISMATCH", seed, framing, fmt, ebno, ppm, P, flush=True)
print("cases", n, "mismatches", bad, "in %.0f s" % (time.time() - t0))
