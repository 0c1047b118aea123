"""Randomised sweep of the oracle restatement against the compiled reference (oracle/_ref): sd_to_llr (incl. the x87 emulation) and
the LDPC decoder at max_iter 10 / 100 over random signal and noise scales.
    python tools/fuzz/oracle_ldpc.py SECONDS
Round 2: 7 634 codewords in 180 s, 0 mismatches (LLRs bit for bit, decoded bits, iteration and parity-check counts)."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
O.build()
port, ref = O.Oracle("port"), O.Oracle("reference")
rng = np.random.default_rng(77)
t0 = time.time(); n = bad = 0
while time.time() - t0 < float(sys.argv[1]):
    data = rng.integers(0, 2, 2064, dtype=np.uint8)
    if rng.random() < 0.1: data[:] = 0
    par = ref.ldpc_encode(data)
    cw = np.concatenate([data, par]).astype(np.float64)
    sd = (1 - 2 * cw) * rng.uniform(0.05, 300) + rng.uniform(0.05, 300) * 10 ** (-rng.uniform(-6, 10) / 20) * rng.standard_normal(2580)
    sd32 = sd.astype(np.float32).astype(np.float64)
    la, lb = port.sd_to_llr(sd32), ref.sd_to_llr(sd32)
    mi = int(rng.choice([10, 100]))
    a, b = port.ldpc_decode(lb, mi, -1), ref.ldpc_decode(lb, mi, -1)
    ok = np.array_equal(np.asarray(la, dtype=np.float32).view(np.uint32), np.asarray(lb, dtype=np.float32).view(np.uint32)) and a[1] == b[1] and a[2] == b[2] and np.array_equal(a[0], b[0])
    n += 1
    if not ok: bad += 1; print("MISMATCH case", n, flush=True)
print("cases", n, "mismatches", bad)
