"""Small run of every kernel for compute-sanitizer (memcheck / racecheck), e.g.
   compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import sys
import numpy as np
sys.path.insert(0, '.')
import os
from wenet_b200 import engine as E, siggen
if os.environ.get("WB_LIB"):            # e.g. the `san` build (no tensor-memory scratch) for synccheck
    E._lib = E.load_library(os.path.abspath(os.environ["WB_LIB"]))

def run(label, raws, **kw):
    e = E.Engine(len(raws), chunk_samples=1 << 17, stats=True, **kw)
    e.feed(raws); e.process(); e.sync()
    pk = [len(e.drain_packets(s)) // 256 for s in range(len(raws))]
    print(label, "packets", pk, "stats", round(e.stats(0).EbNodB, 2))
    e.close()

n = 70000
run("v1 cf32", [siggen.make_stream(s, n_samples=n, ebno_db=9.0, fmt="cf32", clock_ppm=1500.0 * (s - 1))[0]
                for s in range(3)], in_fmt="cf32", framing="v1")
run("v1 cu8", [siggen.make_stream(s, n_samples=n, ebno_db=9.0, fmt="cu8")[0] for s in range(2)],
    in_fmt="cu8", framing="v1")
run("v2 cs16", [siggen.make_stream(s, n_samples=n, ebno_db=9.0, fmt="cs16", framing="v2")[0]
                for s in range(2)], in_fmt="cs16", framing="v2", Fs=960000, Rs=96000)
raws = [siggen.make_4fsk_stream(s, 1200, ebno_db=10.0)[0] for s in range(2)]
e = E.Engine(2, M=4, in_fmt="cf32", framing="none", chunk_samples=16384)
e.feed(raws); e.process(); e.sync()
print("4fsk sd", [e.drain_soft(s).size for s in range(2)])
e.close()
llr = np.random.default_rng(0).standard_normal((4, 2580)).astype(np.float32) * 3
e = E.Engine(1, framing="v1", chunk_samples=4096)
print("ldpc iters", e.ldpc_decode_batch(llr, 10)[1])
e.close()
# transmit side on the device, then the engine decodes its own signal
pl = np.random.default_rng(1).integers(0, 256, size=(3, 2, 256), dtype=np.uint8)
for fmt, fr, cfg in (("cf32", "v1", siggen.V1), ("cs16", "v2", siggen.V2)):
    e = E.Engine(3, Fs=cfg["Fs"], Rs=cfg["Rs"], in_fmt=fmt, framing=fr, chunk_samples=1 << 17)
    e.tx_synthesize(pl, int(cfg["f_lo"]), int(cfg["f_hi"] - cfg["f_lo"]), ebno_db=10.0, seed=3)
    e.process(); e.sync()
    print("tx", fmt, fr, "packets", [len(e.drain_packets(s)) // 256 for s in range(3)])
    e.close()
