#!/usr/bin/env python3
"""Per-phase cycle split of wb_fsk_kernel (debug library, `make -C wenet_b200/csrc dbg`): thread 0 of every CTA
stamps clock() after each CTA barrier; inside the mixer phase B1 lane 0 of the last oscillator warp stamps the way to its
segment (entry -> table loads -> straight-line steps and switch -> rest of the bare recurrence) and every oscillator warp
the moment it is done (cycles from the phase start).  Sums over all CTAs and frames.  Not a benchmark: the counters
perturb the run.  N=2072 gives one CTA per SM, the default 4096 streams two."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wenet_b200 import engine as E      # noqa: E402
import bench                            # noqa: E402

lib = E.load_library(os.path.join(ROOT, "wenet_b200", "libwenet_b200_dbg.so"))
E._lib = lib
n, chunk = int(os.environ.get("N", 4096)), int(os.environ.get("CHUNK", 1 << 20))
src = bench.make_sources(40, chunk, 0)
eng = E.Engine(n, in_fmt="cf32", chunk_samples=chunk, framing="v1")
eng.feed(src + [None] * (n - 40))
eng.sync()
eng.dev_replicate(40, chunk, 4096 + 16 * 37)
out = (C.c_ulonglong * 12)()
for it in range(3):
    eng.dev_set_fill(chunk)
    eng.process()
    eng.sync()
    lib.wb_debug_phase_clk(out)
    v = np.array(list(out), dtype=np.float64)
    ms = eng.last_kernel_ms()
    frames = chunk / 384.0
    ctas = (n + 13) // 14
    names = ["A", "B1", "B2", "B3", "(last B1 warp: entry->tables", "pre-switch steps", "C+loop", "rest of spin)", "w0end", "w1end", "w2end", "w3end"]
    print("fsk %.2f ms; cycles per CTA-frame: " % ms[0] +
          ", ".join("%s %.0f" % (nm, x / ctas / frames) for nm, x in zip(names, v)) +
          "; total %.0f" % ((v[:4].sum() + v[6]) / ctas / frames))
eng.close()
