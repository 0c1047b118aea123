"""GPU parity at the BASELINE.json configuration sizes (configs[1], [2], [3], [4] per GPU), through properties that do
not need the oracle to process the full volume: a few seeded source streams / codewords are decoded by the CPU oracle,
the engine processes thousands of exact replicas of them resident in HBM, and EVERY replica must reproduce the
oracle's result for its source bit for bit (a checksum of checksums over the whole batch)."""
import hashlib

import numpy as np
import pytest

from wenet_b200 import siggen

pytestmark = pytest.mark.gpu
NCODE = 2580


@pytest.fixture(scope="module")
def E():
    from wenet_b200 import engine
    return engine


def _digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def _resident(E, n_streams, sources, nsamp, **kw):
    e = E.Engine(n_streams, chunk_samples=nsamp, **kw)
    e.feed(list(sources) + [None] * (n_streams - len(sources)))
    e.sync()
    e.dev_replicate(len(sources), nsamp, 0)          # stream s := source s % n_src, exact copy
    e.dev_set_fill(nsamp)
    e.process()
    e.sync()
    return e


def test_config2_fsk_only_1024_streams_x_1M(E, oracle_port):
    """configs[1]: FSK-demod only, 1024 streams x 1 Msample, 2-FSK 115.2 kbaud"""
    n, nsamp, n_src = 1024, 1 << 20, 6
    srcs = [siggen.make_stream(500 + i, n_samples=nsamp, ebno_db=6.0 + 2 * i, fmt="cf32",
                               clock_ppm=float((i - 3) * 900))[0] for i in range(n_src)]
    want = [_digest(oracle_port.fsk(921416, 115177).run(r, "cf32")[0]) for r in srcs]
    e = _resident(E, n, srcs, nsamp, in_fmt="cf32", framing="none")
    assert e.last_samples > 0.999 * n * nsamp
    for s in range(n):
        assert _digest(e.drain_soft(s)) == want[s % n_src], s
    e.close()


def test_config3_ldpc_only_1M_codewords(E, oracle_port):
    """configs[2]: LDPC only, 1M codewords, soft-LLR input, max_iter 100 (and the reference's own 10)"""
    rng = np.random.default_rng(77)
    n_src, n = 24, 1 << 20
    llr = []
    for k in range(n_src):
        data = rng.integers(0, 2, 2064, dtype=np.uint8)
        cw = np.concatenate([data, oracle_port.ldpc_encode(data)]).astype(np.float64)
        sd = (1 - 2 * cw) * rng.uniform(0.5, 2.0) + 10 ** (-(1.5 + 0.25 * k) / 20) * rng.standard_normal(NCODE)
        llr.append(oracle_port.sd_to_llr(sd.astype(np.float32).astype(np.float64)))
    llr = np.stack(llr)
    e = E.Engine(1, framing="v1", chunk_samples=4096)
    e.dev_ldpc_setup(llr, n)
    for max_iter in (100, 10):
        ref = [oracle_port.ldpc_decode(llr[k], max_iter, -1) for k in range(n_src)]
        e.dev_ldpc_run(max_iter)
        for first in range(0, n, 1 << 16):
            bits, iters, pcc = e.dev_ldpc_result(first, 1 << 16)
            idx = (first + np.arange(1 << 16)) % n_src
            assert np.array_equal(iters, np.array([ref[k][1] for k in range(n_src)], np.int32)[idx])
            assert np.array_equal(pcc, np.array([ref[k][2] for k in range(n_src)], np.int32)[idx])
            want = np.stack([ref[k][0] for k in range(n_src)])
            assert np.array_equal(bits, want[idx])
    assert len({r[1] for r in ref}) >= 3
    e.close()


def test_config4_end_to_end_4096_streams(E, oracle_port):
    """configs[3] per GPU: demod + deframe + LDPC, 4096 streams, Eb/N0 sweep 4-12 dB (one 256 Ki-sample chunk)"""
    n, nsamp = 4096, 1 << 18
    ebno = [4.0, 6.0, 8.0, 10.0, 12.0]
    srcs, want = [], []
    for i, eb in enumerate(ebno):
        raw, _ = siggen.make_stream(700 + i, n_samples=nsamp, ebno_db=eb, fmt="cf32", clock_ppm=float((i - 2) * 1200))
        srcs.append(raw)
        sd, _, _ = oracle_port.fsk(921416, 115177).run(raw, "cf32")
        res = oracle_port.deframer("v1", 10).feed(sd)
        want.append((_digest(sd), res["packets"], res["iters"].tolist()))
    e = _resident(E, n, srcs, nsamp, in_fmt="cf32", framing="v1")
    cw = e.drain_codewords()
    per_stream = np.bincount(cw["stream"], minlength=n)
    pk = e.drain_all_packets()
    pk_count = np.bincount(pk["stream"], minlength=n)
    for s in range(n):
        d, packets, iters = want[s % 5]
        assert per_stream[s] == len(iters), s
        assert pk_count[s] == len(packets) // 256, s
    # full content check on a stride of streams + digest check of the soft decisions
    starts = np.concatenate([[0], np.cumsum(per_stream)])
    pstarts = np.concatenate([[0], np.cumsum(pk_count)])
    for s in range(0, n, 37):
        d, packets, iters = want[s % 5]
        assert _digest(e.drain_soft(s)) == d, s
        assert cw["iters"][starts[s]:starts[s + 1]].tolist() == iters, s
        assert pk["payload"][pstarts[s]:pstarts[s + 1]].tobytes() == packets, s
    assert len(want[3][1]) >= 256 * 7 and len(want[0][1]) == 0      # 10 dB decodes, 4 dB does not
    e.close()


def test_config5_4fsk_1024_streams(E, oracle_port):
    """configs[4] per GPU: 4-FSK (fsk.c M = 4), 1024 streams (one 512 Ki-sample chunk)"""
    n, nsamp, n_src = 1024, 1 << 19, 4
    srcs = []
    for i in range(n_src):
        raw, _ = siggen.make_4fsk_stream(900 + i, nsamp // 8 + 16, ebno_db=8.0 + 2 * i)
        srcs.append(raw[:2 * nsamp])
    want = [_digest(oracle_port.fsk(921416, 115177, M=4).run(r, "cf32")[0]) for r in srcs]
    e = _resident(E, n, srcs, nsamp, M=4, in_fmt="cf32", framing="none")
    for s in range(n):
        assert _digest(e.drain_soft(s)) == want[s % n_src], s
    e.close()


def test_edge_cases(E, oracle_port):
    """empty feeds, sub-frame dribbles, NaN samples (the reference's NaN guard, src/fsk.c:878-880), saturated input"""
    e = E.Engine(3, in_fmt="cf32", framing="v1", chunk_samples=8192, hard_bits=True)
    e.feed([None, None, None]); e.process(); e.sync()
    assert e.last_samples == 0 and all(e.drain_soft(s).size == 0 for s in range(3))
    raw, _ = siggen.make_stream(5, n_samples=6000, ebno_db=10.0, fmt="cf32")
    raw = raw.copy()
    bad = raw.copy()
    bad[2 * 1000] = np.nan                                   # one NaN sample in frame 2
    sat = np.clip(raw * 1e30, -3e38, 3e38).astype(np.float32)
    ref = [oracle_port.fsk(921416, 115177).run(x, "cf32")[0] for x in (raw, bad, sat)]
    ref_bits = [oracle_port.fsk(921416, 115177).run_bits(x, "cf32") for x in (raw, bad, sat)]
    got, got_bits = [[], [], []], [[], [], []]
    pos = 0
    while pos < raw.size:                                    # 101-sample dribbles: most calls complete no frame
        n = 202
        e.feed([raw[pos:pos + n], bad[pos:pos + n], sat[pos:pos + n]])
        e.process(); e.sync()
        for s in range(3):
            got[s].append(e.drain_soft(s))
            got_bits[s].append(e.drain_hard(s))
        pos += n
    for s in range(3):
        g = np.concatenate(got[s])
        assert g.size == ref[s].size, s
        same = (g.view(np.uint32) == ref[s].view(np.uint32)) | (np.isnan(g) & np.isnan(ref[s]))
        assert same.all(), (s, np.nonzero(~same)[0][:5])
        # hard bits: a frame the NaN guard skips repeats the previous frame's bits (the caller's buffer is left as it was)
        assert np.array_equal(np.concatenate(got_bits[s]), ref_bits[s]), s
    assert set(e.nin().tolist()) <= {380, 384, 388}
    e.close()


# The four remaining 4-FSK instantiations of the kernel template (P < Ts, and 10 samples per symbol); the oracle is
# pinned for them against the compiled reference (tests/test_oracle_vs_ref.py::test_fsk_4fsk_other_geometries).  First seen
# to pass on a B200 by the round-1 driver run (XPASS x4 in GPUTEST_r01.json); a regression now fails the suite.
@pytest.mark.parametrize("Fs,Rs,P,fmt", [(921416, 115177, 4, "cu8"), (921416, 115177, 2, "cf32"), (960000, 96000, 5, "cs16"),
                                         (960000, 96000, 10, "cf32")])
def test_4fsk_other_geometries(E, oracle_port, Fs, Rs, P, fmt):
    raws = [siggen.make_4fsk_stream(41 + s, 2500, ebno_db=6.0 + 3 * s, fmt=fmt, Fs=Fs, Rs=Rs)[0] for s in range(3)]
    per = E.FMT_ELEMS[fmt]
    e = E.Engine(3, Fs=Fs, Rs=Rs, M=4, P=P, in_fmt=fmt, framing="none", chunk_samples=raws[0].size // per + 1024,
                 hard_bits=True)
    try:
        e.feed(raws)
        e.process()
        e.sync()
        for s in range(3):
            f = oracle_port.fsk(Fs, Rs, M=4, P=P)
            sd_o = f.run(raws[s], fmt)[0]
            assert np.array_equal(e.drain_soft(s).view(np.uint32), sd_o.view(np.uint32)), s
            assert np.array_equal(e.drain_hard(s), oracle_port.fsk(Fs, Rs, M=4, P=P).run_bits(raws[s], fmt)), s
    finally:
        e.close()
