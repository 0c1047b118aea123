"""GPU parity tests: the CUDA path, called through the C ABI (wenet_b200.engine -> libwenet_b200.so),
against the CPU oracle (oracle/liboracle.so, itself pinned to the compiled reference by
tests/test_oracle_vs_ref.py and the committed golden vectors) on the same seeded inputs.

Bars: soft decisions, LLRs, iteration counts, parity counts, nin sequence and packet bytes are all
BIT-EXACT (the north-star tolerance for LLRs is 1e-4 relative; we hold them to zero difference).
"""
import os

import numpy as np
import pytest

from wenet_b200 import siggen

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NCODE = 2580


@pytest.fixture(scope="module")
def eng_mod():
    from wenet_b200 import engine
    return engine


def _noisy_codewords(oracle, n, snr_db_list, seed):
    """random valid codewords -> BPSK + AWGN soft decisions -> sd_to_llr (oracle) -> LLR rows"""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        data = rng.integers(0, 2, 2064, dtype=np.uint8)
        if k % 7 == 3:
            data[:] = 0                       # the all-zero-data early exit of run_ldpc_decoder
        par = oracle.ldpc_encode(data)
        cw = np.concatenate([data, par]).astype(np.float64)
        snr = snr_db_list[k % len(snr_db_list)]
        sigma = 10.0 ** (-snr / 20.0)
        sd = (1.0 - 2.0 * cw) * rng.uniform(0.3, 3.0) + sigma * rng.standard_normal(NCODE)
        out.append(oracle.sd_to_llr(sd.astype(np.float32).astype(np.float64)))
    return np.stack(out)


def test_ldpc_known_answer(eng_mod):
    """reference src/H2064_516_sparse.h:27-33: input[] decodes to detected_data[] in 8 iterations"""
    z = np.load(os.path.join(GOLD, "ldpc_kat.npz"))
    e = eng_mod.Engine(1, framing="v1", chunk_samples=4096)
    bits, iters, pcc = e.ldpc_decode_batch(z["llr"][None, :], max_iter=int(z["max_iter"]))
    assert iters[0] == int(z["iters"]) == 8
    assert pcc[0] == int(z["parity_ok"]) == 516
    assert np.array_equal(bits[0], z["detected"].astype(np.uint8))
    e.close()


@pytest.mark.parametrize("max_iter", [10, 100])
def test_ldpc_random_vs_oracle(eng_mod, oracle_port, max_iter):
    n = 96 if max_iter == 10 else 24
    llr = _noisy_codewords(oracle_port, n, [1.0, 2.0, 3.0, 4.0, 6.0, 9.0], seed=7 + max_iter)
    llr[5, 100] = np.float32(40000.0)         # phi0 overflow quirk (x >= 32768 -> 10.0)
    llr[6, 7] = np.float32(0.0)
    e = eng_mod.Engine(1, framing="v1", chunk_samples=4096)
    bits, iters, pcc = e.ldpc_decode_batch(llr, max_iter=max_iter)
    e.close()
    for k in range(n):
        b, it, pc = oracle_port.ldpc_decode(llr[k], max_iter=max_iter, pcc_init=-1)
        assert it == iters[k], (k, it, iters[k])
        assert pc == pcc[k], (k, pc, pcc[k])
        assert np.array_equal(b, bits[k]), k
    assert len(set(iters.tolist())) > 2       # the sweep really exercises different iteration counts


def test_sd_to_llr_vs_oracle(eng_mod, oracle_port):
    rng = np.random.default_rng(11)
    sd = (rng.choice([-1.0, 1.0], size=(40, NCODE)) * rng.uniform(0.1, 30.0, size=(40, 1))
          + rng.standard_normal((40, NCODE)) * rng.uniform(0.01, 2.0, size=(40, 1))).astype(np.float32)
    sd[3, :50] = 0.0
    sd[4] *= 1e-6
    e = eng_mod.Engine(1, framing="v1", chunk_samples=4096)
    llr = e.sd_to_llr_batch(sd)
    e.close()
    for k in range(sd.shape[0]):
        ref = oracle_port.sd_to_llr(sd[k].astype(np.float64))
        assert np.array_equal(ref.view(np.uint32), llr[k].view(np.uint32)), k


def _run_oracle_stream(oracle, raw, fmt, Fs, Rs, M, framing, max_iter=10, P=None):
    f = oracle.fsk(Fs, Rs, M=M, P=P)
    sd, log, consumed = f.run(raw, fmt)
    res = None
    if framing:
        res = oracle.deframer(framing, max_iter).feed(sd)
    return sd, log, consumed, res


STREAM_CASES = [
    # (id, fmt, ebno, ppm, n_packets)
    ("cf32_10dB", "cf32", 10.0, 0.0, 3),
    ("cf32_8dB_ppm", "cf32", 8.0, 2500.0, 3),
    ("cs16_9dB_negppm", "cs16", 9.0, -3000.0, 2),
    ("cu8_12dB", "cu8", 12.0, 0.0, 2),
    ("cf32_5dB", "cf32", 5.0, 0.0, 2),
    ("s16_real_14dB_ppm", "s16", 14.0, 1500.0, 2),      # fsk_demod without -c/-d: real samples, imag = 0 (src/fsk_demod.c:289-295)
]


@pytest.mark.parametrize("case", STREAM_CASES, ids=[c[0] for c in STREAM_CASES])
def test_stream_v1_vs_oracle(eng_mod, oracle_port, case):
    """FSK soft decisions, nin sequence, LLRs, iterations and packets of one v1 stream, one chunk"""
    _, fmt, ebno, ppm, npk = case
    raw, payloads = siggen.make_stream(3, n_packets=npk, ebno_db=ebno, fmt=fmt, clock_ppm=ppm)
    cfg = siggen.V1
    sd_o, log_o, cons_o, res_o = _run_oracle_stream(oracle_port, raw, fmt, cfg["Fs"], cfg["Rs"], 2, "v1")
    nsamp = raw.size // eng_mod.FMT_ELEMS[fmt]
    e = eng_mod.Engine(1, Fs=cfg["Fs"], Rs=cfg["Rs"], in_fmt=fmt, framing="v1", chunk_samples=nsamp + 1024, keep_llr=True)
    e.enable_frame_log(len(log_o) + 4)
    e.feed([raw])
    e.process()
    e.sync()
    sd_g = e.drain_soft(0)
    assert e.last_samples == cons_o
    assert sd_g.size == sd_o.size
    log_g = e.read_frame_log(0, len(log_o))
    assert np.array_equal(log_g[:, 0], log_o[:, 0]), "nin sequence"
    binw = np.float32(cfg["Fs"]) / np.float32(256)
    assert np.array_equal(log_g[:, 1] * binw, log_o[:, 1]) and np.array_equal(log_g[:, 2] * binw, log_o[:, 2]), "f_est"
    assert np.array_equal(log_g[:, 5].view(np.uint32), log_o[:, 5].view(np.uint32)), "norm_rx_timing"
    assert np.array_equal(log_g[:, 6].view(np.uint32), log_o[:, 6].view(np.uint32)), "ppm"
    bad = np.nonzero(sd_g.view(np.uint32) != sd_o.view(np.uint32))[0]
    assert bad.size == 0, "soft decisions differ first at %s" % bad[:5]
    cw, llr = e.drain_codewords(with_llr=True)
    assert len(cw) == len(res_o["iters"])
    assert np.array_equal(llr.view(np.uint32), res_o["llr"].view(np.uint32))
    assert np.array_equal(cw["iters"], res_o["iters"])
    assert np.array_equal(cw["crc_ok"], res_o["crc_ok"])
    assert np.array_equal(cw["bytes"], res_o["bytes258"])
    pk = e.drain_packets(0)
    assert pk == res_o["packets"]
    if ebno >= 9.0:
        assert pk == b"".join(payloads)
    e.close()


def test_stream_chunked_feed_matches_one_shot(eng_mod, oracle_port):
    """feeding ragged chunks (state carried in HBM, half-collected packets carried over) == one pass"""
    cfg = siggen.V1
    raws, refs = [], []
    for s in range(5):
        raw, _ = siggen.make_stream(20 + s, n_packets=3, ebno_db=9.0 + s, fmt="cs16", clock_ppm=(-2000 + 1000 * s))
        raws.append(raw)
        refs.append(_run_oracle_stream(oracle_port, raw, "cs16", cfg["Fs"], cfg["Rs"], 2, "v1"))
    e = eng_mod.Engine(5, Fs=cfg["Fs"], Rs=cfg["Rs"], in_fmt="cs16", framing="v1", chunk_samples=40000)
    rng = np.random.default_rng(5)
    pos = [0] * 5
    sds = [[] for _ in range(5)]
    pk = [b""] * 5
    while any(pos[s] < raws[s].size for s in range(5)):
        chunk = []
        for s in range(5):
            n = int(rng.integers(0, 30000)) * 2
            if rng.random() < 0.15:
                n = 0
            chunk.append(raws[s][pos[s]:pos[s] + n])
            pos[s] += len(chunk[-1])
        e.feed(chunk)
        e.process()
        e.sync()
        for s in range(5):
            sds[s].append(e.drain_soft(s))
            pk[s] += e.drain_packets(s)
    for s in range(5):
        sd = np.concatenate(sds[s])
        assert np.array_equal(sd.view(np.uint32), refs[s][0].view(np.uint32)), s
        assert pk[s] == refs[s][3]["packets"], s
        assert len(pk[s]) == 3 * 256
    e.close()


def test_stream_v2_vs_oracle(eng_mod, oracle_port):
    cfg = siggen.V2
    raws, refs = [], []
    for s in range(3):
        raw, payloads = siggen.make_stream(40 + s, n_packets=2, ebno_db=9.0 + 2 * s, framing="v2", fmt="cf32",
                                           clock_ppm=1500.0 * s)
        raws.append(raw)
        refs.append((_run_oracle_stream(oracle_port, raw, "cf32", cfg["Fs"], cfg["Rs"], 2, "v2"), payloads))
    n = max(r.size for r in raws) // 2
    e = eng_mod.Engine(3, Fs=cfg["Fs"], Rs=cfg["Rs"], in_fmt="cf32", framing="v2", chunk_samples=n + 1024, keep_llr=True)
    e.feed(raws)
    e.process()
    e.sync()
    cw, llr = e.drain_codewords(with_llr=True)
    k = 0
    for s in range(3):
        (sd_o, log_o, cons_o, res_o), payloads = refs[s]
        sd_g = e.drain_soft(s)
        assert np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32)), s
        ncw = len(res_o["iters"])
        assert np.array_equal(cw["stream"][k:k + ncw], np.full(ncw, s))
        assert np.array_equal(llr[k:k + ncw].view(np.uint32), res_o["llr"].view(np.uint32))
        assert np.array_equal(cw["iters"][k:k + ncw], res_o["iters"])
        k += ncw
        pk = e.drain_packets(s)
        assert pk == res_o["packets"] == b"".join(payloads)
    assert k == len(cw)
    e.close()


def test_4fsk_soft_vs_oracle(eng_mod, oracle_port):
    raws, refs = [], []
    for s in range(3):
        raw, sym = siggen.make_4fsk_stream(60 + s, 3000, ebno_db=8.0 + 3 * s)
        raws.append(raw)
        refs.append(oracle_port.fsk(921416, 115177, M=4).run(raw, "cf32"))
    e = eng_mod.Engine(3, M=4, in_fmt="cf32", framing="none", chunk_samples=raws[0].size // 2 + 1024)
    e.feed(raws)
    e.process()
    e.sync()
    for s in range(3):
        sd_g = e.drain_soft(s)
        assert np.array_equal(sd_g.view(np.uint32), refs[s][0].view(np.uint32)), s
    e.close()


@pytest.mark.parametrize("M,fmt,Fs,Rs", [(4, "cf32", 921416, 115177), (4, "cu8", 921416, 115177),
                                         (2, "cs16", 921416, 115177), (2, "cf32", 960000, 96000)])
def test_hard_bits_vs_oracle(eng_mod, oracle_port, M, fmt, Fs, Rs):
    """WB_FLAG_HARD_BITS / wb_drain_hard: rx_bits of fsk_demod() (reference src/fsk.c:936-959, what `fsk_demod` writes
    without -s) next to unchanged soft decisions, fed in uneven pieces so that frames straddle process() calls"""
    raws = []
    for s in range(3):
        if M == 4:
            raw, _ = siggen.make_4fsk_stream(80 + s, 3000, ebno_db=4.0 + 3 * s, fmt=fmt)
        else:
            raw, _ = siggen.make_stream(80 + s, n_packets=1, ebno_db=3.0 + 3 * s, fmt=fmt, clock_ppm=2000.0 * (s - 1),
                                        framing="v1" if Fs == 921416 else "v2")
        raws.append(raw)
    per = 2 if fmt != "s16" else 1
    nmin = min(r.size for r in raws) // per
    raws = [r[:nmin * per] for r in raws]
    refs_b = [oracle_port.fsk(Fs, Rs, M=M).run_bits(r, fmt) for r in raws]
    refs_s = [oracle_port.fsk(Fs, Rs, M=M).run(r, fmt)[0] for r in raws]
    e = eng_mod.Engine(3, Fs=Fs, Rs=Rs, M=M, in_fmt=fmt, framing="none", chunk_samples=nmin // 3 + 2048, hard_bits=True)
    got_b, got_s = [[] for _ in raws], [[] for _ in raws]
    cuts = [0, nmin // 3 - 77, 2 * nmin // 3 + 131, nmin]
    for a, b in zip(cuts[:-1], cuts[1:]):
        e.feed([r[a * per:b * per] for r in raws])
        e.process()
        e.sync()
        for s in range(3):
            got_b[s].append(e.drain_hard(s))
            got_s[s].append(e.drain_soft(s))
    for s in range(3):
        hb, sd = np.concatenate(got_b[s]), np.concatenate(got_s[s])
        assert hb.size == refs_b[s].size > 1000 and np.array_equal(hb, refs_b[s]), s
        assert np.array_equal(sd.view(np.uint32), refs_s[s].view(np.uint32)), s
    if M == 4:      # the arg-max decisions are not the signs of the 4-FSK soft decisions
        assert any(np.any((np.concatenate(got_s[s]) > 0) != np.concatenate(got_b[s])) for s in range(3))
    e.close()
    # without the flag the tap refuses instead of returning stale bytes
    e = eng_mod.Engine(1, Fs=Fs, Rs=Rs, M=M, in_fmt=fmt, framing="none", chunk_samples=4096)
    with pytest.raises(eng_mod.WbError):
        e.drain_hard(0)
    e.close()


def test_hard_bits_golden_and_cli(eng_mod):
    """the reference CLI's own hard-bit bytes (tests/golden/fsk_hard.npz) through the engine and through the
    `fsk_demod` stand-in without -s"""
    import subprocess
    import sys
    z = np.load(os.path.join(GOLD, "fsk_hard.npz"))
    for M, fmt, raw, bits in ((4, "cu8", z["raw4"], z["bits4"]), (2, "cs16", z["raw2"], z["bits2"])):
        e = eng_mod.Engine(1, M=M, in_fmt=fmt, framing="none", chunk_samples=raw.size // 2 + 1024, hard_bits=True)
        e.feed([raw])
        e.process()
        e.sync()
        assert np.array_equal(e.drain_hard(0), bits), M
        e.close()
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-m", "wenet_b200.cli.fsk_demod", "--cu8", "4", "921416", "115177", "-", "-"],
                         input=z["raw4"].tobytes(), stdout=subprocess.PIPE, env=env, check=True, timeout=300).stdout
    assert out == z["bits4"].tobytes()


def test_cli_start_rx_command_line():
    """the stage-1 command of start_rx.sh:126 as it is written there (--stats=100: an option with an optional argument):
    the reference's soft decisions on stdout, JSON lines with the keys rx/fskstatsudp.py:214-226 reads on stderr"""
    import json
    import subprocess
    import sys
    z = np.load(os.path.join(GOLD, "fsk_v1.npz"))
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-m", "wenet_b200.cli.fsk_demod", "--cu8", "-s", "--stats=100", "2", "921416", "115177",
                        "-", "-"], input=z["raw"].tobytes(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env,
                       check=True, timeout=300)
    assert r.stdout == z["sd"].tobytes()
    recs = [json.loads(l) for l in r.stderr.decode().splitlines() if l.startswith("{")]
    assert recs
    for k in ("EbNodB", "ppm", "f1_est", "f2_est", "samp_fft", "eye_diagram"):
        assert k in recs[-1], k
    # the reference's own estimates over the last 60 frames of this stream (oracle/_ref): f1 wanders between estimator
    # bins 39 and 40 (of 3599.28 Hz), f2 sits in bin 76; a record is written once per stats period, not at the last frame
    assert len(recs[-1]["samp_fft"]) == 128 and len(recs) >= 6
    assert abs(recs[-1]["f1_est"] - 140372.0) < 3700 and abs(recs[-1]["f2_est"] - 273545.4) < 3700


def test_cli_testframe_mode():
    """`fsk_demod -f` through the stand-in: the reference CLI's hard bits on stdout and its "errs: ..." lines on stderr
    (tests/golden/testframes.npz, written by the reference binary), hard and soft output modes"""
    import subprocess
    import sys
    z = np.load(os.path.join(GOLD, "testframes.npz"))
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cmd = [sys.executable, "-m", "wenet_b200.cli.fsk_demod", "--cs16", "-f"]
    r = subprocess.run(cmd + ["2", "921416", "115177", "-", "-"], input=z["raw"].tobytes(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, check=True, timeout=300)
    assert r.stdout == z["bits"].tobytes()
    assert r.stderr.decode() == z["stderr"].tobytes().decode()
    r = subprocess.run(cmd + ["-s", "2", "921416", "115177", "-", "-"], input=z["raw"].tobytes(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, check=True, timeout=300)
    assert r.stderr.decode().count("errs:") == 40 and "bits tested 4000" in r.stderr.decode()
    r = subprocess.run(cmd + ["-t", "2", "921416", "115177", "-", "-"], input=z["raw"].tobytes(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, check=True, timeout=300)
    import json
    recs = [json.loads(l) for l in r.stderr.decode().splitlines()]
    assert recs and recs[-1]["frames"] == 40 and recs[-1]["bits"] == 4000 and recs[-1]["errs"] == 184
    assert "eye_diagram" not in recs[-1]
    assert abs(recs[-1]["f1_est"] - 140372.0) < 3700 and abs(recs[-1]["f2_est"] - 273545.4) < 3700   # the reference's bins 39 / 76 +- 1


def test_fine_timing_generic_path_vs_oracle(eng_mod, oracle_port, monkeypatch):
    """the fine-timing chain without the periodic-table shortcut (what wb_create selects when the host finds the
    reference's phi_ft recurrence not periodic) gives the same soft decisions and nin sequence"""
    cfg = siggen.V1
    raw, _ = siggen.make_stream(11, n_packets=2, ebno_db=7.0, fmt="cf32", clock_ppm=1800.0)
    sd_o, log_o, _, _ = _run_oracle_stream(oracle_port, raw, "cf32", cfg["Fs"], cfg["Rs"], 2, None)
    monkeypatch.setenv("WB_FSK_PFT_GENERIC", "1")
    e = eng_mod.Engine(1, in_fmt="cf32", framing="none", chunk_samples=raw.size // 2 + 1024)
    monkeypatch.delenv("WB_FSK_PFT_GENERIC")
    e.enable_frame_log(len(log_o) + 4)
    e.feed([raw])
    e.process()
    e.sync()
    assert np.array_equal(e.drain_soft(0).view(np.uint32), sd_o.view(np.uint32))
    assert np.array_equal(e.read_frame_log(0, len(log_o))[:, 0], log_o[:, 0]), "nin sequence"
    e.close()


def test_many_streams_one_launch(eng_mod, oracle_port):
    """more streams than one CTA holds, ragged lengths, some empty: every stream still matches"""
    cfg = siggen.V1
    n = 37
    raws = []
    for s in range(n):
        if s % 9 == 4:
            raws.append(np.zeros(0, dtype=np.float32))
            continue
        raw, _ = siggen.make_stream(100 + s, n_samples=30000 + 997 * s, ebno_db=6.0 + (s % 7), fmt="cf32",
                                    clock_ppm=float((s % 5 - 2) * 1500))
        raws.append(raw)
    e = eng_mod.Engine(n, in_fmt="cf32", framing="v1", chunk_samples=80000)
    e.feed(raws)
    e.process()
    e.sync()
    for s in range(n):
        sd_g = e.drain_soft(s)
        if raws[s].size == 0:
            assert sd_g.size == 0
            continue
        sd_o, _, _, res_o = _run_oracle_stream(oracle_port, raws[s], "cf32", cfg["Fs"], cfg["Rs"], 2, "v1")
        assert np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32)), s
        assert e.drain_packets(s) == res_o["packets"], s
    e.close()


def test_process_soft_vs_oracle(eng_mod, oracle_port):
    """the drs232_ldpc / wenet_ldpc entry point: soft symbols in (ragged blocks, packets straddling calls)"""
    for framing, cfg in (("v1", siggen.V1), ("v2", siggen.V2)):
        rng = np.random.default_rng(9)
        sds = []
        for s in range(3):
            raw, _ = siggen.make_stream(70 + s, n_packets=3, ebno_db=8.0 + s, framing=framing, fmt="cf32")
            sd, _, _ = oracle_port.fsk(cfg["Fs"], cfg["Rs"]).run(raw, "cf32")
            sds.append(np.concatenate([sd, rng.standard_normal(3000).astype(np.float32)]))
        refs = [oracle_port.deframer(framing, 10).feed(sd) for sd in sds]
        e = eng_mod.Engine(3, Fs=cfg["Fs"], Rs=cfg["Rs"], framing=framing, chunk_samples=1 << 16)
        pos, got, iters = [0] * 3, [b""] * 3, [[] for _ in range(3)]
        while any(pos[s] < sds[s].size for s in range(3)):
            blk = []
            for s in range(3):
                n = int(rng.integers(0, 5000))
                blk.append(sds[s][pos[s]:pos[s] + n])
                pos[s] += len(blk[-1])
            e.process_soft(blk)
            e.sync()
            for cw in e.drain_codewords():
                iters[cw["stream"]].append(int(cw["iters"]))
            for s in range(3):
                got[s] += e.drain_packets(s)
        for s in range(3):
            assert got[s] == refs[s]["packets"], (framing, s)
            assert iters[s] == refs[s]["iters"].tolist(), (framing, s)
        e.close()


def test_cli_pipe_matches_reference_bytes():
    """python -m wenet_b200.cli.fsk_demod --cu8 -s ... | python -m wenet_b200.cli.drs232_ldpc - -  against the bytes
    the reference CLIs produced for the same input (tests/golden/fsk_v1.npz, made by tools/gen_golden.py)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    z = np.load(os.path.join(GOLD, "fsk_v1.npz"))
    sd = subprocess.run([sys.executable, "-m", "wenet_b200.cli.fsk_demod", "--cu8", "-s", "2", "921416", "115177", "-", "-"],
                        input=z["raw"].tobytes(), stdout=subprocess.PIPE, env=env, check=True, timeout=300).stdout
    assert sd == z["sd"].tobytes()
    r = subprocess.run([sys.executable, "-m", "wenet_b200.cli.drs232_ldpc", "-", "-", "-v"], input=sd, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, check=True, timeout=300)
    assert r.stdout == z["packets"].tobytes()
    assert b"packets: 2 packet_errors: 0 PER: 0.000" in r.stderr
    fused = subprocess.run([sys.executable, "-m", "wenet_b200.cli.wenet_rx", "--cu8", "-", "-"], input=z["raw"].tobytes(),
                           stdout=subprocess.PIPE, env=env, check=True, timeout=300).stdout
    assert fused == z["packets"].tobytes()
    z2 = np.load(os.path.join(GOLD, "fsk_v2.npz"))
    sd2 = subprocess.run([sys.executable, "-m", "wenet_b200.cli.fsk_demod", "--cs16", "-s", "2", "960000", "96000", "-", "-"],
                         input=z2["raw"].tobytes(), stdout=subprocess.PIPE, env=env, check=True, timeout=300).stdout
    assert sd2 == z2["sd"].tobytes()
    pk2 = subprocess.run([sys.executable, "-m", "wenet_b200.cli.wenet_ldpc", "-", "-"], input=sd2, stdout=subprocess.PIPE,
                         env=env, check=True, timeout=300).stdout
    assert pk2 == z2["packets"].tobytes()


def test_golden_vectors_through_the_engine(eng_mod):
    """the committed reference outputs, straight against the CUDA path (no oracle in between)"""
    z = np.load(os.path.join(GOLD, "ldpc_llr.npz"))
    e = eng_mod.Engine(1, framing="v1", chunk_samples=4096)
    llr = e.sd_to_llr_batch(z["sd"])
    assert np.array_equal(llr.view(np.uint32), z["llr"].view(np.uint32))
    for mi in (10, 100):
        bits, iters, pcc = e.ldpc_decode_batch(z["llr"], max_iter=mi)
        assert np.array_equal(iters, z["iters%d" % mi]) and np.array_equal(pcc, z["pcc%d" % mi])
        assert np.array_equal(np.packbits(bits, axis=1), z["bits%d" % mi])
    e.close()
    z4 = np.load(os.path.join(GOLD, "fsk_4fsk.npz"))
    e = eng_mod.Engine(1, M=4, in_fmt="cu8", framing="none", chunk_samples=z4["raw"].size // 2 + 1024)
    e.feed([z4["raw"]])
    e.process()
    e.sync()
    assert np.array_equal(e.drain_soft(0).view(np.uint32), z4["sd"].view(np.uint32))
    e.close()


def test_stats_surface(eng_mod, oracle_port):
    """wb_get_stats (the stderr JSON fields of src/fsk_demod.c:345-401): f_est, ppm, timing exact; snr_est (EbNodB) is the
    reference's 0.5/0.5 IIR replayed over the last 32 frames; eye diagram where the reference's own indexing is defined"""
    cfg = siggen.V1
    raw, _ = siggen.make_stream(90, n_packets=2, ebno_db=11.0, fmt="cf32", clock_ppm=800.0)
    f = oracle_port.fsk(cfg["Fs"], cfg["Rs"])
    sd_o, log_o, _ = f.run(raw, "cf32")
    so = f.state()
    e = eng_mod.Engine(1, in_fmt="cf32", framing="none", chunk_samples=raw.size // 2 + 1024, stats=True)
    e.feed([raw]); e.process(); e.sync()
    st = e.stats(0)
    assert st.f_est[0] == so[8] and st.f_est[1] == so[9]
    assert np.float32(st.ppm) == so[13] and np.float32(st.norm_rx_timing) == so[12] and np.float32(st.rx_timing) == so[16]
    assert np.float32(st.foff) == so[17]
    assert abs(st.EbNodB - so[15]) < 1e-4, (st.EbNodB, so[15])
    assert st.frames == len(log_o) and st.nfft == 128
    assert np.array_equal(np.array(st.samp_fft[:128], dtype=np.float32).view(np.uint32), f.fft_est().view(np.uint32))
    eye_o = f.eye()
    eye_g = np.array([[st.rx_eye[i][j] for j in range(st.neyesamp)] for i in range(st.neyetr)], dtype=np.float32)
    assert eye_g.shape == eye_o.shape == (8, 16)
    if so[16] >= -1.0:            # high_sample + 1 >= 0: the reference's indices stay inside f_int
        assert np.allclose(eye_g, eye_o, rtol=0, atol=1e-6)
    e.close()


# ---- transmit side on the device ----

@pytest.mark.gpu
@pytest.mark.parametrize("framing", ["v1", "v2"])
def test_tx_frames_and_modulator_vs_oracle(eng_mod, oracle_port, framing):
    """wb_tx_synthesize without noise: the on-air bits equal the TX restatement's (which the reference receiver accepts)
    and the samples equal the reference modulator's arithmetic bit for bit; then the engine decodes its own signal"""
    cfg = siggen.V1 if framing == "v1" else siggen.V2
    n, npk = 5, 3
    rng = np.random.default_rng(21)
    pl = rng.integers(0, 256, size=(n, npk, 256), dtype=np.uint8)
    f1, fs = int(cfg["f_lo"]), int(cfg["f_hi"] - cfg["f_lo"])
    e = eng_mod.Engine(n, Fs=cfg["Fs"], Rs=cfg["Rs"], in_fmt="cf32", framing=framing, chunk_samples=200000)
    ns = e.tx_synthesize(pl, f1, fs, ebno_db=None, lead_in=2000, gap=24, tail=400)
    for s in range(n):
        parts = [np.ones(2000, np.uint8)]
        for k in range(npk):
            parts += [oracle_port.tx_frame_bits(pl[s, k].tobytes(), framing), np.ones(24, np.uint8)]
        bits = np.concatenate(parts + [np.ones(400, np.uint8)])
        bits = np.concatenate([bits, np.ones((-bits.size) % 48, np.uint8)])
        assert np.array_equal(e.tx_read_bits(s), bits), s
        x = oracle_port.fsk_mod(bits, cfg["Fs"], cfg["Rs"], f1, fs)
        assert ns * 2 == x.size
        y = e.dev_read_input(s, ns)
        assert np.array_equal(y.view(np.uint32), x.view(np.uint32)), s
    e.process()
    e.sync()
    for s in range(n):
        assert e.drain_packets(s) == pl[s].tobytes(), s
    e.close()


@pytest.mark.gpu
def test_tx_golden_modulator(eng_mod):
    """the device modulator against the reference's own fsk_mod_c output (tests/golden/tx.npz), 2-FSK and 4-FSK"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tx.npz"))
    for M, bits, mod, f1, fs in ((2, g["bits2"], g["mod2"], 129763, 143594), (4, g["bits4"], g["mod4"], 46071, 115177)):
        nblk = bits.size // 2048 + 1
        pad = np.concatenate([bits, np.ones(nblk * 2048 - bits.size, np.uint8)])
        pl = np.packbits(pad).reshape(1, nblk, 256)
        e = eng_mod.Engine(1, M=M, in_fmt="cf32", framing="none", chunk_samples=100000)
        ns = e.tx_synthesize(pl, f1, fs, ebno_db=None, lead_in=0, gap=0, tail=0)
        y = e.dev_read_input(0, ns)
        assert np.array_equal(y[:mod.size].view(np.uint32), mod.view(np.uint32)), M
        e.close()


@pytest.mark.gpu
def test_tx_noisy_round_trip(eng_mod, oracle_port):
    """AWGN + peak normalisation on the device (benchmarking/generate_lowsnr.py): 10 dB streams decode completely,
    4 dB streams do not; the oracle demodulates the very same samples to the very same soft decisions; other input
    formats are converted on the device"""
    cfg = siggen.V1
    n, npk = 6, 4
    rng = np.random.default_rng(22)
    pl = rng.integers(0, 256, size=(n, npk, 256), dtype=np.uint8)
    f1, fs = int(cfg["f_lo"]), int(cfg["f_hi"] - cfg["f_lo"])
    e = eng_mod.Engine(n, in_fmt="cf32", framing="v1", chunk_samples=200000)
    ns = e.tx_synthesize(pl, f1, fs, ebno_db=10.0, seed=5)
    ys = [e.dev_read_input(s, ns) for s in range(n)]
    for s in range(n):
        m = np.abs(ys[s].view(np.complex64))
        assert abs(m.max() - 1.0) < 1e-6                               # peak normalised
        if s:
            assert not np.array_equal(ys[s], ys[0])
    # noise level: the unit-amplitude signal sits at 1/peak, noise variance Ts/EbN0 relative to it
    x = ys[0].view(np.complex64).astype(np.complex128)
    pw = np.mean(np.abs(x) ** 2)
    g = 1.0 / (pw / (1.0 + 8.0 / 10.0)) ** 0.5                         # = peak, from E|y|^2 = (1 + nvar) / peak^2
    assert 2.0 < g < 6.0
    e.process()
    e.sync()
    for s in range(n):
        assert e.drain_packets(s) == pl[s].tobytes(), s
        sd_o, _, _, res_o = _run_oracle_stream(oracle_port, ys[s], "cf32", cfg["Fs"], cfg["Rs"], 2, "v1")
        assert res_o["packets"] == pl[s].tobytes()
    e.close()
    e = eng_mod.Engine(n, in_fmt="cf32", framing="v1", chunk_samples=200000)
    e.tx_synthesize(pl, f1, fs, ebno_db=4.0, seed=6)
    e.process()
    e.sync()
    assert sum(len(e.drain_packets(s)) for s in range(n)) < n * npk * 256 // 4
    e.close()
    for fmt in ("cs16", "cu8"):
        e = eng_mod.Engine(n, in_fmt=fmt, framing="v1", chunk_samples=200000)
        e.tx_synthesize(pl, f1, fs, ebno_db=12.0, seed=7)
        e.process()
        e.sync()
        for s in range(n):
            assert e.drain_packets(s) == pl[s].tobytes(), (fmt, s)
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,P,fmt", [(siggen.V1, 4, "cf32"), (siggen.V1, 2, "cu8"), (siggen.V2, 5, "cs16")])
def test_general_P_vs_oracle(eng_mod, oracle_port, cfg, P, fmt):
    """fsk_demod -p P with P < Ts (reference src/fsk_demod.c:186-188, src/fsk.c:139): the kernel's general-geometry
    instantiation (BLK = false) -- soft decisions, nin sequence, timing state and packets against the oracle"""
    framing = "v1" if cfg is siggen.V1 else "v2"
    raw, payloads = siggen.make_stream(31, n_packets=2, ebno_db=9.0, framing=framing, fmt=fmt, clock_ppm=-1500.0)
    sd_o, log_o, cons_o, res_o = _run_oracle_stream(oracle_port, raw, fmt, cfg["Fs"], cfg["Rs"], 2, framing, P=P)
    nsamp = raw.size // eng_mod.FMT_ELEMS[fmt]
    e = eng_mod.Engine(1, Fs=cfg["Fs"], Rs=cfg["Rs"], P=P, in_fmt=fmt, framing=framing, chunk_samples=nsamp + 1024)
    e.enable_frame_log(len(log_o) + 4)
    e.feed([raw])
    e.process()
    e.sync()
    sd_g = e.drain_soft(0)
    log_g = e.read_frame_log(0, len(log_o))
    assert np.array_equal(log_g[:, 0], log_o[:, 0]), "nin sequence"
    assert np.array_equal(log_g[:, 5].view(np.uint32), log_o[:, 5].view(np.uint32)), "norm_rx_timing"
    assert np.array_equal(log_g[:, 6].view(np.uint32), log_o[:, 6].view(np.uint32)), "ppm"
    assert np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32))
    assert e.drain_packets(0) == res_o["packets"]
    e.close()


def test_c_example_decodes_the_golden_stream(tmp_path):
    """examples/decode_file.c -- a plain C99 caller, no Python between it and the C ABI -- compiled, RUN on the GPU
    on the reference's own input (tests/golden/fsk_v1.npz: cu8 IQ) and compared with the packets the reference
    binaries wrote for it"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = str(tmp_path / "decode_file")
    libdir = os.path.join(root, "wenet_b200")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "decode_file.c"), "-L" + libdir, "-lwenet_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    z = np.load(os.path.join(GOLD, "fsk_v1.npz"))
    src, dst = tmp_path / "in.cu8", tmp_path / "out.bin"
    src.write_bytes(z["raw"].tobytes())
    r = subprocess.run([exe, str(src), str(dst)], stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    assert dst.read_bytes() == z["packets"].tobytes()
    assert b"packets: 2" in r.stderr
    # and through a pipe, the way start_rx.sh runs the reference (stdin -> stdout)
    r = subprocess.run([exe, "-", "-"], input=z["raw"].tobytes(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0 and r.stdout == z["packets"].tobytes()


def test_drain_all_packets_seq_is_the_codeword_number(eng_mod, oracle_port):
    """(stream, seq) of wb_drain_all_packets is unique for the life of the engine: seq = wb_codeword.seq of the codeword
    the payload came from, counting on across calls, with a gap where a codeword failed its CRC"""
    cfg = siggen.V1
    raws = [siggen.make_stream(300 + s, n_packets=5, ebno_db=eb, fmt="cf32")[0] for s, eb in enumerate((12.0, 6.5, 12.0))]
    n = max(r.size for r in raws) // 2
    e = eng_mod.Engine(3, in_fmt="cf32", framing="v1", chunk_samples=n // 2 + 2048)
    seen, cws = [], []
    for half in range(2):
        e.feed([r[:2 * (n // 2)] if half == 0 else r[2 * (n // 2):] for r in raws])
        e.process()
        e.sync()
        cws.append(e.drain_codewords())
        seen.append(e.drain_all_packets())
    pk = np.concatenate(seen)
    cw = np.concatenate(cws)
    assert len(seen[0]) and len(seen[1])
    keys = list(zip(pk["stream"].tolist(), pk["seq"].tolist()))
    assert len(set(keys)) == len(keys)                                     # no duplicate (stream, seq) across drains
    good = cw[cw["crc_ok"] == 1]
    assert sorted(keys) == sorted(zip(good["stream"].tolist(), good["seq"].tolist()))
    for part in seen:                                                      # each call: sorted by (stream, seq)
        k = list(zip(part["stream"].tolist(), part["seq"].tolist()))
        assert k == sorted(k)
    for s in range(3):
        want = oracle_port.deframer("v1", 10).feed(oracle_port.fsk(cfg["Fs"], cfg["Rs"]).run(raws[s], "cf32")[0])
        assert pk["payload"][pk["stream"] == s].tobytes() == want["packets"], s
        assert cw["seq"][cw["stream"] == s].tolist() == list(range(len(want["iters"]))), s
    e.close()


def test_process_soft_between_feed_and_process_keeps_the_fed_samples(eng_mod, oracle_port):
    """the ABI allows wb_feed, wb_process_soft, wb_feed, wb_process on one engine: the soft-only pass consumes no IQ
    samples, so what was fed before it must still be demodulated afterwards"""
    cfg = siggen.V1
    raw, _ = siggen.make_stream(77, n_packets=3, ebno_db=11.0, fmt="cf32")
    n = raw.size // 2
    sd_o = oracle_port.fsk(cfg["Fs"], cfg["Rs"]).run(raw, "cf32")[0]
    e = eng_mod.Engine(1, in_fmt="cf32", framing="v1", chunk_samples=n + 2048)
    a = (n // 3) & ~1
    e.feed([raw[:2 * a]])
    e.process()
    e.sync()
    got = [e.drain_soft(0)]
    e.feed([raw[2 * a:2 * 2 * a]])
    e.process_soft([np.zeros(100, dtype=np.float32)])       # a deframer-only pass in between
    e.sync()
    e.feed([raw[2 * 2 * a:]])
    e.process()
    e.sync()
    got.append(e.drain_soft(0))
    g = np.concatenate(got)
    assert g.size == sd_o.size and np.array_equal(g.view(np.uint32), sd_o.view(np.uint32))
    e.close()


def _device_count():
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
        except OSError:
            return 1
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 1


def test_multi_engine_equals_one_engine(eng_mod, oracle_port):
    """MultiEngine (one engine + one feeder thread per slot; slots = every visible GPU, or two engines on GPU 0 on a
    one-GPU box; weighted placement) against a single engine and the oracle on the same global block: same packets under
    the same global stream numbers, same soft decisions"""
    from wenet_b200.multi import MultiEngine
    cfg = siggen.V1
    n = 7
    raws = [siggen.make_stream(600 + s, n_packets=2, ebno_db=7.0 + s, fmt="cu8", clock_ppm=float(300 * (s - 3)))[0] for s in range(n)]
    nsamp = min(r.size for r in raws) // 2
    ndev = _device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    from wenet_b200.multi import probe_copy_rates
    rates = probe_copy_rates(sorted(set(devices)), seconds=0.05, nbytes=8 << 20)
    assert len(rates) == len(set(devices)) and all(r > 1.0 for r in rates)       # GB/s, pinned host -> device
    me = MultiEngine(n, devices=devices, weights=[1.0 + 0.5 * (i % 2) for i in range(len(devices))], in_fmt="cu8", framing="v1",
                     chunk_samples=nsamp + 1024)
    pb = me.pinned_block(nsamp)
    for s in range(n):
        pb.array[s, :] = raws[s][:2 * nsamp]
    me.step(pb.array)
    me.sync()
    pk = me.drain_all_packets()
    one = eng_mod.Engine(n, in_fmt="cu8", framing="v1", chunk_samples=nsamp + 1024)
    one.feed_strided(pb.array)
    one.process()
    one.sync()
    pk1 = one.drain_all_packets()
    assert len(pk) and np.array_equal(pk["stream"], pk1["stream"]) and np.array_equal(pk["seq"], pk1["seq"])
    assert np.array_equal(pk["payload"], pk1["payload"])
    assert me.last_samples == one.last_samples and len({g.device for g in me.engines}) == min(ndev, len(devices))
    for s in range(n):
        sd_o = oracle_port.fsk(cfg["Fs"], cfg["Rs"]).run(raws[s][:2 * nsamp], "cu8")[0]
        assert np.array_equal(me.drain_soft(s).view(np.uint32), sd_o.view(np.uint32)), s
        want = oracle_port.deframer("v1", 10).feed(sd_o)["packets"]
        assert pk["payload"][pk["stream"] == s].tobytes() == want, s
    one.close()
    # the pipelined receiver form: two more blocks through stream_step / flush, packets arrive one call later and carry on
    # the per-stream codeword numbering
    before = {s_: int(pk["seq"][pk["stream"] == s_].max()) for s_ in set(pk["stream"].tolist())}
    assert len(me.stream_step(pb.array)) == 0
    a = me.stream_step(pb.array)
    b = me.flush()
    assert len(a) and len(b) and len(me.flush()) == 0
    for part in (a, b):
        assert set(part["stream"].tolist()) <= set(range(n))
    for s_, last in before.items():
        assert a["seq"][a["stream"] == s_].min() > last and b["seq"][b["stream"] == s_].min() > a["seq"][a["stream"] == s_].max()
    me.close()
