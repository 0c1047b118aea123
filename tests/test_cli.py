"""argv surface of the CLI stand-ins (no GPU needed: only usage / exit-code behaviour, reference
src/fsk_demod.c:156-178, src/drs232_ldpc.c:142-146); the byte-stream parity of the shims is in the GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(mod, *args):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-m", "wenet_b200.cli." + mod] + list(args), env=env, input=b"",
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)


def test_fsk_demod_usage_errors():
    r = run("fsk_demod", "2", "921416")
    assert r.returncode == 1 and b"Too few arguments" in r.stderr and b"usage:" in r.stderr
    r = run("fsk_demod", "2", "921416", "115177", "-", "-", "extra")
    assert r.returncode == 1 and b"Too many arguments" in r.stderr
    r = run("fsk_demod", "3", "921416", "115177", "-", "-")
    assert r.returncode == 1 and b"Mode 3 is not valid" in r.stderr
    r = run("fsk_demod", "-l", "2", "921416", "115177", "-", "-")
    assert r.returncode == 1 and b"low-rate" in r.stderr


def test_fsk_demod_parses_the_reference_command_lines():
    """-t[r] / --stats[=r] has an OPTIONAL argument in the reference (getopt_long "t::", src/fsk_demod.c:92-131):
    the command lines of start_rx.sh:126,136, start_rx_headless.sh:131 and rx/start_rx_docker.sh:80 must parse"""
    from wenet_b200.cli import fsk_demod as F
    o = F.parse(["fsk_demod", "--cu8", "-s", "--stats=100", "2", "921416", "115177", "-", "-"])
    assert (o["fmt"], o["soft"], o["stats"], o["stats_rate"], o["M"], o["Fs"], o["Rs"]) == ("cu8", True, True, 100, 2, 921416, 115177)
    o = F.parse(["fsk_demod", "-s", "--stats=100", "-b", "1", "-u", "23500", "2", "48000", "4800", "-", "-"])
    assert (o["fmt"], o["stats_rate"], o["lo"], o["hi"], o["Fs"], o["Rs"]) == ("s16", 100, 1, 23500, 48000, 4800)
    o = F.parse(["fsk_demod", "--cs16", "-s", "--stats=100", "2", "960000", "96000", "-", "-"])
    assert (o["fmt"], o["stats"], o["Fs"]) == ("cs16", True, 960000)
    # a bare -t / --stats never swallows the next word; an attached value counts; atoi() == 0 falls back to 8
    for argv, rate in ((["-t", "2"], 8), (["--stats", "2"], 8), (["-t25", "2"], 25), (["-st5", "2"], 5), (["-t0", "2"], 8),
                       (["-tx", "2"], 8)):
        o = F.parse(["fsk_demod"] + argv + ["921416", "115177", "-", "-"])
        assert o["stats"] and o["stats_rate"] == rate and o["M"] == 2, argv
    o = F.parse(["fsk_demod", "-p", "4", "-f", "4", "921416", "115177", "in", "out"])
    assert (o["P"], o["testframes"], o["stats"], o["M"], o["fin"], o["fout"]) == (4, True, False, 4, "in", "out")


def test_ldpc_usage_errors():
    for mod in ("drs232_ldpc", "wenet_ldpc"):
        r = run(mod, "-")
        assert r.returncode == 1 and b"usage:" in r.stderr
        r = run(mod, "/nonexistent/in", "-")
        assert r.returncode == 1 and b"Error opening input file" in r.stderr


def test_per_formatting_matches_the_reference_printf():
    """(float)packet_errors/packets through %4.3f (src/drs232_ldpc.c:261-265, :280); with no packets at all the
    reference binary prints -nan (checked against oracle/_ref/drs232_ldpc on empty input)"""
    from wenet_b200.cli._ldpc_cli import _per
    assert (_per(0, 0), _per(0, 2), _per(1, 3), _per(2, 3), _per(5, 5)) == ("-nan", "0.000", "0.333", "0.667", "1.000")


class _FakeStats:
    EbNodB, ppm, rx_timing = 9.5, 12.0, 0.1
    f_est = [140372.0, 273545.4, 0.0, 0.0]
    neyetr, neyesamp, nfft = 2, 3, 4
    rx_eye = [[0.5] * 160 for _ in range(8)]
    samp_fft = [0.25] * 512


class _FakeEngine:
    """stands in for wenet_b200.engine.Engine in the host-logic test below: whole frames of N samples in, Nbits soft
    decisions / hard bits per frame out (no GPU in the CPU suite; the real thing is in tests/test_gpu_parity.py)"""
    N, Nbits = 384, 48
    made = []

    def __init__(self, n, **kw):
        self.kw, self.pending, self.rest, self.frames, self.blocks = kw, 0, 0, 0, []
        _FakeEngine.made.append(self)

    def feed(self, arrs):
        self.rest += arrs[0].size // 2
        self.blocks.append(arrs[0].size // 2)

    def process(self):
        self.pending, self.rest = self.rest // self.N, self.rest % self.N
        self.frames += self.pending

    def sync(self):
        pass

    def drain_soft(self, s):
        import numpy as np
        return np.full(self.pending * self.Nbits, -1.0, dtype=np.float32)

    def drain_hard(self, s):
        import numpy as np
        assert self.kw["hard_bits"]
        return np.ones(self.pending * self.Nbits, dtype=np.uint8)

    def stats(self, s):
        assert self.kw["stats"]
        return _FakeStats()

    def close(self):
        pass


def _run_main(monkeypatch, argv, nbytes):
    import io
    import sys as _sys
    from wenet_b200 import engine as E
    from wenet_b200.cli import fsk_demod as F
    monkeypatch.setattr(E, "Engine", _FakeEngine)
    _FakeEngine.made.clear()
    out, err = io.BytesIO(), io.StringIO()
    fake_in = type("I", (), {"buffer": io.BytesIO(bytes(nbytes))})()
    fake_out = type("O", (), {"buffer": out})()
    monkeypatch.setattr(_sys, "stdin", fake_in)
    monkeypatch.setattr(_sys, "stdout", fake_out)
    monkeypatch.setattr(_sys, "stderr", err)
    assert F.main(["fsk_demod"] + argv) == 0
    return out.getvalue(), err.getvalue(), _FakeEngine.made[-1]


def test_fsk_demod_main_loop_host_logic(monkeypatch):
    """the stand-in's block loop around a fake engine: output sizes, the reference's stats cadence (--stats=100 at
    921416 / 384: one JSON line per 24 frames, src/fsk_demod.c:247-251), hard / soft / testframe branches"""
    import json
    nframes = 240
    out, err, eng = _run_main(monkeypatch, ["--cu8", "-s", "--stats=100", "2", "921416", "115177", "-", "-"], nframes * 384 * 2 + 100)
    assert len(out) == nframes * 48 * 4 and eng.kw["stats"] and not eng.kw["hard_bits"] and eng.kw["in_fmt"] == "cu8"
    recs = [json.loads(l) for l in err.splitlines()]
    assert len(recs) == nframes // 24 and set(eng.blocks[:-1]) == {24 * 384}
    for k in ("secs", "EbNodB", "ppm", "f1_est", "f2_est", "eye_diagram", "samp_fft"):     # rx/fskstatsudp.py:29, :214-226
        assert k in recs[0], k
    assert len(recs[0]["samp_fft"]) == 4 and len(recs[0]["eye_diagram"]) == 2 and "f3_est" not in recs[0]
    # hard-bit output, no stats: 64-frame blocks, one byte per bit from the engine's hard-bit tap
    out, err, eng = _run_main(monkeypatch, ["--cs16", "4", "921416", "115177", "-", "-"], nframes * 384 * 4)
    assert out == b"\x01" * (nframes * 48) and err == "" and eng.kw["hard_bits"] and eng.kw["M"] == 4
    assert eng.blocks[0] == 64 * 384
    # testframe mode: all-ones bits never match the known frame -> no lines, but the branch runs (soft and hard, with -t)
    for extra in (["-f"], ["-f", "-s"], ["-f", "-t"], ["-f", "-s", "-t5"]):
        out, err, eng = _run_main(monkeypatch, ["--cs16"] + extra + ["2", "921416", "115177", "-", "-"], 100 * 384 * 4)
        assert err == "" and len(out) == 100 * 48 * (4 if "-s" in extra else 1), extra


def test_ldpc_cli_main_loop_host_logic(monkeypatch, tmp_path):
    """drs232_ldpc stand-in around a fake engine: packet bytes out, the reference's per-packet and summary lines
    (src/drs232_ldpc.c:261-265, :280) with uint16 counters and printf's PER"""
    import numpy as np
    from wenet_b200 import engine as E
    from wenet_b200.cli import _ldpc_cli as L

    class Fake:
        sd_cap = 1 << 20

        def __init__(self, n, **kw):
            self.calls = 0

        def process_soft(self, arrs):
            self.calls += 1

        def sync(self):
            pass

        def drain_codewords(self):
            if self.calls != 1:
                return []
            bad = np.zeros(258, dtype=np.uint8)
            bad[256], bad[257] = 0x34, 0x12
            return [{"crc_ok": 1, "iters": 2}, {"crc_ok": 0, "iters": 10, "bytes": bad}, {"crc_ok": 1, "iters": 3}]

        def drain_packets(self, s):
            return b"\x55" * 512 if self.calls == 1 else b""

        def close(self):
            pass

    monkeypatch.setattr(E, "Engine", Fake)
    fin, fout = tmp_path / "in.f32", tmp_path / "out.bin"
    np.zeros(5000, dtype=np.float32).tofile(fin)
    import io
    import sys as _sys
    err = io.StringIO()
    monkeypatch.setattr(_sys, "stderr", err)
    assert L.main(["drs232_ldpc", str(fin), str(fout), "-v"], "v1", "drs232") == 0
    assert fout.read_bytes() == b"\x55" * 512
    lines = ["packets: 1 packet_errors: 0 PER: 0.000 iter: 2\n", "packets: 2 packet_errors: 1 PER: 0.500 iter: 10\n",
             "packets: 3 packet_errors: 1 PER: 0.333 iter: 3\n", "packets: 3 packet_errors: 1 PER: 0.333\n"]
    assert err.getvalue() == "".join(lines)
    # -vv (what start_rx.sh passes) adds the two checksums of a packet that fails its CRC, src/drs232_ldpc.c:246-251
    err.seek(0); err.truncate()
    from wenet_b200.siggen import crc16_ccitt_false
    assert L.main(["drs232_ldpc", str(fin), str(fout), "-vv"], "v1", "drs232") == 0
    lines.insert(1, "tx_checksum: 0x1234 rx_checksum: 0x%02x\n" % crc16_ccitt_false(bytes(256)))
    assert err.getvalue() == "".join(lines)


def test_fsk_demod_numbers_parse_like_atoi():
    """Mode / SampleRate / SymbolRate go through atoi() in the reference (src/fsk_demod.c:181-183), and a bare
    --fsk_lower / --fsk_upper (optional_argument without a value, :97-98) is accepted and ignored"""
    from wenet_b200.cli import fsk_demod as F
    o = F.parse(["fsk_demod", "--cu8", "-s", "--fsk_lower", "--fsk_upper", "2", "921416Hz", " 115177", "-", "-"])
    assert (o["M"], o["Fs"], o["Rs"], o["lo"], o["hi"]) == (2, 921416, 115177, 0, 0)
    with pytest.raises(SystemExit) as ex:
        F.parse(["fsk_demod", "two", "921416", "115177", "-", "-"])      # atoi("two") = 0: not a valid mode, exit(1)
    assert ex.value.code == 1


def test_stats_line_feeds_the_reference_relay_parser():
    """the stand-in's stderr JSON line goes through the UNMODIFIED rx/fskstatsudp.py FSKDemodStats.update
    (the parser between fsk_demod's stderr and wenetserver.py's UDP port) and comes out as snr / ppm / fest / fft_db"""
    import importlib.util
    import os
    import sys
    import types
    ref = "/root/reference/rx/fskstatsudp.py"
    if not os.path.exists(ref):
        pytest.skip("needs the reference tree")
    from wenet_b200 import engine as E
    from wenet_b200.cli import fsk_demod as F
    st = E.WbStats()
    st.EbNodB, st.ppm = 11.26, -37.9
    st.f_est[0], st.f_est[1] = 140372.3, 259148.2
    st.neyetr, st.neyesamp, st.nfft = 8, 16, 128
    for i in range(8):
        for j in range(16):
            st.rx_eye[i][j] = float("nan") if (i, j) == (0, 0) else (i + j) / 23.0     # the reference prints nan there too
    for k in range(128):
        st.samp_fft[k] = 0.001 * k
    line = F.stats_line(st, 2)
    sys.modules.setdefault("WenetPackets", types.SimpleNamespace(WENET_IMAGE_UDP_PORT=7890))   # its only import from rx/
    spec = importlib.util.spec_from_file_location("ref_fskstatsudp", ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p = mod.FSKDemodStats(averaging_time=5.0, peak_hold=True, freq=441200000, sample_rate=921416)
    errors = []
    p.log_error = errors.append
    p.update(line)
    assert not errors, errors
    assert abs(p.snr - 11.3) < 1e-9 and p.ppm == -37 and p.fest == [140372.3, 259148.2]
    assert len(p.fft_db) == 128 and len(p.fft_freq) == 128 and abs(p.fcentre - (441200000 + (140372.3 + 259148.2) / 2)) < 1e-6
    # 4-FSK adds f3_est / f4_est, which the parser ignores
    st.f_est[2], st.f_est[3] = 300000.0, 400000.0
    p.update(F.stats_line(st, 4))
    assert not errors
