"""argv surface of the CLI stand-ins (no GPU needed: only usage / exit-code behaviour, reference
src/fsk_demod.c:156-178, src/drs232_ldpc.c:142-146); the byte-stream parity of the shims is in the GPU tests."""
import os
import subprocess
import sys

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
