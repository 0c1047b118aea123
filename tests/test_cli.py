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
    assert r.returncode == 1


def test_ldpc_usage_errors():
    for mod in ("drs232_ldpc", "wenet_ldpc"):
        r = run(mod, "-")
        assert r.returncode == 1 and b"usage:" in r.stderr
        r = run(mod, "/nonexistent/in", "-")
        assert r.returncode == 1 and b"Error opening input file" in r.stderr
