"""bench.py's reference arm on the host cores (the arm the driver runs beside ours): it must print ONE JSON line with the
same metric / unit / config as the GPU arm, `impl: reference`, a `cpu_baseline` that describes this very run and an `e2e`
that repeats the line's value with no copies.  Runs the reference binaries of oracle/_ref when they are built here (else
the oracle port), on a bounded sample: a few seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference"
    assert d["metric"] == "IQ Msamples/s demodulated+decoded" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None                          # BASELINE.md holds no published number for this metric
    cfg = d["config"]
    assert "configs[3]" in cfg["workload"] and cfg["streams_per_gpu"] == 4096 and cfg["chunk_samples"] == 1 << 20
    assert "model" not in cfg
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["unit"] == d["unit"]
    assert cb["value"] == d["value"]
    assert (cb["kind"] == "reference") == os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fsk_demod"))
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    """under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without output"""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    """without a CUDA device the GPU arm must fail loudly and print no number (skipped on a GPU box)"""
    import pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
