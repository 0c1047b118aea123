#!/usr/bin/env python3
"""bench.py -- IQ Msamples/s demodulated AND decoded (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W                  # this engine
    python bench.py --impl reference --gpus N --steps K --warmup W # the reference CPU pipe on the host cores

Workload (config.workload): BASELINE.json configs[3], "end-to-end demod+deframe+LDPC: 4096 streams, Eb/N0
sweep 4-12 dB", time-chunked: one STEP = one pass of the whole path (K1 fsk -> K2 deframe -> K3 llr stats
-> K4 ldpc+crc) over one HBM-resident chunk of `--chunk` samples of every one of the 4096 streams of this
GPU (v1 RS232 framing, 921416 sps / 115177 baud 2-FSK, cf32).  configs[3] as written (8 Msample per stream)
is 256 GB of cf32 and does not fit one GPU, so it runs as consecutive chunks with the per-stream state carried
in HBM; every step demodulates a full chunk.  Multi-GPU: weak scaling, 4096 streams per GPU, no collective.

`value`  : whole-job Msamples/s with the chunk already resident in HBM (CUDA events on the engine's stream).
`e2e`    : the same metric through the public API with HOST buffers: wb_feed_strided (pinned host -> HBM),
           wb_process, wb_sync, wb_drain_all_packets (HBM -> host) every step, wall clock around device syncs.
           The host buffers hold the streams as cu8, 2 bytes per IQ sample: what rtl_sdr delivers to the
           reference receiver (start_rx.sh:125, `fsk_demod --cu8`), what the reference's own benchmark pipes into it
           (benchmarking/README.md: `csdr convert_f_u8 | fsk_demod --cu8`) and what this bench's reference arm reads
           -- the reference cannot read float IQ at all (src/fsk_demod.c:92-106) -- so both arms of the headline
           ratio consume the same bytes.  The same run with cs16 (4 B/sample) and cf32 (8 B/sample) host buffers is
           reported beside it (e2e_cs16, e2e_cf32): all three are bound by the PCIe copy.
`roofline`: the dominant kernel (wb_fsk_kernel): algorithmic bytes (8.5 B per IQ sample) / its
           event-timed duration, against the measured HBM peak in MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own binaries (oracle/_ref: fsk_demod | drs232_ldpc, built from the unmodified
           sources) on the host cores, bounded sample.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EBNO_SWEEP = [4.0, 6.0, 8.0, 10.0, 12.0]
ALG_BYTES_PER_SAMPLE_FSK = 8.5          # cf32 in (8 B) + 48 floats out per 384 samples (0.5 B), SURVEY 8(d)
METRIC = "IQ Msamples/s demodulated+decoded"
UNIT = "Msamples/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------- synthetic input

def make_sources(n_src, nsamp, seed_base, mode="v1"):
    """n_src distinct streams (seeded), Eb/N0 cycling through the 4-12 dB sweep."""
    from wenet_b200 import siggen
    out = []
    for i in range(n_src):
        eb = EBNO_SWEEP[i % len(EBNO_SWEEP)]
        if mode == "fsk4":
            raw, _ = siggen.make_4fsk_stream(seed_base + i, nsamp // 8 + 16, ebno_db=eb + 3.0)
            raw = raw[:2 * nsamp]
        else:
            raw, _ = siggen.make_stream(seed_base + i, n_samples=nsamp, ebno_db=eb, framing=mode, fmt="cf32",
                                        clock_ppm=float((i % 7 - 3) * 400))
        out.append(raw)
    return out


# ---------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.proc = None
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, "-i", str(device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 9]
        if not rows:
            rows = [r for (t, r) in self.rows if len(r) >= 9][-5:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


# ---------------------------------------------------------------- the reference CPU pipe

def ref_binaries():
    d = os.path.join(ROOT, "oracle", "_ref")
    f, g = os.path.join(d, "fsk_demod"), os.path.join(d, "drs232_ldpc")
    return (f, g) if os.path.exists(f) and os.path.exists(g) else None


def run_cpu_pipes(paths, n_pipes):
    """n_pipes x (fsk_demod --cu8 -s 2 921416 115177 file - | drs232_ldpc - -), all at once, pipe i on paths[i % len].
    -> (wall s, bytes out)"""
    f, g = ref_binaries()
    cmd = "%s --cu8 -s 2 921416 115177 %s - 2>/dev/null | %s - - 2>/dev/null | wc -c"
    t0 = time.perf_counter()
    procs = [subprocess.Popen(["bash", "-c", cmd % (f, paths[i % len(paths)], g)], stdout=subprocess.PIPE, text=True)
             for i in range(n_pipes)]
    outs = [p.communicate()[0] for p in procs]
    dt = time.perf_counter() - t0
    return dt, sum(int(o.strip() or 0) for o in outs)


def run_cpu_port(raw_cu8, n_threads):
    """fallback when oracle/_ref is absent: the C restatement (oracle/liboracle.so), one stream per thread"""
    from oracle import oracle as O
    port = O.Oracle("port")
    nbytes = [0] * n_threads

    def work(i):
        sd, _, _ = port.fsk(921416, 115177, M=2).run(raw_cu8, "cu8")
        nbytes[i] = len(port.deframer("v1", 10).feed(sd)["packets"])

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    return time.perf_counter() - t0, sum(nbytes)


def cpu_sample_file(nsamp, tmpdir):
    """cu8 file: five v1 streams, one per Eb/N0 of the GPU workload's 4-12 dB sweep (same generator, same clock offsets),
    1 Mi samples each, back to back, tiled to about nsamp samples: every pipe does the same, balanced work"""
    from wenet_b200 import siggen
    seg = [siggen.make_stream(i, n_samples=1 << 20, ebno_db=eb, fmt="cu8", clock_ppm=float((i % 7 - 3) * 400))[0]
           for i, eb in enumerate(EBNO_SWEEP)]
    base = np.concatenate(seg)
    per = len(seg) << 20
    reps = max(1, nsamp // per)
    path = os.path.join(tmpdir, "wb_cpu_sample.cu8")
    with open(path, "wb") as fh:
        for _ in range(reps):
            fh.write(base.tobytes())
    return [path], reps * per, base


def cpu_baseline(nsamp_per_pipe, reps=1):
    cores = os.cpu_count() or 1
    tmpdir = tempfile.mkdtemp(prefix="wb_bench_")
    try:
        path, ns, base = cpu_sample_file(nsamp_per_pipe, tmpdir)
        if ref_binaries():
            kind, pipes = "reference", max(1, cores // 2)
            best = None
            for _ in range(reps):
                dt, nb = run_cpu_pipes(path, pipes)
                best = dt if best is None else min(best, dt)
            used = min(cores, 2 * pipes)
        else:
            kind, pipes = "port", max(1, cores)
            raw = np.tile(base, ns // (base.size // 2))
            best, nb = run_cpu_port(raw, pipes)
            used = pipes
        return {"value": round(pipes * ns / best / 1e6, 3), "unit": UNIT, "cores": used, "kind": kind,
                "sample": "%d parallel pipes (fsk_demod --cu8 -s 2 921416 115177 | drs232_ldpc) x %d samples each (v1 streams at the "
                          "workload's Eb/N0 sweep 4-12 dB back to back), %.2f s wall, %d B decoded" % (pipes, ns, best, nb)}
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tmpdir = tempfile.mkdtemp(prefix="wb_bench_")
    try:
        path, ns, base = cpu_sample_file(args.cpu_samples, tmpdir)
        have = ref_binaries() is not None
        pipes = max(1, cores // 2) if have else max(1, cores)
        raw = None if have else np.tile(base, ns // (base.size // 2))
        times, nb = [], 0
        for i in range(args.warmup + args.steps):
            dt, nb = run_cpu_pipes(path, pipes) if have else run_cpu_port(raw, pipes)
            if i >= args.warmup:
                times.append(dt)
        total = sum(times)
        value = pipes * ns * len(times) / total / 1e6
        line = {
            "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": min(cores, 2 * pipes) if have else pipes,
                             "kind": "reference" if have else "port",
                             "sample": "each step: %d parallel pipes x %d samples (v1 streams at the Eb/N0 sweep 4-12 dB back to back, cu8), %d B decoded per step"
                                       % (pipes, ns, nb)},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


# ---------------------------------------------------------------- this engine

def workload_config(args, world):
    if args.mode != "v1":
        return {"workload": "exploration mode %s: %d streams/GPU x %d-sample chunks" % (args.mode, args.streams, args.chunk),
                "streams_per_gpu": args.streams, "chunk_samples": args.chunk, "in_fmt": "cf32", "mode": args.mode}
    return {"workload": "BASELINE.json configs[3] time-chunked: end-to-end 2-FSK demod + v1 deframe + LDPC(2580,2064) "
                        "max_iter 10 + CRC, %d streams/GPU x %d-sample chunks, Eb/N0 sweep 4-12 dB, Fs 921416 Rs 115177"
                        % (args.streams, args.chunk),
            "streams_per_gpu": args.streams, "chunk_samples": args.chunk, "in_fmt": "cf32", "framing": "v1",
            "ldpc_max_iter": 10, "parallelism": "stream-sharded x%d, no collective" % world,
            "l2": "inputs (%.1f GB per step per GPU) larger than L2" % (args.streams * args.chunk * 8 / 1e9),
            "e2e_in_fmt": "cu8 (rtl_sdr's format = the reference arm's input bytes)", "e2e_host_bytes_per_step": args.e2e_bytes,
            "synth": args.synth}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--chunk", type=int, default=1 << 20, help="samples per stream per step (HBM-resident)")
    ap.add_argument("--e2e-bytes", type=int, default=1 << 20, help="host bytes per stream per e2e step (pinned host)")
    ap.add_argument("--sources", type=int, default=40, help="distinct synthetic streams generated on the host")
    ap.add_argument("--cpu-samples", type=int, default=32 << 20, help="samples per CPU pipe (reference arm)")
    ap.add_argument("--mode", default="v1", choices=["v1", "v2", "fsk4", "fskonly", "ldpc"],
                    help="v1 = the headline workload; exploration modes (no e2e / cpu_baseline, not the graded line): v2 "
                         "(960000/96000, wenet_ldpc framing), fsk4 (4-FSK demod only, BASELINE configs[4]), fskonly (2-FSK demod "
                         "only, BASELINE configs[1]: use --streams 1024), ldpc (BASELINE configs[2]: --codewords H2064_516 "
                         "codewords of LLRs resident in HBM, --ldpc-iter max iterations; the metric is Mcodewords/s)")
    ap.add_argument("--codewords", type=int, default=1 << 20)
    ap.add_argument("--ldpc-iter", type=int, default=100)
    ap.add_argument("--synth", default="host", choices=["host", "device"],
                    help="host = 40 seeded numpy streams (with transmitter clock offsets) replicated on the device (default); "
                         "device = every stream distinct, built in HBM by wb_tx_synthesize (frame_packet + fsk_mod_c + AWGN)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return 0

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from wenet_b200 import engine as E          # raises if libwenet_b200.so is missing: no CPU fallback

    if args.mode == "ldpc":
        # BASELINE configs[2]: LDPC only.  24 seeded noisy codewords (LLRs through the engine's own sd_to_llr) replicated
        # to --codewords in HBM; one step = one decode of all of them.
        from wenet_b200 import siggen
        rng = np.random.default_rng(77 + rank)
        sd = []
        for k in range(24):
            data = rng.integers(0, 2, 2064).astype(np.uint8)
            cw = np.concatenate([data, siggen.ldpc_parity_bits(data)]).astype(np.float64)
            sd.append((1 - 2 * cw) * rng.uniform(0.5, 2.0) + 10 ** (-(1.5 + 0.25 * k) / 20) * rng.standard_normal(2580))
        eng = E.Engine(1, framing="v1", chunk_samples=4096, device=local)
        llr = eng.sd_to_llr_batch(np.stack(sd).astype(np.float32))
        eng.dev_ldpc_setup(llr, args.codewords)
        for _ in range(max(args.warmup, 1)):
            eng.dev_ldpc_run(args.ldpc_iter)
        eng.sync()
        barrier()
        eng.timer_start()
        for _ in range(args.steps):
            eng.dev_ldpc_run(args.ldpc_iter)
        ms = max_over_ranks(eng.timer_stop())
        _, iters, _ = eng.dev_ldpc_result(0, 24)
        if rank == 0:
            total = sum_over_ranks(float(args.codewords)) * args.steps
            print(json.dumps({"metric": "LDPC Mcodewords/s (exploration mode, BASELINE configs[2])", "value": round(total / (ms * 1e-3) / 1e6, 3),
                              "unit": "Mcodewords/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "dtype": "f32",
                              "data": "synthetic", "config": {"workload": "H2064_516 (2580, 2064), %d codewords/GPU resident, max_iter %d, "
                                                              "iterations of the 24 sources %s" % (args.codewords, args.ldpc_iter, iters.tolist())},
                              "roofline": {"bound": "hbm", "achieved": round(10582.0 * total / (ms * 1e-3) / 1e9, 2), "unit": "GB/s",
                                           "note": "10 582 algorithmic bytes per codeword; the kernel is shared-memory / issue bound"}}),
                  flush=True)
        else:
            sum_over_ranks(float(args.codewords))
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    n, chunk = args.streams, args.chunk
    n_src = min(args.sources, n)
    sources = make_sources(n_src, chunk, seed_base=rank * 1000, mode="v1" if args.mode == "fskonly" else args.mode)
    if args.mode != "v1":
        args.no_e2e = args.no_cpu_baseline = True
    ekw = {"v1": dict(framing="v1"), "v2": dict(Fs=960000, Rs=96000, framing="v2"), "fsk4": dict(M=4, framing="none"),
           "fskonly": dict(framing="none")}[args.mode]

    eng = E.Engine(n, in_fmt="cf32", chunk_samples=chunk, device=local, **ekw)
    fill = chunk
    if args.synth == "device" and args.mode in ("v1", "v2"):
        from wenet_b200 import siggen
        cfgs = siggen.V2 if args.mode == "v2" else siggen.V1
        ts = cfgs["Fs"] // cfgs["Rs"]
        frame_samples = (343 * (10 if args.mode == "v1" else 8)) * ts
        npk = max(1, (chunk - 2448 * ts) // frame_samples)
        rng = np.random.default_rng(4242 + rank)
        pl = rng.integers(0, 256, size=(n, npk, 256), dtype=np.uint8)
        pl[:, :, 0] = 0x55
        ebno = np.array([EBNO_SWEEP[s_ % len(EBNO_SWEEP)] for s_ in range(n)], dtype=np.float32)
        fill = eng.tx_synthesize(pl, int(cfgs["f_lo"]), int(cfgs["f_hi"] - cfgs["f_lo"]), ebno_db=ebno, seed=rank)
    else:
        eng.feed(sources + [None] * (n - n_src))
        eng.sync()
        eng.dev_replicate(n_src, chunk, 4096 + 16 * 37)     # stream s = source s % n_src rotated by (s // n_src) * 4688 samples
    eng.dev_set_fill(fill)

    def one_step():
        eng.dev_set_fill(fill)
        eng.process()

    attempts = 0
    while True:
        attempts += 1
        clocks = ClockSampler(local)
        for _ in range(args.warmup):
            one_step()
        eng.sync()
        eng.drain_all_packets()          # warm-up output is not part of the report
        barrier()
        l0 = eng.launch_count
        t0 = time.perf_counter()
        eng.timer_start()
        for _ in range(args.steps):
            one_step()
        ms = eng.timer_stop()
        t1 = time.perf_counter()
        l1 = eng.launch_count
        barrier()
        clk = clocks.summary(t0, t1)
        clocks.stop()
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clk.get("reasons", []))
        if not bad or attempts >= 2:
            break
    # per-kernel split and work done, from the last timed step (every step does the same work)
    kms = eng.last_kernel_ms().astype(np.float64)
    samples_step = eng.last_samples
    codewords_step = eng.last_codewords
    pk = eng.drain_all_packets()
    packets_last = int(len(pk))
    ms = max_over_ranks(ms)
    total_samples = sum_over_ranks(float(samples_step)) * args.steps
    value = total_samples / (ms * 1e-3) / 1e6

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bps = 9.0 if args.mode == "fsk4" else ALG_BYTES_PER_SAMPLE_FSK
    fsk_gbs = alg_bps * samples_step / (kms[0] * 1e-3) / 1e9 if kms[0] > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch: one `ncu --set full` capture of this same workload
    # (profiles/r01_traffic.json says how it was taken); null for any other workload
    traffic = issue_pct = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            tj = json.load(fh)
        w = tj["wb_fsk_kernel"]["workload"]
        if w["streams"] == n and w["chunk_samples"] == chunk and w["in_fmt"] == "cf32" and args.mode == "v1":
            traffic = tj["wb_fsk_kernel"]["dram_bytes_read"] + tj["wb_fsk_kernel"]["dram_bytes_write"]
            # what actually bounds the kernel (SURVEY 8d asks for it beside the HBM fraction): issue-slot utilisation
            # from the same ncu capture, smsp__issue_active.avg.pct_of_peak_sustained_active
            issue_pct = tj["wb_fsk_kernel"].get("issue_active_pct")
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "wb_fsk_kernel", "achieved": round(fsk_gbs, 2), "peak": peak, "unit": "GB/s",
                "frac": round(fsk_gbs / peak, 4), "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "kernel_ms": {"fsk": round(float(kms[0]), 3), "deframe": round(float(kms[1]), 3),
                              "llr_stats": round(float(kms[2]), 3), "ldpc": round(float(kms[3]), 3)},
                "alg_bytes_per_launch": alg_bps * samples_step,
                "issue_active_pct": issue_pct,
                "note": "the kernel is a chain of dependent fp32 phases per frame (bit-exact with the reference's operation "
                        "order): issue slots, not HBM, are the binding resource; see DESIGN.md section 4"}

    # ---- e2e through the public API with host buffers ----
    def run_e2e(fmt, n_eng):
        """feed (pinned host -> HBM) + process + sync + drain (HBM -> host) every step.  The streams are split over
        n_eng engines on the same GPU and each engine's drain of step k is issued right before its feed of step
        k + 1, so one engine's copies overlap the other's kernels across step boundaries too."""
        ec = min(chunk, args.e2e_bytes // E.FMT_BPS[fmt])
        per = [n // n_eng + (1 if i < n % n_eng else 0) for i in range(n_eng)]
        elems = E.FMT_ELEMS[fmt]
        engs, pins, base = [], [], 0
        for cnt in per:
            engs.append(E.Engine(cnt, in_fmt=fmt, framing="v1", chunk_samples=ec + 1024, device=local))
            pb = E.PinnedBuffer((cnt, elems * ec), E.FMT_DTYPE[fmt])
            for j in range(cnt):
                s_ = base + j
                src = sources[s_ % n_src]
                r = ((s_ // n_src) * 4688) % (chunk - ec) if chunk > ec else 0
                seg = src[2 * r:2 * r + 2 * ec]
                if fmt == "cf32":
                    pb.array[j, :] = seg
                elif fmt == "cs16":                     # siggen.to_format: the reference divides by 1000 (FDMDV_SCALE)
                    pb.array[j, :] = np.round(seg.astype(np.float64) * 1000.0).astype(np.int16)
                else:                                   # the same samples as rtl_sdr would deliver them (cu8)
                    pb.array[j, :] = np.clip(np.round(seg * 127.0 + 127.0), 0, 255).astype(np.uint8)
            pins.append(pb)
            base += cnt
        stat = {"d2h": 0, "samples": 0, "packets": 0}
        pending = [False] * n_eng

        def collect(i):
            g = engs[i]
            g.sync()
            out = g.drain_all_packets()
            stat["d2h"] += out.nbytes + g.n_streams * 40 + g.last_codewords * 280
            stat["samples"] += g.last_samples
            stat["packets"] += len(out)
            pending[i] = False

        def step():
            for i, (g, pb) in enumerate(zip(engs, pins)):
                if pending[i]:
                    collect(i)
                g.feed_strided(pb.array)
                g.process()
                pending[i] = True

        def flush():
            for i in range(n_eng):
                if pending[i]:
                    collect(i)

        for _ in range(max(args.warmup, 1)):
            step()
        flush()
        barrier()
        stat.update(d2h=0, samples=0, packets=0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        flush()                                     # every step's results are on the host when the clock stops
        dt = time.perf_counter() - t0
        barrier()
        dt = max_over_ranks(dt)
        consumed = sum_over_ranks(float(stat["samples"]))
        res = {"value": round(consumed / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(n * ec * E.FMT_BPS[fmt]),
               "d2h_bytes_per_step": int(stat["d2h"] // args.steps), "ms_per_step": round(1e3 * dt / args.steps, 3), "in_fmt": fmt,
               "chunk_samples": ec, "engines": n_eng, "crc_valid_packets_per_step": stat["packets"] // args.steps,
               "api": "wb_feed_strided(pinned host) + wb_process + wb_sync + wb_drain_all_packets"}
        for g in engs:
            g.close()
        del pins
        return res

    e2e = e2e_cf32 = e2e_cs16 = None
    if not args.no_e2e:
        eng.close()
        e2e = run_e2e("cu8", 2)
        e2e_cs16 = run_e2e("cs16", 2)
        e2e_cf32 = run_e2e("cf32", 2)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(32 << 20)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clk, "e2e": e2e, "e2e_cs16": e2e_cs16, "e2e_cf32": e2e_cf32, "gpu_launches": int(l1 - l0),
            "roofline": roofline, "cpu_baseline": cpu,
            "work_per_step_per_gpu": {"samples": int(samples_step), "codewords": int(codewords_step),
                                      "crc_valid_packets": packets_last},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
