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
`parity`  : the gate every number passes first ("parity gates must pass for a number to count").  Outside
           the timed regions the CPU oracle demodulates + deframes + decodes >= 10 of the ACTUAL rows of this run --
           the rows of the 4096-stream HBM-resident launch (read back from the device: one per Eb/N0 plus rotated
           replicas, the full chunk) and the same rows of the e2e cu8 host buffers -- and soft decisions (bit for bit),
           packets and LDPC iteration counts must match; on a mismatch no `value` is printed and the exit code is 1.
`ebn0_ladder`: decoded bytes per Eb/N0 (the reference's own benchmark output, benchmarking/README.md:65-81): the five
           4/6/8/10/12 dB streams of the e2e cu8 buffers through this engine and through the reference's CPU pipe.
`extra`   : the other BASELINE.json configurations, each with its own roofline fraction, measured after the headline:
           configs[1] FSK-demod only 1024 streams x 1 Msample, configs[2] LDPC only 1 M codewords at max_iter 10 and
           100, configs[4] 4-FSK 1024 streams per GPU (8192 over 8 GPUs).
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
        single = None
        if ref_binaries():
            kind, pipes = "reference", max(1, cores // 2)
            best = None
            for _ in range(reps):
                dt, nb = run_cpu_pipes(path, pipes)
                best = dt if best is None else min(best, dt)
            used = min(cores, 2 * pipes)
            dt1, _ = run_cpu_pipes(path, 1)          # one stream through one pipe (2 processes)
            single = {"value": round(ns / dt1 / 1e6, 3), "unit": UNIT, "cores": 2, "sample": "one pipe x %d samples" % ns}
        else:
            kind, pipes = "port", max(1, cores)
            raw = np.tile(base, ns // (base.size // 2))
            best, nb = run_cpu_port(raw, pipes)
            used = pipes
        return {"value": round(pipes * ns / best / 1e6, 3), "unit": UNIT, "cores": used, "kind": kind,
                "sample": "%d parallel pipes (fsk_demod --cu8 -s 2 921416 115177 | drs232_ldpc) x %d samples each (v1 streams at the "
                          "workload's Eb/N0 sweep 4-12 dB back to back), %.2f s wall, %d B decoded" % (pipes, ns, best, nb),
                "single_stream": single}
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


# ---------------------------------------------------------------- parity gate + Eb/N0 ladder (test infrastructure)

def to_cu8(seg):
    """float IQ -> what rtl_sdr would deliver (siggen.to_format 'cu8')"""
    return np.clip(np.round(seg * 127.0 + 127.0), 0, 255).astype(np.uint8)


def check_rows(n, n_src):
    """rows of the batch the oracle re-does: one per Eb/N0 (the unrotated sources 0..4) plus rotated replicas spread over
    the batch, incl. the last row"""
    rows = list(range(min(len(EBNO_SWEEP), n)))
    for s in (n_src + 5, 2 * n_src + 17, n // 3 + 1, n // 2 + 3, (3 * n) // 4 + 2, n - 1):
        if 0 <= s < n and s not in rows:
            rows.append(s)
    return rows


def oracle_decode(port, raw, fmt):
    """the CPU restatement of the whole path on one row -> (soft decisions, deframer result dict)"""
    sd, _, _ = port.fsk(921416, 115177, M=2).run(raw, fmt)
    return sd, port.deframer("v1", 10).feed(sd)


def ladder_rows(args):
    """the five Eb/N0 streams (4..12 dB) as the e2e cu8 buffers hold them in rows 0..4 of rank 0: both arms decode these"""
    from wenet_b200 import siggen
    ec = min(args.chunk, args.e2e_bytes // 2)
    out = []
    for i, eb in enumerate(EBNO_SWEEP):
        raw, _ = siggen.make_stream(i, n_samples=args.chunk, ebno_db=eb, framing="v1", fmt="cf32", clock_ppm=float((i % 7 - 3) * 400))
        out.append((eb, to_cu8(raw[:2 * ec])))
    return out


def reference_ladder(rows):
    """decoded bytes per Eb/N0 from the reference's own binaries (or the port where they are absent)"""
    from oracle import oracle as O
    out, kind = [], "reference" if ref_binaries() else "port"
    port = None if ref_binaries() else O.Oracle("port")
    for eb, cu8 in rows:
        if port is None:
            nb = len(O.run_ref_pipe(cu8.tobytes(), fmt="cu8"))
        else:
            nb = len(oracle_decode(port, cu8, "cu8")[1]["packets"])
        out.append({"ebno_db": eb, "samples": int(cu8.size // 2), "bytes": int(nb)})
    return out, kind


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
        try:
            lad, kind = reference_ladder(ladder_rows(args))
            line["ebn0_ladder"] = {"arm": kind, "rows": lad}
        except Exception as ex:                      # the ladder is a report, not the measurement
            line["ebn0_ladder"] = {"error": str(ex)[:200]}
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


class Ranks:
    """the bench's only use of torch: barrier / reductions between the one-process-per-GPU ranks of a torchrun launch"""

    def __init__(self, world, local):
        self.world, self.dist = world, None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
            self.dist, self.torch = dist, torch

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.dist else None)

    def min(self, x):
        return self._red(x, self.dist.ReduceOp.MIN if self.dist else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.dist else None)

    def gather(self, x):
        """-> [x of rank 0, x of rank 1, ...] on every rank"""
        if self.dist is None:
            return [x]
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def ncu_static(kernel, n, chunk, mode):
    """per-launch counters of one `ncu --set full` capture of this same workload (profiles/r02_traffic.json says how it
    was taken); None for any other workload"""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                tj = json.load(fh)
            k = tj[kernel]
            w = k["workload"]
            if w["streams"] == n and w["chunk_samples"] == chunk and w["in_fmt"] == "cf32" and mode == "v1":
                return k
        except Exception:
            continue
    return None


def bench_ldpc(E, R, args, local, rank, n_cw, max_iter, steps):
    """BASELINE configs[2]: n_cw codewords of LLRs resident in HBM (24 seeded noisy codewords through the engine's own
    sd_to_llr, replicated), one step = one decode of all of them"""
    from wenet_b200 import siggen
    rng = np.random.default_rng(77 + rank)
    sd = []
    for k in range(24):
        data = rng.integers(0, 2, 2064).astype(np.uint8)
        cw = np.concatenate([data, siggen.ldpc_parity_bits(data)]).astype(np.float64)
        sd.append((1 - 2 * cw) * rng.uniform(0.5, 2.0) + 10 ** (-(1.5 + 0.25 * k) / 20) * rng.standard_normal(2580))
    eng = E.Engine(1, framing="v1", chunk_samples=4096, device=local)
    try:
        llr = eng.sd_to_llr_batch(np.stack(sd).astype(np.float32))
        eng.dev_ldpc_setup(llr, n_cw)
        out = {}
        for mi in max_iter:
            eng.dev_ldpc_run(mi)
            eng.sync()
            R.barrier()
            eng.timer_start()
            for _ in range(steps):
                eng.dev_ldpc_run(mi)
            ms = R.max(eng.timer_stop())
            _, iters, _ = eng.dev_ldpc_result(0, 24)
            total = R.sum(float(n_cw)) * steps
            out[mi] = (total, ms, iters.tolist())
        return out
    finally:
        eng.close()


def bench_fsk_only(E, R, local, sources, n, chunk, M, warmup, steps):
    """FSK demodulator alone (framing none) over n resident streams (sources replicated + rotated) -> (samples/step all
    ranks, ms for `steps` steps, kernel ms of the last step)"""
    eng = E.Engine(n, in_fmt="cf32", chunk_samples=chunk, device=local, M=M, framing="none")
    try:
        n_src = min(len(sources), n)
        eng.feed(list(sources[:n_src]) + [None] * (n - n_src))
        eng.sync()
        eng.dev_replicate(n_src, chunk, 4096 + 16 * 37)
        for _ in range(max(warmup, 1)):
            eng.dev_set_fill(chunk)
            eng.process()
        eng.sync()
        R.barrier()
        eng.timer_start()
        for _ in range(steps):
            eng.dev_set_fill(chunk)
            eng.process()
        ms = R.max(eng.timer_stop())
        kms = float(eng.last_kernel_ms()[0])
        return R.sum(float(eng.last_samples)), ms, kms, eng.streams_per_cta
    finally:
        eng.close()


def run_multiengine(args):
    """`python bench.py --gpus N` WITHOUT torchrun: the same workload through wenet_b200.MultiEngine -- one process, one
    engine pair and one feeder thread per GPU, no torch, no collective ("one host thread/process per GPU; outputs
    gathered on the host").  Prints the same JSON line; `value` = input resident in HBM (every GPU's 4096 cf32 streams),
    `e2e` = cu8 host buffers through MultiEngine.stream_step."""
    from wenet_b200 import engine as E
    from wenet_b200.multi import MultiEngine, probe_copy_rates
    G, n, chunk = args.gpus, args.streams, args.chunk
    total, n_src = n * G, min(args.sources, args.streams)
    sources = make_sources(n_src, chunk, seed_base=0, mode="v1")
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    port = None
    if not args.no_parity:
        from oracle import oracle as O
        O.build()
        port = O.Oracle("port")
    parity = {"oracle": "CPU restatement oracle/liboracle.so"} if port else {"skipped": True}
    ok_all = True

    # ---- value: every GPU's streams resident in HBM, one thread per GPU queues the steps ----
    res = MultiEngine(total, devices=range(G), in_fmt="cf32", framing="v1", chunk_samples=chunk)

    def setup(k, g):
        g.feed(sources[:min(n_src, g.n_streams)] + [None] * (g.n_streams - min(n_src, g.n_streams)))
        g.sync()
        g.dev_replicate(min(n_src, g.n_streams), chunk, 4096 + 16 * 37)
        g.dev_set_fill(chunk)
    res.each(setup)

    def steps(k, g, cnt, timed):
        if timed:
            g.timer_start()
        for _ in range(cnt):
            g.dev_set_fill(chunk)
            g.process()
        return g.timer_stop() if timed else g.sync()
    clocks = ClockSampler(0)
    res.each(lambda k, g: steps(k, g, max(args.warmup, 1), False))
    l0 = res.launch_count
    t0 = time.perf_counter()
    ms = max(res.each(lambda k, g: steps(k, g, args.steps, True)))
    t1 = time.perf_counter()
    l1 = res.launch_count
    clk = clocks.summary(t0, t1)
    clocks.stop()
    samples_step = res.last_samples
    kms = np.max(np.stack(res.each(lambda k, g: g.last_kernel_ms().astype(np.float64))), axis=0)
    value = samples_step * args.steps / (ms * 1e-3) / 1e6
    fsk_gbs = ALG_BYTES_PER_SAMPLE_FSK * (samples_step / G) / (kms[0] * 1e-3) / 1e9
    res.close()

    # ---- e2e: cu8 host buffers, streams placed by the measured concurrent copy rates ----
    ec = min(chunk, args.e2e_bytes // 2)
    rates = probe_copy_rates(range(G), seconds=1.0)
    weighted = min(rates) < 0.9 * max(rates)
    me = MultiEngine(total, devices=range(G), weights=rates if weighted else None, engines_per_device=2, in_fmt="cu8",
                     framing="v1", chunk_samples=ec + 1024)
    pb = me.pinned_block(ec)
    for s_ in range(total):
        src = sources[s_ % n_src]
        r = ((s_ // n_src) * 4688) % (chunk - ec) if chunk > ec else 0
        pb.array[s_, :] = to_cu8(src[2 * r:2 * r + 2 * ec])
    if port is not None:
        me.step(pb.array)
        me.sync()
        rows = check_rows(total, n_src)
        sd_ok = pk_ok = True
        for s_ in rows:
            sd_o, ref = oracle_decode(port, pb.array[s_], "cu8")
            sd_g = me.drain_soft(s_)
            sd_ok &= sd_g.size == sd_o.size and bool(np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32)))
            pk_ok &= me.drain_packets(s_) == ref["packets"]
        me.drain_all_packets()
        parity.update(e2e_streams_checked=len(rows), rows=rows, e2e_sd_equal=bool(sd_ok), e2e_packets_equal=bool(pk_ok))
        ok_all &= sd_ok and pk_ok
    for _ in range(max(args.warmup, 1)):
        me.stream_step(pb.array)
    me.flush()
    npk = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        npk += len(me.stream_step(pb.array))
    npk += len(me.flush())
    dt = time.perf_counter() - t0
    consumed = me.last_samples * args.steps
    h2d = total * ec * 2
    e2e = {"value": round(consumed / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(npk // args.steps * 264), "ms_per_step": round(1e3 * dt / args.steps, 3), "in_fmt": "cu8",
           "chunk_samples": ec, "engines": 2 * G, "crc_valid_packets_per_step": npk // args.steps,
           "h2d_gbs": round(h2d * args.steps / dt / 1e9, 2), "h2d_rate_per_gpu_gbs": [round(x, 1) for x in rates],
           "h2d_ceiling_gbs": round(sum(rates), 1), "frac_of_ceiling": round(h2d * args.steps / dt / 1e9 / sum(rates), 4),
           "placement": "rate-weighted (MultiEngine weights = probe_copy_rates)" if weighted else "equal blocks",
           "streams_per_engine": [b - a for a, b in me.ranges],
           "api": "MultiEngine.stream_step(pinned host block) + flush: wb_feed_strided + wb_process + wb_sync + wb_drain_all_packets per engine thread"}
    me.close()
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, G), parallelism="wenet_b200.MultiEngine: one process, one feeder thread per GPU, "
                                                                 "no torch, no collective"),
            "clocks": clk, "e2e": e2e, "gpu_launches": int(l1 - l0), "parity": parity,
            "roofline": {"bound": "hbm", "kernel": "wb_fsk_kernel", "achieved": round(fsk_gbs, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(fsk_gbs / peak, 4), "traffic": None,
                         "kernel_ms": {"fsk": round(float(kms[0]), 3), "deframe": round(float(kms[1]), 3),
                                       "llr_stats": round(float(kms[2]), 3), "ldpc": round(float(kms[3]), 3)}},
            "cpu_baseline": None}
    if not ok_all:
        line.update(value=None, e2e=None, error="parity gate failed: see `parity`")
    print(json.dumps(line), flush=True)
    return 0 if ok_all else 1


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
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configurations (the `extra` object)")
    ap.add_argument("--no-parity", action="store_true", help="exploration only: a line without the parity gate says so")
    ap.add_argument("--e2e-all-formats", action="store_true", help="also time the e2e leg with cs16 and cf32 host buffers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return 0

    if world == 1 and args.gpus > 1 and args.mode == "v1":
        return run_multiengine(args)             # N GPUs from one process: the MultiEngine host object (no torchrun, no torch)

    R = Ranks(world, local)
    from wenet_b200 import engine as E          # raises if libwenet_b200.so is missing: no CPU fallback
    from wenet_b200 import sharding
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    if args.mode == "ldpc":
        res = bench_ldpc(E, R, args, local, rank, args.codewords, [args.ldpc_iter], args.steps)
        total, ms, iters = res[args.ldpc_iter]
        if rank == 0:
            print(json.dumps({"metric": "LDPC Mcodewords/s (exploration mode, BASELINE configs[2])", "value": round(total / (ms * 1e-3) / 1e6, 3),
                              "unit": "Mcodewords/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "dtype": "f32",
                              "data": "synthetic", "config": {"workload": "H2064_516 (2580, 2064), %d codewords/GPU resident, max_iter %d, "
                                                              "iterations of the 24 sources %s" % (args.codewords, args.ldpc_iter, iters)},
                              "roofline": {"bound": "hbm", "achieved": round(10582.0 * total / (ms * 1e-3) / 1e9, 2), "unit": "GB/s",
                                           "note": "10 582 algorithmic bytes per codeword; the kernel is shared-memory / issue bound"}}),
                  flush=True)
        R.close()
        return 0

    n, chunk = args.streams, args.chunk
    n_src = min(args.sources, n)
    sources = make_sources(n_src, chunk, seed_base=rank * 1000, mode="v1" if args.mode == "fskonly" else args.mode)
    if args.mode != "v1":
        args.no_e2e = args.no_cpu_baseline = args.no_extra = args.no_parity = True
    ekw = {"v1": dict(framing="v1"), "v2": dict(Fs=960000, Rs=96000, framing="v2"), "fsk4": dict(M=4, framing="none"),
           "fskonly": dict(framing="none")}[args.mode]

    port = None
    if not args.no_parity:
        from oracle import oracle as O           # the checker (test infrastructure): never inside a timed region
        if rank == 0:
            O.build()                            # (one rank compiles it if the copy on this box looks stale; the others wait)
        R.barrier()
        port = O.Oracle("port")
    parity = {"oracle": "CPU restatement oracle/liboracle.so (pinned against the compiled reference, tests/test_oracle_vs_ref.py)"} \
        if port else {"skipped": True}

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

    # ---- parity gate on the HBM-resident launch: the very first pass of the 4096-stream batch from fresh state, the
    #      checked rows read back from the device (= exactly what the kernel consumed) and re-done by the oracle ----
    ok_all = True
    if port is not None:
        rows = check_rows(n, n_src)
        one_step()
        eng.sync()
        cws = eng.drain_codewords()
        sd_ok = pk_ok = it_ok = True
        npk_checked = 0
        for s in rows:
            raw = eng.dev_read_input(s, fill)
            sd_o, res = oracle_decode(port, raw, "cf32")
            sd_g = eng.drain_soft(s)
            sd_ok &= (sd_g.size == sd_o.size) and bool(np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32)))
            pk_ok &= eng.drain_packets(s) == res["packets"]
            it_ok &= cws["iters"][cws["stream"] == s].tolist() == res["iters"].tolist()
            npk_checked += len(res["packets"]) // 256
        eng.drain_all_packets()
        parity.update(streams_checked=len(rows), rows=rows, samples_per_row=int(fill), sd_sha_equal=bool(sd_ok),
                      packets_equal=bool(pk_ok), iters_equal=bool(it_ok), packets_checked=int(npk_checked))
        ok_all &= sd_ok and pk_ok and it_ok

    attempts = 0
    while True:
        attempts += 1
        clocks = ClockSampler(local)
        for _ in range(args.warmup):
            one_step()
        eng.sync()
        eng.drain_all_packets()          # warm-up output is not part of the report
        R.barrier()
        l0 = eng.launch_count
        t0 = time.perf_counter()
        eng.timer_start()
        for _ in range(args.steps):
            one_step()
        ms = eng.timer_stop()
        t1 = time.perf_counter()
        l1 = eng.launch_count
        R.barrier()
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
    # decoded bytes per Eb/N0 over the whole batch of the last timed step (stream s is at EBNO_SWEEP[(s % n_src) % 5])
    by_eb = {}
    if args.mode == "v1" and args.synth == "host":
        eb_of = np.array([EBNO_SWEEP[(s_ % n_src) % len(EBNO_SWEEP)] for s_ in range(n)])
        cnt = np.bincount(pk["stream"], minlength=n) if len(pk) else np.zeros(n, dtype=np.int64)
        for eb in EBNO_SWEEP:
            sel = eb_of == eb
            by_eb["%g" % eb] = {"streams": int(sel.sum()), "decoded_bytes": int(cnt[sel].sum()) * 256}
    ms = R.max(ms)
    total_samples = R.sum(float(samples_step)) * args.steps
    value = total_samples / (ms * 1e-3) / 1e6

    alg_bps = 9.0 if args.mode == "fsk4" else ALG_BYTES_PER_SAMPLE_FSK
    fsk_gbs = alg_bps * samples_step / (kms[0] * 1e-3) / 1e9 if kms[0] > 0 else 0.0
    st_fsk = ncu_static("wb_fsk_kernel", n, chunk, args.mode)
    st_ldpc = ncu_static("wb_ldpc_kernel", n, chunk, args.mode)
    traffic = (st_fsk["dram_bytes_read"] + st_fsk["dram_bytes_write"]) if st_fsk else None
    roofline = {"bound": "hbm", "kernel": "wb_fsk_kernel", "achieved": round(fsk_gbs, 2), "peak": peak, "unit": "GB/s",
                "frac": round(fsk_gbs / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": {"fsk": round(float(kms[0]), 3), "deframe": round(float(kms[1]), 3),
                              "llr_stats": round(float(kms[2]), 3), "ldpc": round(float(kms[3]), 3)},
                "alg_bytes_per_launch": alg_bps * samples_step,
                "issue_active_pct": st_fsk.get("issue_active_pct") if st_fsk else None,
                "warp_inst_per_sample": st_fsk.get("warp_inst_per_sample") if st_fsk else None,
                "note": "the kernel is a chain of dependent fp32 phases per frame (bit-exact with the reference's operation "
                        "order): issue slots, not HBM, are the binding resource; see DESIGN.md section 4"}
    # the decoder is bound by shared-memory bandwidth, not HBM (SURVEY 8d asks for that estimate beside the HBM fraction):
    # wavefronts per launch from the ncu capture of this workload / (its live event-timed duration x one wavefront per
    # cycle and SM)
    roofline_ldpc = None
    if args.mode == "v1" and kms[3] > 0:
        sm_clk = float(clk.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        roofline_ldpc = {"bound": "smem", "kernel": "wb_ldpc_kernel", "codewords_per_launch": int(codewords_step),
                         "mcodewords_per_s": round(codewords_step / (kms[3] * 1e-3) / 1e6, 3),
                         "hbm_achieved_gbs": round(10582.0 * codewords_step / (kms[3] * 1e-3) / 1e9, 2),
                         "hbm_frac": round(10582.0 * codewords_step / (kms[3] * 1e-3) / 1e9 / peak, 4),
                         "ideal_wavefronts_per_iter": 1354,
                         "note": "ideal = (516 checks x 14 slots + 2064 x 3 + 516 x 2 variable edges) x (message load + phi0 "
                                 "look-up + message store) / 32 lanes, reference src/mpdecode_core.c:385-489"}
        if st_ldpc and st_ldpc.get("smem_wavefronts"):
            wf = float(st_ldpc["smem_wavefronts"])
            roofline_ldpc.update(smem_wavefronts_per_launch=wf, peak_wavefronts_per_s=148 * sm_clk,
                                 achieved_frac_of_smem_peak=round(wf / (kms[3] * 1e-3) / (148 * sm_clk), 4),
                                 wavefronts_per_codeword=round(wf / max(1, st_ldpc.get("codewords", codewords_step)), 1))
    eng.close()

    # ---- e2e through the public API with host buffers ----
    ceiling = None

    def measure_ceiling():
        """what the host -> device leg alone can do with all ranks copying at once: flat copies from pinned memory for a
        fixed wall time on every rank (wb_copy_probe_*), rates gathered"""
        pr = E.CopyProbe(local, 256 << 20)
        try:
            pr.run(2)
            R.barrier()
            t0_ = time.perf_counter()
            nb = 0
            while time.perf_counter() - t0_ < 1.0:
                pr.run(1)
                nb += pr.nbytes
            rate = nb / (time.perf_counter() - t0_) / 1e9
            R.barrier()
        finally:
            pr.close()
        return R.gather(rate)

    def run_e2e(fmt, n_eng, n_mine, row0, check):
        """feed (pinned host -> HBM) + process + sync + drain (HBM -> host) every step.  This rank's n_mine streams
        (global rows row0 ..) are split over n_eng engines on the same GPU and each engine's drain of step k is issued
        right before its feed of step k + 1, so one engine's copies overlap the other's kernels across step boundaries."""
        ec = min(chunk, args.e2e_bytes // E.FMT_BPS[fmt])
        per = [n_mine // n_eng + (1 if i < n_mine % n_eng else 0) for i in range(n_eng)]
        elems = E.FMT_ELEMS[fmt]
        engs, pins, base = [], [], 0
        where = {}
        for i, cnt in enumerate(per):
            engs.append(E.Engine(cnt, in_fmt=fmt, framing="v1", chunk_samples=ec + 1024, device=local))
            pb = E.PinnedBuffer((cnt, elems * ec), E.FMT_DTYPE[fmt])
            for j in range(cnt):
                s_ = base + j
                where[s_] = (i, j)
                src = sources[s_ % n_src]
                r = ((s_ // n_src) * 4688) % (chunk - ec) if chunk > ec else 0
                seg = src[2 * r:2 * r + 2 * ec]
                if fmt == "cf32":
                    pb.array[j, :] = seg
                elif fmt == "cs16":                     # siggen.to_format: the reference divides by 1000 (FDMDV_SCALE)
                    pb.array[j, :] = np.round(seg.astype(np.float64) * 1000.0).astype(np.int16)
                else:                                   # the same samples as rtl_sdr would deliver them (cu8)
                    pb.array[j, :] = to_cu8(seg)
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

        res_par = None
        if check and port is not None:
            # parity gate on the host-buffer path: the first step of the fresh engines, checked rows re-done by the oracle
            # from the bytes in the pinned buffers; rows 0..4 (rank 0) double as the Eb/N0 ladder of this arm
            for i, (g, pb) in enumerate(zip(engs, pins)):
                g.feed_strided(pb.array)
                g.process()
            sd_ok = pk_ok = True
            rows_ = [s_ for s_ in check_rows(n_mine, n_src) if s_ in where]
            ladder = []
            for s_ in rows_:
                i, j = where[s_]
                g = engs[i]
                g.sync()
                sd_o, res = oracle_decode(port, pins[i].array[j], fmt)
                sd_g = g.drain_soft(j)
                sd_ok &= (sd_g.size == sd_o.size) and bool(np.array_equal(sd_g.view(np.uint32), sd_o.view(np.uint32)))
                got = g.drain_packets(j)
                pk_ok &= got == res["packets"]
                if s_ < len(EBNO_SWEEP) and row0 == 0:
                    ladder.append({"ebno_db": EBNO_SWEEP[s_], "samples": int(ec), "bytes": len(got)})
            for g in engs:
                g.sync()
                g.drain_all_packets()
            res_par = {"streams_checked": len(rows_), "sd_equal": bool(sd_ok), "packets_equal": bool(pk_ok), "ladder": ladder}

        for _ in range(max(args.warmup, 1)):
            step()
        flush()
        R.barrier()
        stat.update(d2h=0, samples=0, packets=0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        flush()                                     # every step's results are on the host when the clock stops
        dt = time.perf_counter() - t0
        R.barrier()
        dt_ranks = R.gather(dt)
        dt = max(dt_ranks)
        consumed = R.sum(float(stat["samples"]))
        h2d = R.sum(float(n_mine * ec * E.FMT_BPS[fmt]))
        res = {"value": round(consumed / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(R.sum(float(stat["d2h"])) // args.steps), "ms_per_step": round(1e3 * dt / args.steps, 3),
               "in_fmt": fmt, "chunk_samples": ec, "engines": n_eng,
               "crc_valid_packets_per_step": int(R.sum(float(stat["packets"])) // args.steps),
               "h2d_gbs": round(h2d * args.steps / dt / 1e9, 2), "s_per_rank": [round(x, 4) for x in dt_ranks],
               "api": "wb_feed_strided(pinned host) + wb_process + wb_sync + wb_drain_all_packets"}
        for g in engs:
            g.close()
        del pins
        return res, res_par

    e2e = e2e_cf32 = e2e_cs16 = e2e_equal = None
    ladder_gpu = None
    if not args.no_e2e:
        rates = measure_ceiling()
        if os.environ.get("WB_BENCH_SKEW_RATES"):          # test hook: pretend the ranks reach unequal rates (exercises the weighted path on any box)
            rates = [r * (1.0 - 0.3 * (i % 2)) for i, r in enumerate(rates)]
        ceiling = sum(rates)
        # placement of the job's n x world streams over the ranks: equal blocks, or -- when the GPUs of the box do not reach
        # the same host -> device rate with all of them copying (HGX: GPUs behind a busier host bridge) -- blocks in
        # proportion to the measured rates, so that every rank's feed takes the same time (sharding.weighted_ranges)
        weighted = world > 1 and min(rates) < 0.9 * max(rates)
        rng_eq = sharding.equal_ranges(n * world, world)
        rng_w = sharding.weighted_ranges(n * world, rates) if weighted else rng_eq
        lo, hi = rng_w[rank]
        e2e, par_e2e = run_e2e("cu8", 2, hi - lo, lo, True)
        placement = "equal blocks"
        if weighted:
            placement = "rate-weighted (sharding.weighted_ranges over the measured concurrent copy rates)"
            # one refinement from what the run itself showed: ranks that took longer than the others get fewer streams
            # (the 0.6 s flat-copy probe is only a first estimate of what a rank sustains with the strided feed)
            t_r = e2e["s_per_rank"]
            if max(t_r) > 1.04 * min(t_r):
                w2 = [(b - a) / t for (a, b), t in zip(rng_w, t_r)]
                rng_2 = sharding.weighted_ranges(n * world, w2)
                lo2, hi2 = rng_2[rank]
                e2e_2, _ = run_e2e("cu8", 2, hi2 - lo2, lo2, False)
                if e2e_2["value"] > e2e["value"]:
                    e2e, rng_w = e2e_2, rng_2
                    placement = "rate-weighted, refined once from the measured per-rank step times (sharding.weighted_ranges)"
        e2e.update(placement=placement,
                   streams_per_rank=[b - a for a, b in rng_w], h2d_rate_per_rank_gbs=[round(x, 1) for x in rates],
                   h2d_ceiling_gbs=round(ceiling, 1), frac_of_ceiling=round(e2e["h2d_gbs"] / ceiling, 4),
                   ceiling_note="sum over ranks of the pinned-host -> device rate each GPU reaches with all ranks copying at once "
                                "(wb_copy_probe, measured in this run)")
        if weighted:
            e2e_equal, _ = run_e2e("cu8", 2, n, rank * n, False)
            e2e_equal.update(placement="equal blocks", h2d_ceiling_gbs=round(ceiling, 1),
                             frac_of_ceiling=round(e2e_equal["h2d_gbs"] / ceiling, 4),
                             note="with equal blocks every rank moves the same bytes, so the job runs at world x the slowest rank's rate")
        if par_e2e is not None:
            bad_any = R.min(1.0 if (par_e2e["sd_equal"] and par_e2e["packets_equal"]) else 0.0) < 1.0
            parity.update(e2e_streams_checked=par_e2e["streams_checked"], e2e_sd_equal=par_e2e["sd_equal"],
                          e2e_packets_equal=par_e2e["packets_equal"])
            ok_all &= not bad_any
            ladder_gpu = par_e2e["ladder"]
        if args.e2e_all_formats or world == 1:
            e2e_cs16, _ = run_e2e("cs16", 2, n, rank * n, False)
            e2e_cf32, _ = run_e2e("cf32", 2, n, rank * n, False)

    # ---- decoded bytes per Eb/N0, both arms on the same five streams (rank 0) ----
    ladder = None
    if rank == 0 and ladder_gpu:
        try:
            ref_rows, kind = reference_ladder(ladder_rows(args))
            same = [a["bytes"] == b["bytes"] and a["samples"] == b["samples"] for a, b in zip(ladder_gpu, ref_rows)]
            ladder = {"streams": "rows 0..4 of the e2e cu8 host buffers: one v1 stream per Eb/N0, %d samples each" % ladder_gpu[0]["samples"],
                      "reference_arm": kind,
                      "rows": [{"ebno_db": a["ebno_db"], "gpu_bytes": a["bytes"], "reference_bytes": b["bytes"]} for a, b in zip(ladder_gpu, ref_rows)],
                      "equal": bool(all(same) and len(same) == len(EBNO_SWEEP)),
                      "whole_batch_last_step": by_eb}
            parity["ladder_equal"] = ladder["equal"]
            ok_all &= ladder["equal"]
        except Exception as ex:
            ladder = {"error": str(ex)[:200]}
    ok_all = R.min(1.0 if ok_all else 0.0) >= 1.0

    # ---- the other BASELINE configurations (after the headline; each with its own roofline fraction) ----
    extra = None
    if not args.no_extra:
        extra = {}
        try:
            tot, ms1, k1, spb1 = bench_fsk_only(E, R, local, sources, 1024, chunk, 2, 2, 5)
            g1 = ALG_BYTES_PER_SAMPLE_FSK * (tot / world) / (k1 * 1e-3) / 1e9
            extra["configs[1]"] = {"workload": "FSK-demod only: 1024 streams/GPU x %d samples, 2-FSK 115177 baud, cf32 resident" % chunk,
                                   "value": round(tot * 5 / (ms1 * 1e-3) / 1e6, 2), "unit": UNIT, "ms_per_step": round(ms1 / 5, 3),
                                   "streams_per_cta": spb1,
                                   "roofline": {"bound": "hbm", "achieved": round(g1, 2), "peak": peak, "unit": "GB/s", "frac": round(g1 / peak, 4)}}
        except Exception as ex:
            extra["configs[1]"] = {"error": str(ex)[:200]}
        try:
            res = bench_ldpc(E, R, args, local, rank, 1 << 20, [10, 100], 3)
            for mi in (10, 100):
                tot, msl, iters = res[mi]
                g2 = 10582.0 * (tot / world) / (msl * 1e-3) / 1e9
                extra["configs[2] max_iter %d" % mi] = {
                    "workload": "LDPC only: 1048576 H2064_516 codewords/GPU of LLRs resident, max_iter %d (24 seeded noisy codewords "
                                "replicated; their iteration counts %s)" % (mi, iters),
                    "value": round(tot / (msl * 1e-3) / 1e6, 3), "unit": "Mcodewords/s", "ms_per_step": round(msl / 3, 3),
                    "roofline": {"bound": "hbm", "achieved": round(g2, 2), "peak": peak, "unit": "GB/s", "frac": round(g2 / peak, 4),
                                 "note": "10 582 algorithmic bytes per codeword; the decoder is shared-memory bound (roofline_ldpc)"}}
        except Exception as ex:
            extra["configs[2]"] = {"error": str(ex)[:200]}
        try:
            src4 = make_sources(8, chunk, seed_base=rank * 1000 + 500, mode="fsk4")
            tot, ms4, k4, spb4 = bench_fsk_only(E, R, local, src4, 1024, chunk, 4, 2, 5)
            g4 = 9.0 * (tot / world) / (k4 * 1e-3) / 1e9
            extra["configs[4]"] = {"workload": "4-FSK demod only: 1024 streams/GPU (8192 over 8 GPUs) x %d-sample chunks, cf32 resident" % chunk,
                                   "value": round(tot * 5 / (ms4 * 1e-3) / 1e6, 2), "unit": UNIT, "ms_per_step": round(ms4 / 5, 3),
                                   "streams_per_cta": spb4,
                                   "roofline": {"bound": "hbm", "achieved": round(g4, 2), "peak": peak, "unit": "GB/s", "frac": round(g4 / peak, 4)}}
        except Exception as ex:
            extra["configs[4]"] = {"error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(32 << 20)

    rc = 0
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clk, "e2e": e2e, "e2e_equal_blocks": e2e_equal, "e2e_cs16": e2e_cs16, "e2e_cf32": e2e_cf32,
            "gpu_launches": int(l1 - l0), "parity": parity, "ebn0_ladder": ladder,
            "roofline": roofline, "roofline_ldpc": roofline_ldpc, "cpu_baseline": cpu,
            "work_per_step_per_gpu": {"samples": int(samples_step), "codewords": int(codewords_step),
                                      "crc_valid_packets": packets_last},
            "extra": extra,
        }
        if not ok_all:
            # a number whose results differ from the reference's is not a number: no value, non-zero exit
            line.update(value=None, e2e=None, error="parity gate failed: see `parity`")
            rc = 1
        print(json.dumps(line), flush=True)
    elif not ok_all:
        rc = 1
    R.close()
    return rc


if __name__ == "__main__":
    sys.exit(main())
