#!/usr/bin/env python3
"""fsk_demod [-l] [-p P] [-s] [(-c|-d)] [-t[r]] [-f] [-b lo] [-u hi] (2|4) SampleRate SymbolRate In Out

Same argv, input formats, output stream and exit codes as reference src/fsk_demod.c:89-206, demodulating on the
GPU through libwenet_b200.so.  Differences, all outside the hot path: -l (low-rate mode: another frame geometry,
used by no Wenet script) is not implemented and exits 1; the stats JSON (stderr, -t) has the reference's keys (secs,
EbNodB, ppm, f1_est, f2_est[, f3_est, f4_est], eye_diagram, samp_fft -- what rx/fskstatsudp.py parses) at the reference's cadence, with
the statistics as they stand at the end of a block of frames (with -t a block is one stats period).  -f (testframe mode, src/fsk_demod.c:226-243, :304-343) counts bit errors against the reference's
known frame on the host (wenet_b200/cli/_testframes.py): the same "errs: ..." lines, or with -t one JSON line per modem
frame that completed a testframe (modem statistics in it are the block's, the frames/bits/errs counters exact).
Without -s the output is the demodulator's own hard bits (arg-max tone, src/fsk.c:936-959), one byte per bit.
Output is written per block of frames, not per frame.
"""
import getopt
import json
import math
import signal
import sys
import time

import numpy as np

USAGE = ("usage: %s [-l] [-p P]  [-s] [(-c|-d)] [-t [r]] [-f] (2|4) SampleRate SymbolRate InputModemRawFile OutputFile\n"
         " -p P --conv=P     -  P specifies the rate at which symbols are down-converted before further processing\n"
         " -c --cs16         -  The raw input file will be in complex signed 16 bit format.\n"
         " -d --cu8          -  The raw input file will be in complex unsigned 8 bit format.\n"
         "                        If neither -c nor -d are used, the input should be in signed 16 bit format.\n"
         " -t[r] --stats=[r] -  Print out modem statistics to stderr in JSON.\n"
         " -s --soft-dec     -  The output file will be in a soft-decision format, with one 32-bit float per bit.\n"
         " -b lo -u hi       -  frequency estimator limits in Hz\n")

BLOCK_FRAMES = 64


def usage(prog, msg=None):
    if msg:
        sys.stderr.write(msg + "\n")
    sys.stderr.write(USAGE % prog)
    sys.exit(1)


def _atoi(text):
    """C atoi(): optional sign and leading digits, 0 if there are none (the reference parses every number with it)"""
    import re
    m = re.match(r"\s*([+-]?\d+)", text or "")
    return int(m.group(1)) if m else 0


def _split_stats_option(args):
    """-t[r] / --stats[=r] takes an OPTIONAL argument (getopt_long "t::", src/fsk_demod.c:92-110), which Python's getopt
    does not know: the value only counts when attached (`-t100`, `--stats=100`, the form every start_rx script uses),
    a bare `-t` / `--stats` never swallows the next word.  Returns (remaining args, enabled, rate string or None)."""
    rest, enabled, rate = [], False, None
    takes_value = False                       # the previous word was -p / -b / -u: this one is its value
    for a in args:
        if takes_value or a == "--" or not a.startswith("-") or a == "-":
            takes_value = False
            rest.append(a)
            continue
        if a == "--stats" or a.startswith("--stats="):
            enabled, rate = True, (a[8:] if a.startswith("--stats=") else None)
            continue
        if a in ("--fsk_lower", "--fsk_upper"):   # optional_argument without a value: accepted and ignored (optarg NULL)
            continue
        if not a.startswith("--"):
            k = 1
            while k < len(a) and a[k] in "fhlcds":      # flags without an argument may be clustered in front
                k += 1
            if k < len(a) and a[k] == "t":
                enabled, rate = True, (a[k + 1:] or None)
                if k > 1:
                    rest.append(a[:k])
                continue
            takes_value = a[-1] in "pbu" and all(c in "fhlcds" for c in a[1:-1])
        rest.append(a)
    return rest, enabled, rate


def parse(argv):
    rest, stats_on, stats_rate = _split_stats_option(argv[1:])
    try:
        opts, args = getopt.gnu_getopt(rest, "fhlp:cdsb:u:",
                                       ["help", "lbr", "conv=", "cs16", "cu8", "fsk_lower=", "fsk_upper=", "soft-dec",
                                        "testframes"])
    except getopt.GetoptError as e:
        usage(argv[0], str(e))
    o = dict(fmt="s16", soft=False, stats=stats_on, stats_rate=8, P=0, lo=0, hi=0, testframes=False)
    if stats_rate:
        o["stats_rate"] = _atoi(stats_rate) or 8          # atoi() == 0 -> 8, src/fsk_demod.c:126-129
    for k, v in opts:
        if k in ("-h", "--help"):
            usage(argv[0])
        elif k in ("-l", "--lbr"):
            usage(argv[0], "low-rate mode (-l) is not supported by the B200 engine")
        elif k in ("-f", "--testframes"):
            o["testframes"] = True
        elif k in ("-c", "--cs16"):
            o["fmt"] = "cs16"
        elif k in ("-d", "--cu8"):
            o["fmt"] = "cu8"
        elif k in ("-s", "--soft-dec"):
            o["soft"] = True
        elif k in ("-p", "--conv"):
            o["P"] = _atoi(v)
        elif k in ("-b", "--fsk_lower"):
            o["lo"] = _atoi(v)
        elif k in ("-u", "--fsk_upper"):
            o["hi"] = _atoi(v)
    if len(args) < 5:
        usage(argv[0], "Too few arguments")
    if len(args) > 5:
        usage(argv[0], "Too many arguments")
    o["M"], o["Fs"], o["Rs"] = _atoi(args[0]), _atoi(args[1]), _atoi(args[2])      # atoi(), src/fsk_demod.c:181-183
    if o["M"] not in (2, 4):
        usage(argv[0], "Mode %d is not valid. Mode must be 2 or 4." % o["M"])
    o["fin"], o["fout"] = args[3], args[4]
    return o


def stats_line(st, M):
    """One stderr line of the reference's `-t` / `--stats` output (src/fsk_demod.c:345-392) from a wb_stats record:
    the keys rx/fskstatsudp.py requires (EbNodB, ppm, f1_est, f2_est, samp_fft) plus eye_diagram and, for 4-FSK,
    f3_est / f4_est."""
    ppm = int(st.ppm) if math.isfinite(st.ppm) else 0           # C's (int) of a non-finite float does not raise
    d = {"secs": int(time.time()), "EbNodB": round(st.EbNodB, 1), "ppm": ppm,
         "f1_est": round(st.f_est[0], 1), "f2_est": round(st.f_est[1], 1)}
    if M == 4:
        d["f3_est"], d["f4_est"] = round(st.f_est[2], 1), round(st.f_est[3], 1)
    d["eye_diagram"] = [[round(float(st.rx_eye[i][j]), 6) for j in range(st.neyesamp)] for i in range(st.neyetr)]
    d["samp_fft"] = [round(float(v), 6) for v in st.samp_fft[:st.nfft]]
    return json.dumps(d)


def main(argv=None):
    argv = argv or sys.argv
    o = parse(argv)
    from wenet_b200 import engine as E
    try:
        fin = sys.stdin.buffer if o["fin"] == "-" else open(o["fin"], "rb")
        fout = sys.stdout.buffer if o["fout"] == "-" else open(o["fout"], "wb")
    except OSError:
        sys.stderr.write("Couldn't open files\n")
        sys.exit(1)
    limits = (o["lo"], o["hi"]) if (o["lo"] > 0 and o["hi"] > o["lo"]) else None
    try:
        eng = E.Engine(1, Fs=o["Fs"], Rs=o["Rs"], M=o["M"], P=o["P"], in_fmt=o["fmt"], framing="none",
                       chunk_samples=(BLOCK_FRAMES + 2) * 512, est_limits=limits, stats=o["stats"], hard_bits=not o["soft"])
    except E.WbError as e:
        sys.stderr.write("Couldn't open files\n%s\n" % e)
        sys.exit(1)
    if limits:
        sys.stderr.write("Setting estimator limits to %d to %d Hz.\n" % limits)
    signal.signal(signal.SIGTERM, lambda *_: sys.exit(0))
    bps = E.FMT_BPS[o["fmt"]]
    dt = E.FMT_DTYPE[o["fmt"]]
    stats_every = int(1 / (o["stats_rate"] * eng.N / o["Fs"])) + 1 if o["stats"] else 0
    frames_since = 0
    # with -t the blocks shrink to the reference's stats cadence (one JSON line every stats_loop + 1 frames,
    # src/fsk_demod.c:247-251, :394-400), so that fskstatsudp.py sees as many lines per second as from the reference
    block_frames = max(1, min(BLOCK_FRAMES, stats_every)) if (o["stats"] and not o["testframes"]) else BLOCK_FRAMES
    tf = None
    if o["testframes"]:
        from wenet_b200.cli._testframes import TestFrames
        tf = TestFrames()
    # a pipe hands over whatever has arrived (read1): frames go out as soon as their samples are in, like the reference's
    # per-frame fread/fwrite; a regular file comes in full blocks.  A split sample waits for its other half.
    read = getattr(fin, "read1", fin.read)
    left = b""
    while True:
        raw = read(block_frames * eng.N * bps)
        if not raw:
            break
        raw = left + raw
        left = raw[len(raw) - len(raw) % bps:]
        raw = raw[:len(raw) - len(left)]
        if not raw:
            continue
        eng.feed([np.frombuffer(raw, dtype=dt)])
        eng.process()
        eng.sync()
        sd = eng.drain_soft(0)
        if sd.size:
            if o["soft"]:
                fout.write(sd.tobytes())
            else:
                fout.write(eng.drain_hard(0).tobytes())
            fout.flush()
        if tf is not None and sd.size:
            # src/fsk_demod.c:304-343: soft mode slices sd < 0 (whatever M is), hard mode takes the demodulator's bits
            hits = tf.feed((sd < 0).astype(np.uint8) if o["soft"] else eng.drain_hard(0))
            if not o["stats"]:
                for h in hits:
                    sys.stderr.write(tf.line(h))
            elif hits:
                st = eng.stats(0)
                last = {}
                for h in hits:                      # one line per modem frame that saw a testframe, counters as at its end
                    last[h[0] // eng.Nbits] = h
                for h in last.values():
                    d = '{"secs": %d, "EbNodB": %5.1f, "ppm": %4d, "f1_est":%.1f, "f2_est":%.1f' % (
                        int(time.time()), st.EbNodB, int(st.ppm) if math.isfinite(st.ppm) else 0, st.f_est[0], st.f_est[1])
                    if o["M"] == 4:
                        d += ', "f3_est":%.1f, "f4_est":%.1f' % (st.f_est[2], st.f_est[3])
                    sys.stderr.write(d + ', "frames":%d, "bits":%d, "errs":%d}\n' % (h[2], h[3], h[4]))
            continue
        if o["stats"] and sd.size:
            frames_since += sd.size // eng.Nbits
            if frames_since >= stats_every:
                frames_since -= stats_every       # nin wanders around N: keep the average cadence
                sys.stderr.write(stats_line(eng.stats(0), o["M"]) + "\n")
    fout.flush()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
