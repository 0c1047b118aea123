#!/usr/bin/env python3
"""wenet_rx [--cu8|--cs16] [--v2] [Fs Rs] In Out -- `fsk_demod ... | drs232_ldpc - -` (or `| wenet_ldpc - -`) fused in
one process: IQ in, 256-byte CRC-valid packets out, soft decisions never leave the GPU.  A drop-in for the two
middle stages of start_rx.sh:125-128 when both would run on the same machine."""
import sys

import numpy as np


def main(argv=None):
    argv = list(argv or sys.argv)
    fmt, framing = "s16", "v1"
    args = []
    for a in argv[1:]:
        if a in ("--cu8", "-d"):
            fmt = "cu8"
        elif a in ("--cs16", "-c"):
            fmt = "cs16"
        elif a == "--v2":
            framing = "v2"
        else:
            args.append(a)
    if len(args) not in (2, 4):
        sys.stderr.write("usage: %s [--cu8|--cs16] [--v2] [SampleRate SymbolRate] In Out\n" % argv[0])
        sys.exit(1)
    Fs, Rs = (int(args[0]), int(args[1])) if len(args) == 4 else ((921416, 115177) if framing == "v1" else (960000, 96000))
    fin = sys.stdin.buffer if args[-2] == "-" else open(args[-2], "rb")
    fout = sys.stdout.buffer if args[-1] == "-" else open(args[-1], "wb")
    from wenet_b200 import engine as E
    eng = E.Engine(1, Fs=Fs, Rs=Rs, in_fmt=fmt, framing=framing, chunk_samples=1 << 17)
    bps = E.FMT_BPS[fmt]
    while True:
        raw = fin.read((1 << 16) * bps)
        if not raw:
            break
        raw = raw[:len(raw) - len(raw) % bps]
        eng.feed([np.frombuffer(raw, dtype=E.FMT_DTYPE[fmt])])
        eng.process()
        pk = eng.drain_packets(0)
        if pk:
            fout.write(pk)
            fout.flush()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
