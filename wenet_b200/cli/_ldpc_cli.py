"""shared body of the drs232_ldpc / wenet_ldpc stand-ins (reference src/drs232_ldpc.c:105-286, src/wenet_ldpc.c)"""
import sys

import numpy as np

BLOCK_SYMS = 1 << 15


def _per(errors, packets):
    """(float)packet_errors/packets through %4.3f: 0/0 is the x86 default NaN, which glibc prints as -nan"""
    if packets == 0:
        return "-nan" if errors == 0 else "inf"
    return "%4.3f" % float(np.float32(errors) / np.float32(packets))


def main(argv, framing, name):
    if len(argv) < 3:
        sys.stderr.write("usage: %s InputOneSymbolPerFloat OutputPackets [-v[v]]\n" % name)
        sys.exit(1)
    try:
        fin = sys.stdin.buffer if argv[1] == "-" else open(argv[1], "rb")
    except OSError as e:
        sys.stderr.write("Error opening input file: %s: %s.\n" % (argv[1], e.strerror))
        sys.exit(1)
    try:
        fout = sys.stdout.buffer if argv[2] == "-" else open(argv[2], "wb")
    except OSError as e:
        sys.stderr.write("Error opening output file: %s: %s.\n" % (argv[2], e.strerror))
        sys.exit(1)
    verbose = {"-v": 1, "-vv": 2}.get(argv[3], 0) if len(argv) > 3 else 0
    from wenet_b200 import engine as E
    try:
        eng = E.Engine(1, framing=framing, chunk_samples=BLOCK_SYMS * 16)
    except E.WbError as e:
        sys.stderr.write("%s\n" % e)
        sys.exit(1)
    packets = errors = 0
    # a pipe hands over whatever has arrived (read1), so a packet is decoded as soon as its last symbol is there, as in
    # the reference's per-symbol loop; a regular file comes in full blocks.  A split float waits for its other half.
    read = getattr(fin, "read1", fin.read)
    left = b""
    while True:
        raw = read(4 * min(BLOCK_SYMS, eng.sd_cap))
        if not raw:
            break
        raw = left + raw
        left = raw[len(raw) - len(raw) % 4:]
        raw = raw[:len(raw) - len(left)]
        if not raw:
            continue
        eng.process_soft([np.frombuffer(raw, dtype=np.float32)])
        eng.sync()
        for cw in eng.drain_codewords():
            packets = (packets + 1) & 0xFFFF                     # uint16_t counters, src/drs232_ldpc.c:116
            if not cw["crc_ok"]:
                errors = (errors + 1) & 0xFFFF
                if verbose == 2:                                 # src/drs232_ldpc.c:246-251
                    from wenet_b200.siggen import crc16_ccitt_false
                    body = bytes(bytearray(cw["bytes"]))
                    sys.stderr.write("tx_checksum: 0x%02x rx_checksum: 0x%02x\n"
                                     % (body[256] + (body[257] << 8), crc16_ccitt_false(body[:256])))
            if verbose:
                sys.stderr.write("packets: %d packet_errors: %d PER: %s iter: %d\n" % (packets, errors, _per(errors, packets), cw["iters"]))
        pk = eng.drain_packets(0)
        if pk:
            fout.write(pk)
            fout.flush()
    fout.flush()
    sys.stderr.write("packets: %d packet_errors: %d PER: %s\n" % (packets, errors, _per(errors, packets)))
    eng.close()
    return 0
