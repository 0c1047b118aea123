#!/usr/bin/env python3
"""drs232_ldpc InputOneSymbolPerFloat OutputPackets [-v[v]] -- Wenet v1 (RS232-framed) deframer + LDPC decoder + CRC
gate on the GPU; same argv and byte streams as reference src/drs232_ldpc.c:142-169."""
import sys

from wenet_b200.cli._ldpc_cli import main

if __name__ == "__main__":
    sys.exit(main(sys.argv, "v1", "drs232"))
