#!/usr/bin/env python3
"""wenet_ldpc InputOneSymbolPerFloat OutputPackets [-v[v]] -- Wenet v2 (scrambled, 32-bit unique word) deframer +
LDPC decoder + CRC gate on the GPU; same argv and byte streams as reference src/wenet_ldpc.c."""
import sys

from wenet_b200.cli._ldpc_cli import main

if __name__ == "__main__":
    sys.exit(main(sys.argv, "v2", "wenet_ldpc"))
