"""argv-compatible stand-ins for the reference's receive-pipe programs (src/fsk_demod.c, src/drs232_ldpc.c,
src/wenet_ldpc.c), each a one-stream engine on the GPU behind the same stdin/stdout/stderr byte surface."""
