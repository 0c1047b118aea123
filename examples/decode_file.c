/*
 * decode_file.c -- a C caller of include/wenet_b200.h: what a C program that today links src/fsk.c +
 * src/mpdecode_core.c (fsk_create_hbr / fsk_nin / fsk_demod_sd, then the drs232_ldpc symbol loop) looks like
 * against libwenet_b200.so.  One stream here; n_streams is where the GPU earns its keep.
 *
 *   gcc -Iinclude examples/decode_file.c -Lwenet_b200 -lwenet_b200 -Wl,-rpath,$PWD/wenet_b200 -o decode_file
 *   ./decode_file iq.cu8 packets.bin          (cu8 IQ at 921416 samples/s, 115177 baud, RS232 framing)
 *
 * Exit codes: 0 ok, 1 usage / file error, 2 engine error (e.g. no CUDA device: there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wenet_b200.h"

#define BLOCK_SAMPLES (64 * 384)

static int fail(const char *what)
{
    fprintf(stderr, "%s: %s\n", what, wb_last_error());
    return 2;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s InputCu8IQ OutputPackets\n", argv[0]);
        return 1;
    }
    FILE *fin = strcmp(argv[1], "-") ? fopen(argv[1], "rb") : stdin;
    FILE *fout = strcmp(argv[2], "-") ? fopen(argv[2], "wb") : stdout;
    if (!fin || !fout) {
        fprintf(stderr, "Couldn't open files\n");
        return 1;
    }

    wb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.n_streams = 1;
    cfg.Fs = 921416; cfg.Rs = 115177; cfg.M = 2;       /* fsk_create_hbr(Fs, Rs, P = Fs/Rs, M, ..) */
    cfg.in_fmt = WB_FMT_CU8;                             /* fsk_demod --cu8 */
    cfg.framing = WB_FRAMING_V1;                         /* | drs232_ldpc */
    cfg.chunk_samples = BLOCK_SAMPLES + 1024;
    wb_engine *e = NULL;
    if (wb_create(&cfg, &e) != WB_OK) return fail("wb_create");

    static uint8_t iq[2 * BLOCK_SAMPLES], pk[WB_PACKET_BYTES * 64];
    unsigned long packets = 0;
    size_t got;
    while ((got = fread(iq, 2, BLOCK_SAMPLES, fin)) > 0) {
        const void *bufs[1] = {iq};
        uint64_t ns[1] = {got};
        size_t nbytes = 0;
        if (wb_feed(e, bufs, ns) != WB_OK) return fail("wb_feed");
        if (wb_process(e) != WB_OK) return fail("wb_process");
        do {                                              /* the fwrite(packet) of drs232_ldpc.c:254 */
            if (wb_drain_packets(e, 0, pk, sizeof(pk), &nbytes) != WB_OK) return fail("wb_drain_packets");
            fwrite(pk, 1, nbytes, fout);
            packets += nbytes / WB_PACKET_BYTES;
        } while (nbytes == sizeof(pk));
        fflush(fout);
    }
    wb_stats st;
    if (wb_get_stats(e, 0, &st) == WB_OK)
        fprintf(stderr, "packets: %lu  frames: %llu  f1_est %.1f f2_est %.1f ppm %.0f\n", packets,
                (unsigned long long)st.frames, st.f_est[0], st.f_est[1], st.ppm);
    wb_destroy(e);
    fclose(fout);
    return 0;
}
