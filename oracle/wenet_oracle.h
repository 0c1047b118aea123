/*
 * wenet_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, single-threaded) of the reference hot
 * path  fsk_demod | drs232_ldpc  /  fsk_demod | wenet_ldpc  of
 * projecthorus/wenet.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product
 * (wenet_b200/, libwenet_b200.so) never does.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against
 * the unmodified reference compiled from source (oracle/_ref, see
 * oracle/Makefile) by tests/test_oracle_vs_ref.py, and against the reference's
 * own LDPC known-answer vector (src/H2064_516_sparse.h:27-33, committed as
 * tests/golden/ldpc_kat.npz).
 */
#ifndef WENET_ORACLE_H
#define WENET_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/phi0.c:13 */
float wo_phi0(float x);
void wo_phi0_array(const float *x, float *y, long n);

/* reference src/mpdecode_core.c:569 (sd_to_llr) */
void wo_sd_to_llr(float *llr, const double *sd, int n);

/* reference src/mpdecode_core.c:494 (run_ldpc_decoder) = :152 + :385.
   *pcc is left untouched on the all-zero-data exit, as in the reference. */
int wo_ldpc_decode(const float *llr, uint8_t *bits, int max_iter, int *pcc);

/* reference src/mpdecode_core.c:72 (encode) */
void wo_ldpc_encode(const uint8_t *ibits, uint8_t *pbits);

/* reference src/drs232_ldpc.c:91 (gen_crc16) */
uint16_t wo_crc16(const uint8_t *data, int n);

/* ---- deframer + decoder, reference src/drs232_ldpc.c:176-275 (mode 1) and
 *      src/wenet_ldpc.c:171-259 (mode 2) ---- */
typedef struct wo_deframer wo_deframer;
wo_deframer *wo_deframer_create(int mode, int max_iter);
void wo_deframer_destroy(wo_deframer *d);
/* feed n soft symbols; CRC-valid 256-byte packets are appended to out
 * (capacity out_cap bytes).  Optional taps per codeword (any may be NULL):
 *   tap_llr    [cw][2580] float, tap_iters [cw] int, tap_pcc [cw] int,
 *   tap_crc_ok [cw] uint8, tap_pos [cw] long (index, in the whole fed stream,
 *   of the first collected symbol), tap_bytes [cw][258] all decoded bytes.
 * tap_cap = capacity in codewords.  Returns bytes written to out;
 * *n_cw = codewords decoded in this call. */
long wo_deframer_feed(wo_deframer *d, const float *sd, long n,
                      uint8_t *out, long out_cap,
                      float *tap_llr, int *tap_iters, int *tap_pcc,
                      uint8_t *tap_crc_ok, long *tap_pos, uint8_t *tap_bytes,
                      long tap_cap, long *n_cw);
/* packets / packet_errors counters as uint16_t like the reference */
void wo_deframer_counts(wo_deframer *d, int *packets, int *packet_errors);

/* ---- FSK demodulator, reference src/fsk.c:128 (fsk_create_hbr),
 *      :540 (fsk_demod_freq_est), :679 (fsk2_demod) and the frame loop of
 *      src/fsk_demod.c:270-299 ---- */
typedef struct wo_fsk wo_fsk;
wo_fsk *wo_fsk_create(int Fs, int Rs, int P, int M);
void wo_fsk_destroy(wo_fsk *f);
void wo_fsk_set_est_limits(wo_fsk *f, int lo, int hi);
int wo_fsk_nin(wo_fsk *f);
int wo_fsk_nbits(wo_fsk *f);
/* one frame: in = nin interleaved float pairs; sd (Nbits floats) and/or bits
 * (Nbits bytes) may be NULL */
void wo_fsk_demod(wo_fsk *f, float *sd, uint8_t *bits, const float *in);
/* same layout as ref_fsk_state(): 19 floats */
void wo_fsk_state(wo_fsk *f, float *out);
void wo_fsk_fft_est(wo_fsk *f, float *out);
void wo_fsk_eye(wo_fsk *f, int *neyetr, int *neyesamp, float *out);
/* whole-stream frame loop, same contract as ref_fsk_run() in ref_harness.c:
 * fmt 0 = cf32, 1 = cu8, 2 = cs16, 3 = real s16; frame_log = 8 floats/frame
 * {nin, f_est0..3, norm_rx_timing, ppm, EbNodB} */
long wo_fsk_run(wo_fsk *f, int fmt, const void *raw, long nsamp,
                float *sd_out, long sd_cap, long *n_sd,
                float *frame_log, long log_cap, long *consumed);
/* the same loop through the hard-decision output of fsk_demod() (src/fsk.c:936-959): one byte per bit */
long wo_fsk_run_bits(wo_fsk *f, int fmt, const void *raw, long nsamp, uint8_t *bits_out, long cap, long *n_bits);

/* ---- transmit side ---- */
/* one on-air frame (preamble + unique word + body) as 0/1 bytes in transmit order: reference tx/PacketTX.py:123-137,
 * tx/radio_wrappers.py:385-417 / :502-559; mode 1 = RS232 (3430 bits), 2 = scrambled v2 (2744 bits) */
int wo_tx_frame_bits(const uint8_t *payload, int payload_len, int mode, uint8_t *bits);
/* reference src/fsk.c:1162-1204 (fsk_mod_c); wo_fsk_set_tx = the tx_f1 / tx_fs arguments of fsk_create_hbr */
void wo_fsk_set_tx(wo_fsk *f, int f1_tx, int fs_tx);
void wo_fsk_mod_c(wo_fsk *f, float *out, const uint8_t *tx_bits);

#ifdef __cplusplus
}
#endif
#endif
