/*
 * wenet_b200.h -- C ABI of libwenet_b200.so: a batched, B200-resident
 * replacement for the two C programs in the middle of the Wenet receive pipe
 *
 *     rtl_sdr | fsk_demod | drs232_ldpc (or wenet_ldpc) | rx_ssdv.py
 *
 * One engine = one GPU = n_streams independent IQ streams, each with its own
 * demodulator + deframer state resident in HBM.  Plain pointers and sizes
 * only; every function returns 0 or a negative WB_E* code and never aborts;
 * wb_last_error() gives the message for the calling thread.  Calls on one
 * engine must be serialised by the caller (like one `struct FSK`), different
 * engines are independent.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the projecthorus/wenet tree).
 */
#ifndef WENET_B200_H
#define WENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_ABI_VERSION 1

/* input sample formats -- src/fsk_demod.c:273-296 (plus raw float pairs) */
#define WB_FMT_CF32 0   /* interleaved float32 I/Q, used as is                       */
#define WB_FMT_CU8  1   /* --cu8 : (u8 - 127) / 128                                  */
#define WB_FMT_CS16 2   /* --cs16: s16 / 1000                                        */
#define WB_FMT_S16  3   /* real s16 / 1000, imag = 0 (fsk_demod without -c/-d)       */

/* framing after the demodulator */
#define WB_FRAMING_NONE 0   /* demodulate only (fsk_demod alone)                     */
#define WB_FRAMING_V1   1   /* src/drs232_ldpc.c: 40-bit UW, RS232 start/stop bits   */
#define WB_FRAMING_V2   2   /* src/wenet_ldpc.c : 32-bit UW, +-1 descramble          */

#define WB_FLAG_KEEP_LLR   1u  /* keep per-codeword LLRs for wb_drain_codewords (test tap) */
#define WB_FLAG_STATS      2u  /* compute the Eb/N0 + eye-diagram statistics every frame   */
#define WB_FLAG_HARD_BITS  4u  /* also keep the demodulator's hard bits for wb_drain_hard      */

#define WB_OK        0
#define WB_EINVAL   -1   /* bad argument / unsupported configuration */
#define WB_ENOMEM   -2
#define WB_ECUDA    -3   /* CUDA runtime error (see wb_last_error)    */
#define WB_ENODEV   -4   /* no CUDA device: the engine has NO CPU fallback */
#define WB_ERANGE   -5   /* buffer too small / capacity exceeded      */

#define WB_PACKET_BYTES 256
#define WB_CODE_BITS    2580

typedef struct wb_engine wb_engine;

typedef struct wb_config {
    uint32_t struct_size;     /* = sizeof(wb_config) */
    int32_t  device;          /* CUDA ordinal */
    int32_t  n_streams;
    /* modem -- arguments of fsk_create_hbr(Fs, Rs, P, M, ..) src/fsk.h:110; P = 0 -> Fs/Rs like
       src/fsk_demod.c:186-188 */
    int32_t  Fs, Rs, M, P;
    /* fsk_set_est_limits(lo, hi) src/fsk.h:160 / fsk_demod -b -u; <= 0 keeps the defaults */
    int32_t  est_lo, est_hi;
    int32_t  in_fmt;          /* WB_FMT_*     */
    int32_t  framing;         /* WB_FRAMING_* */
    int32_t  ldpc_max_iter;   /* struct LDPC.max_iter (src/mpdecode_core.h:19); 0 -> MAX_ITER = 10 */
    uint32_t flags;           /* WB_FLAG_*    */
    uint64_t chunk_samples;   /* capacity, per stream, of the resident input buffer */
} wb_config;

/* per-stream modem statistics: the fields of the stderr JSON line of src/fsk_demod.c:345-401
   (struct MODEM_STATS, src/modem_stats.h:46-71) */
typedef struct wb_stats {
    float   EbNodB;           /* stats.snr_est */
    float   ppm;
    float   f_est[4];
    float   rx_timing, foff, norm_rx_timing;
    int32_t nin;
    int32_t neyetr, neyesamp;
    float   rx_eye[8][160];
    int32_t nfft;             /* Ndft / 2 */
    float   samp_fft[512];    /* fsk->fft_est */
    uint64_t frames;          /* modem frames demodulated so far */
    uint32_t packets, packet_errors;  /* drs232_ldpc.c:116 counters (uint16_t there; not wrapped here) */
} wb_stats;

/* one decoded codeword, as wb_drain_codewords returns it */
typedef struct wb_codeword {
    int32_t  stream;
    uint32_t seq;             /* per-stream codeword number, decode order */
    int32_t  iters;           /* return value of run_ldpc_decoder */
    int32_t  parity_ok;       /* parityCheckCount, -1 if the decoder never assigned it */
    int32_t  crc_ok;
    uint8_t  bytes[258];      /* payload + CRC as packed by drs232_ldpc.c:234-239 */
    uint8_t  pad[2];
} wb_codeword;

/* ---- lifecycle -------------------------------------------------------- */
/* fsk_create_hbr() + the struct LDPC set-up of drs232_ldpc.c:128-138, for n_streams streams */
int  wb_create(const wb_config *cfg, wb_engine **out);
/* fsk_destroy() */
void wb_destroy(wb_engine *e);
const char *wb_last_error(void);
int  wb_abi_version(void);

/* ---- streaming path (host buffers in, host buffers out) ---------------- */
/* fsk_nin(): samples the NEXT frame of each stream will consume; nin[n_streams] */
int  wb_nin(wb_engine *e, uint32_t *nin);
/* the fread() of fsk_demod.c:270: append nsamp[s] samples of cfg.in_fmt from iq[s] to stream s.
   iq[s] may be NULL when nsamp[s] == 0. */
int  wb_feed(wb_engine *e, const void *const *iq, const uint64_t *nsamp);
/* same, all streams equal length from one strided host block (stream s at base + s*stride_bytes) */
int  wb_feed_strided(wb_engine *e, const void *base, uint64_t stride_bytes, uint64_t nsamp);
/* run the whole path on everything resident: per stream, fsk_demod_sd() for every complete frame
   (fsk_demod.c:299), then the UW search / collect / sd_to_llr / run_ldpc_decoder / CRC gate of
   drs232_ldpc.c:176-259.  Asynchronous; the drain calls and wb_sync wait for it. */
int  wb_process(wb_engine *e);
/* the same without a demodulator in front: the fread(&symbol) loop of drs232_ldpc.c:176 / wenet_ldpc.c:171.
   sd[s] = nsym[s] float32 soft symbols (negative = 1) of stream s, at most sd_cap (wb_geometry) per call */
int  wb_process_soft(wb_engine *e, const float *const *sd, const uint64_t *nsym);
int  wb_sync(wb_engine *e);
/* the fwrite(packet) of drs232_ldpc.c:254: CRC-valid 256-byte payloads of `stream`, decode order,
   produced since the last drain of that stream */
int  wb_drain_packets(wb_engine *e, int stream, uint8_t *buf, size_t cap, size_t *nbytes);
/* every stream at once: records of {int32 stream, uint32 seq, 256 bytes}, sorted by (stream, seq); seq is the
   per-stream codeword number the payload came from (= wb_codeword.seq: it keeps counting across calls, and CRC
   failures leave gaps), so (stream, seq) is unique for the life of the engine.  The per-stream queues behind
   wb_drain_packets / wb_drain_all_packets grow until drained. */
int  wb_drain_all_packets(wb_engine *e, uint8_t *buf, size_t cap, size_t *nbytes, uint64_t *npackets);
/* the fwrite(sdbuf) of fsk_demod.c:403: soft decisions of `stream` produced by the LAST wb_process */
int  wb_drain_soft(wb_engine *e, int stream, float *buf, size_t cap_floats, size_t *n);
/* the fwrite(bitbuf) of fsk_demod.c:405 (fsk_demod without -s): rx_bits of fsk_demod() (src/fsk.h:183,
   src/fsk.c:936-959), one byte per bit, of `stream` from the LAST wb_process.  These are the arg-max tone
   decisions, which for 4-FSK are not the signs of the soft decisions.  Needs WB_FLAG_HARD_BITS. */
int  wb_drain_hard(wb_engine *e, int stream, uint8_t *buf, size_t cap, size_t *n);
/* test tap: all codewords decoded by the LAST wb_process (CRC-valid or not), sorted by (stream, seq);
   llr (may be NULL; needs WB_FLAG_KEEP_LLR) receives 2580 floats per codeword */
int  wb_drain_codewords(wb_engine *e, wb_codeword *cw, float *llr, size_t cap, size_t *n);
/* fsk_get_demod_stats() + the JSON fields of fsk_demod.c:351-392 */
int  wb_get_stats(wb_engine *e, int stream, wb_stats *st);
/* fsk_clear_estimators() src/fsk.h:166 for every stream */
int  wb_clear_estimators(wb_engine *e);

/* ---- stage-level entry points ------------------------------------------ */
/* run_ldpc_decoder() (src/mpdecode_core.h:37) over n codewords of 2580 LLRs each (host memory):
   bits_packed = 323 bytes per codeword (2580 bits MSB first, 4 pad bits), iters / parity_ok per codeword */
int  wb_ldpc_decode_batch(wb_engine *e, const float *llr, size_t n, int max_iter,
                          uint8_t *bits_packed, int32_t *iters, int32_t *parity_ok);
/* sd_to_llr() (src/mpdecode_core.h:39) over n blocks of 2580 float soft decisions */
int  wb_sd_to_llr_batch(wb_engine *e, const float *sd, size_t n, float *llr);

/* ---- transmit side on the device (synthetic input in HBM) ---------------- */
/* What the receive path decodes, built where wb_feed would have put it: for every stream n_packets frames
   (reference tx/PacketTX.py:123-137 frame_packet: payload + CRC16 + RA-LDPC parity of tx/ldpc_encoder.py /
   src/mpdecode_core.c:72-91, preamble + unique word, v1 UART framing or v2 scrambling per the engine's framing),
   idle '1' bits around them, modulated by the reference's fsk_mod_c (src/fsk.c:1162-1204; without noise the samples
   are bit-identical to it, amplitude 2) and, if ebno_db is finite, AWGN + peak normalisation as
   benchmarking/generate_lowsnr.py:70-89 (unit-amplitude signal, max |y| = 1).  With framing NONE the payload bytes
   are sent as they are, MSB first (M = 4: two bits per symbol).  The bit stream is padded with idle bits to whole
   modulator calls of 48 symbols.  Equivalent to a wb_feed of *nsamp_per_stream samples to every stream. */
typedef struct wb_tx_config {
    uint32_t struct_size;      /* sizeof(wb_tx_config) */
    int32_t  n_packets;        /* frames (framing NONE: 256-byte blocks) per stream */
    int32_t  lead_in_bits;     /* idle bits before the first frame */
    int32_t  gap_bits;         /* idle bits after every frame */
    int32_t  tail_bits;        /* idle bits at the end */
    int32_t  f1_tx, fs_tx;     /* Hz: tone of symbol 0 and tone spacing (tx_f1 / tx_fs of fsk_create_hbr, src/fsk.h:110) */
    float    ebno_db;          /* AWGN at this Eb/N0; NaN = none (the reference modulator's output as it is) */
    uint64_t seed;             /* noise generator seed */
    const float *ebno_db_per_stream;   /* host, [n_streams]: overrides ebno_db stream by stream (NaN entries = no noise
                                          is not supported here: all streams are noisy or none); NULL = ebno_db for all */
} wb_tx_config;
/* payloads: host memory, [n_streams][n_packets][256] bytes */
int  wb_tx_synthesize(wb_engine *e, const uint8_t *payloads, const wb_tx_config *cfg, uint64_t *nsamp_per_stream);
/* test tap: the on-air bits (one byte each) of a stream built by the last wb_tx_synthesize */
int  wb_tx_read_bits(wb_engine *e, int stream, uint8_t *bits, size_t cap, size_t *n);

/* ---- HBM-resident benchmarking helpers --------------------------------- */
/* device address / stride / capacity of the resident input buffer */
int  wb_dev_input(wb_engine *e, void **dptr, uint64_t *stride_bytes, uint64_t *capacity_samples);
/* test tap: samples [first, first + nsamp) of a stream's resident input row (counted from the headroom mark, where a
   fresh engine's first wb_feed / wb_tx_synthesize lands), in the engine's input format */
int  wb_dev_read_input(wb_engine *e, int stream, uint64_t first, uint64_t nsamp, void *out);
/* declare nsamp samples resident in every stream and rewind the read positions to 0 */
int  wb_dev_set_fill(wb_engine *e, uint64_t nsamp);
/* replicate stream 0..n_src-1 of the resident input into all streams, stream s = source (s % n_src)
   rotated by (s / n_src) * rot samples */
int  wb_dev_replicate(wb_engine *e, int n_src, uint64_t nsamp, uint64_t rot);
/* resident LDPC benchmark: upload n_src codewords of LLRs, replicate to n, decode in place */
int  wb_dev_ldpc_setup(wb_engine *e, const float *llr, size_t n_src, size_t n);
int  wb_dev_ldpc_run(wb_engine *e, int max_iter);
int  wb_dev_ldpc_result(wb_engine *e, size_t first, size_t n, uint8_t *bits_packed, int32_t *iters, int32_t *parity_ok);
/* CUDA-event timing on the engine's stream */
int  wb_timer_start(wb_engine *e);
int  wb_timer_stop(wb_engine *e, float *ms);
/* per-kernel event timing of the last wb_process: ms[0..3] = fsk, deframe, sd_to_llr, ldpc */
int  wb_last_kernel_ms(wb_engine *e, float *ms);
/* number of kernels this engine has launched so far */
uint64_t wb_launch_count(wb_engine *e);
/* codewords found by the last wb_process */
uint64_t wb_last_codewords(wb_engine *e);
/* samples consumed (all streams) by the last wb_process */
uint64_t wb_last_samples(wb_engine *e);


/* ---- host-side helpers --------------------------------------------------- */
/* pinned host memory for staging buffers handed to wb_feed / wb_feed_strided */
void *wb_host_alloc(size_t bytes);
void  wb_host_free(void *p);
/* the pinned-host <-> device copy leg alone (what bounds the host-buffer path end to end): `bytes` of pinned and of
   device memory on `device`; run = `iters` flat copies back to back, *ms = their CUDA-event duration.  Run on all GPUs of a
   box at once it gives the box's concurrent-copy ceiling (tools/micro/pcie_multi.cu is the stand-alone form). */
typedef struct wb_copy_probe wb_copy_probe;
int  wb_copy_probe_create(int device, size_t bytes, wb_copy_probe **out);
int  wb_copy_probe_run(wb_copy_probe *p, int iters, int d2h, float *ms);
void wb_copy_probe_destroy(wb_copy_probe *p);
/* derived geometry: out[0..13] = N, Nbits, Ts, P, Ndft, nmax, job_cap, sd_cap, Nsym, M, ldpc_max_iter,
   symbols collected per packet, streams per CTA and shared-memory bytes per CTA of the demodulator kernel */
int  wb_geometry(wb_engine *e, int32_t *out, int n);
/* test tap: log {nin, f_est bins[4], norm_rx_timing, ppm, rx_timing} of the first frames_per_stream frames of
   each following wb_process (what oracle/ref_harness.c logs per frame of src/fsk_demod.c:270-299) */
int  wb_enable_frame_log(wb_engine *e, int frames_per_stream);
int  wb_read_frame_log(wb_engine *e, int stream, float *buf, size_t cap_frames);

#ifdef __cplusplus
}
#endif
#endif /* WENET_B200_H */
