/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin ctypes-friendly harness of OUR OWN around the UNMODIFIED reference
 * objects (fsk.o kiss_fft.o mpdecode_core.o phi0.o compiled from
 * /root/reference/src by oracle/Makefile).  It exposes stage-level taps that
 * the reference only offers inside its main() functions:
 *
 *   - the frame loop of src/fsk_demod.c:270-299 (sample conversion + one
 *     fsk_demod_sd() call per nin samples) over a whole in-memory stream,
 *     logging nin, f_est, norm_rx_timing, ppm and EbNodB per frame;
 *   - sd_to_llr / run_ldpc_decoder (src/mpdecode_core.h:35-39) with a
 *     settable max_iter;
 *   - phi0 (src/phi0.c:13), encode (src/mpdecode_core.c:72);
 *   - the code tables and the known-answer vector of
 *     src/H2064_516_sparse.h and the scramble code of src/wenet_scramble.h
 *     (read through pointers so tools/gen_tables.py can extract them).
 *
 * The deframers live inside main() of drs232_ldpc.c / wenet_ldpc.c and are
 * therefore exercised through the compiled binaries (oracle/_ref/drs232_ldpc,
 * oracle/_ref/wenet_ldpc) over pipes, not through this harness.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "fsk.h"
#include "modem_stats.h"
#include "codec2_fdmdv.h"
#include "mpdecode_core.h"
#include "phi0.h"
#include "H2064_516_sparse.h"
#include "wenet_scramble.h"

/* ---- tables ------------------------------------------------------------ */

const uint16_t *ref_H_rows(void) { return H_rows; }
const uint16_t *ref_H_cols(void) { return H_cols; }
int ref_H_rows_len(void) { return (int)(sizeof(H_rows) / sizeof(H_rows[0])); }
int ref_H_cols_len(void) { return (int)(sizeof(H_cols) / sizeof(H_cols[0])); }
const float *ref_kat_input(void) { return input; }
int ref_kat_input_len(void) { return (int)(sizeof(input) / sizeof(input[0])); }
const char *ref_kat_detected(void) { return (const char *)detected_data; }
int ref_kat_detected_len(void) { return (int)(sizeof(detected_data) / sizeof(detected_data[0])); }
int ref_kat_detected_elsize(void) { return (int)sizeof(detected_data[0]); }
int ref_kat_input_elsize(void) { return (int)sizeof(input[0]); }
const double *ref_scramble_code(void) { return scramble_code; }
int ref_scramble_len(void) { return (int)(sizeof(scramble_code) / sizeof(scramble_code[0])); }
int ref_max_iter(void) { return MAX_ITER; }

/* ---- LDPC -------------------------------------------------------------- */

static void fill_ldpc(struct LDPC *l, int max_iter)
{
    l->max_iter = max_iter;
    l->dec_type = 0;
    l->q_scale_factor = 1;
    l->r_scale_factor = 1;
    l->CodeLength = CODELENGTH;
    l->NumberParityBits = NUMBERPARITYBITS;
    l->NumberRowsHcols = NUMBERROWSHCOLS;
    l->max_row_weight = MAX_ROW_WEIGHT;
    l->max_col_weight = MAX_COL_WEIGHT;
    l->H_rows = H_rows;
    l->H_cols = H_cols;
}

/* returns iterations; *pcc is pre-set to pcc_init because the reference leaves
   it untouched on the all-zero-data early exit (src/mpdecode_core.c:467-476) */
int ref_ldpc_decode(const float *llr, uint8_t *out_bits, int max_iter, int *pcc)
{
    struct LDPC l;
    float tmp[CODELENGTH];
    fill_ldpc(&l, max_iter);
    memcpy(tmp, llr, sizeof(tmp));
    return run_ldpc_decoder(&l, out_bits, tmp, pcc);
}

void ref_sd_to_llr(float *llr, const double *sd, int n)
{
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(tmp, sd, sizeof(double) * (size_t)n);
    sd_to_llr(llr, tmp, n);
    free(tmp);
}

float ref_phi0(float x) { return phi0(x); }

void ref_phi0_array(const float *x, float *y, long n)
{
    long i;
    for (i = 0; i < n; i++) y[i] = phi0(x[i]);
}

void ref_encode(const uint8_t *ibits, uint8_t *pbits)
{
    struct LDPC l;
    uint8_t tmp[CODELENGTH];
    fill_ldpc(&l, MAX_ITER);
    memcpy(tmp, ibits, CODELENGTH - NUMBERPARITYBITS);
    encode(&l, tmp, pbits);
}

/* ---- FSK --------------------------------------------------------------- */

void *ref_fsk_create(int Fs, int Rs, int P, int M) { return fsk_create_hbr(Fs, Rs, P, M, 1200, 400); }
void ref_fsk_destroy(void *h) { fsk_destroy((struct FSK *)h); }
void ref_fsk_set_est_limits(void *h, int lo, int hi) { fsk_set_est_limits((struct FSK *)h, lo, hi); }
int ref_fsk_nin(void *h) { return (int)fsk_nin((struct FSK *)h); }
int ref_fsk_nbits(void *h) { return ((struct FSK *)h)->Nbits; }
int ref_fsk_N(void *h) { return ((struct FSK *)h)->N; }
int ref_fsk_Ndft(void *h) { return ((struct FSK *)h)->Ndft; }
int ref_fsk_nstash(void *h) { return ((struct FSK *)h)->nstash; }

/* one frame straight through fsk_demod_sd(): in = nin interleaved float pairs */
void ref_fsk_demod_sd(void *h, float *sd, const float *in)
{
    fsk_demod_sd((struct FSK *)h, sd, (COMP *)in);
}

void ref_fsk_demod_bits(void *h, uint8_t *bits, const float *in)
{
    fsk_demod((struct FSK *)h, bits, (COMP *)in);
}

/* state tap: phi_c[4] (8 floats), f_est[4], norm_rx_timing, ppm, EbNodB,
   snr_est, rx_timing, foff, clock_offset -> 19 floats */
void ref_fsk_state(void *h, float *out)
{
    struct FSK *f = (struct FSK *)h;
    int i;
    for (i = 0; i < 4; i++) { out[2 * i] = f->phi_c[i].real; out[2 * i + 1] = f->phi_c[i].imag; }
    for (i = 0; i < 4; i++) out[8 + i] = f->f_est[i];
    out[12] = f->norm_rx_timing;
    out[13] = f->ppm;
    out[14] = f->EbNodB;
    out[15] = f->stats->snr_est;
    out[16] = f->stats->rx_timing;
    out[17] = f->stats->foff;
    out[18] = f->stats->clock_offset;
}

void ref_fsk_fft_est(void *h, float *out)
{
    struct FSK *f = (struct FSK *)h;
    memcpy(out, f->fft_est, sizeof(float) * (size_t)(f->Ndft / 2));
}

void ref_fsk_samp_old(void *h, float *out)
{
    struct FSK *f = (struct FSK *)h;
    memcpy(out, f->samp_old, sizeof(COMP) * (size_t)f->nstash);
}

/* eye diagram tap: neyetr, neyesamp and rx_eye[neyetr][neyesamp] flattened */
void ref_fsk_eye(void *h, int *neyetr, int *neyesamp, float *out)
{
    struct FSK *f = (struct FSK *)h;
    struct MODEM_STATS st;
    int i, j;
    fsk_get_demod_stats(f, &st);
    *neyetr = st.neyetr;
    *neyesamp = st.neyesamp;
    for (i = 0; i < st.neyetr; i++)
        for (j = 0; j < st.neyesamp; j++)
            out[i * st.neyesamp + j] = st.rx_eye[i][j];
}

/* sample conversion of one frame, the arithmetic of src/fsk_demod.c:273-296 */
static void load_frame(COMP *modbuf, int fmt, const void *raw, long pos, int nin)
{
    int i;
    if (fmt == 0) {
        const float *p = (const float *)raw + 2 * pos;
        for (i = 0; i < nin; i++) { modbuf[i].real = p[2 * i]; modbuf[i].imag = p[2 * i + 1]; }
    } else if (fmt == 1) {
        const uint8_t *p = (const uint8_t *)raw + 2 * pos;
        for (i = 0; i < nin; i++) {
            modbuf[i].real = ((float)p[2 * i] - 127.0) / 128.0;
            modbuf[i].imag = ((float)p[2 * i + 1] - 127.0) / 128.0;
        }
    } else if (fmt == 2) {
        const int16_t *p = (const int16_t *)raw + 2 * pos;
        for (i = 0; i < nin; i++) {
            modbuf[i].real = ((float)p[2 * i]) / FDMDV_SCALE;
            modbuf[i].imag = ((float)p[2 * i + 1] / FDMDV_SCALE);
        }
    } else {
        const int16_t *p = (const int16_t *)raw + pos;
        for (i = 0; i < nin; i++) {
            modbuf[i].real = ((float)p[i]) / FDMDV_SCALE;
            modbuf[i].imag = 0.0;
        }
    }
}

/*
 * The frame loop of src/fsk_demod.c:270-299 over an in-memory stream.
 *   fmt: 0 = cf32 (test tap: samples are used as they are), 1 = cu8,
 *        2 = cs16, 3 = real s16  (src/fsk_demod.c:273-296)
 *   raw / nsamp: the stream, nsamp in (complex) samples
 *   sd_out: capacity sd_cap floats;  frame_log: 8 floats per frame
 *           {nin_used, f_est0..3, norm_rx_timing, ppm, EbNodB}
 * Returns the number of frames processed; *n_sd = soft decisions written;
 * *consumed = samples consumed.
 */
long ref_fsk_run(void *h, int fmt, const void *raw, long nsamp,
                 float *sd_out, long sd_cap, long *n_sd,
                 float *frame_log, long log_cap, long *consumed)
{
    struct FSK *f = (struct FSK *)h;
    long pos = 0, frames = 0, nsd = 0;
    int nmax = f->N + 2 * f->Ts;
    COMP *modbuf = (COMP *)malloc(sizeof(COMP) * (size_t)nmax);
    float *sdbuf = (float *)malloc(sizeof(float) * (size_t)f->Nbits);
    memset(sdbuf, 0, sizeof(float) * (size_t)f->Nbits);
    while (pos + (long)fsk_nin(f) <= nsamp) {
        int nin = (int)fsk_nin(f);
        load_frame(modbuf, fmt, raw, pos, nin);
        fsk_demod_sd(f, sdbuf, modbuf);
        pos += nin;
        if (nsd + f->Nbits <= sd_cap) {
            memcpy(sd_out + nsd, sdbuf, sizeof(float) * (size_t)f->Nbits);
            nsd += f->Nbits;
        }
        if (frame_log && frames < log_cap) {
            float *l = frame_log + 8 * frames;
            l[0] = (float)nin;
            l[1] = f->f_est[0]; l[2] = f->f_est[1]; l[3] = f->f_est[2]; l[4] = f->f_est[3];
            l[5] = f->norm_rx_timing; l[6] = f->ppm; l[7] = f->EbNodB;
        }
        frames++;
    }
    free(modbuf);
    free(sdbuf);
    *n_sd = nsd;
    *consumed = pos;
    return frames;
}

/* The same loop through fsk_demod() (fsk_demod without -s, src/fsk_demod.c:301,405): hard bits, one byte each */
long ref_fsk_run_bits(void *h, int fmt, const void *raw, long nsamp, uint8_t *bits_out, long cap, long *n_bits)
{
    struct FSK *f = (struct FSK *)h;
    long pos = 0, frames = 0, nb = 0;
    COMP *modbuf = (COMP *)malloc(sizeof(COMP) * (size_t)(f->N + 2 * f->Ts));
    uint8_t *bitbuf = (uint8_t *)calloc((size_t)f->Nbits, 1);
    while (pos + (long)fsk_nin(f) <= nsamp) {
        int nin = (int)fsk_nin(f);
        load_frame(modbuf, fmt, raw, pos, nin);
        fsk_demod(f, bitbuf, modbuf);
        pos += nin;
        if (nb + f->Nbits <= cap) { memcpy(bits_out + nb, bitbuf, (size_t)f->Nbits); nb += f->Nbits; }
        frames++;
    }
    free(modbuf);
    free(bitbuf);
    *n_bits = nb;
    return frames;
}

/* ---- transmit side: the reference's own modulator (src/fsk.c:1162-1204) ---- */
void ref_fsk_set_tx(void *h, int f1_tx, int fs_tx)
{
    struct FSK *f = (struct FSK *)h;
    f->f1_tx = f1_tx; f->fs_tx = fs_tx;          /* what fsk_create_hbr stores from its tx_f1 / tx_fs arguments, src/fsk.c:161-162 */
    f->tx_phase_c.real = cosf(0); f->tx_phase_c.imag = sinf(0);     /* comp_exp_j(0), src/fsk.c:237 */
}
void ref_fsk_mod_c(void *h, float *out, const uint8_t *tx_bits)
{
    fsk_mod_c((struct FSK *)h, (COMP *)out, (uint8_t *)tx_bits);
}

