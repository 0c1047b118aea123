/*
 * wenet_oracle.c -- TEST INFRASTRUCTURE ONLY (see wenet_oracle.h).
 *
 * Scalar CPU restatement of the reference FSK-demod + deframe + LDPC path.
 * It is written flat (edge lists, one state struct, no per-codeword graph
 * allocation) but performs the same IEEE-754 operations in the same order as
 * the reference built with `gcc -O3` on x86-64 (no FMA contraction), so the
 * results are bit-identical; tests/test_oracle_vs_ref.py proves that against
 * oracle/_ref.  Build with -ffp-contract=off (oracle/Makefile does).
 */
#include "wenet_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "wb_tables.h"

/* ===================================================================== */
/* phi0 -- reference src/phi0.c:13-218.                                   */
/* The reference converts the argument to Q16 with a C cast               */
/* ((int32_t)(xf*65536), src/phi0.c:11) and walks a compare tree.  The     */
/* tree is a monotone step function of the Q16 integer, so it is fully     */
/* described by its breakpoints (wb_phi0_brk / wb_phi0_val, measured from  */
/* the compiled reference by tools/gen_tables.py).                         */
/* x86 cvttss2si returns INT32_MIN for NaN and out-of-range products, so   */
/* x >= 32768 lands in the "smaller than every threshold" leaf -> 10.0.    */
/* ===================================================================== */

static int32_t q16_like_x86(float xf)
{
    float p = xf * 65536.0f;
    if (!(p > -2147483904.0f && p < 2147483648.0f)) return INT32_MIN; /* NaN, inf, overflow */
    return (int32_t)p;
}

float wo_phi0(float xf)
{
    int32_t q = q16_like_x86(xf);
    int lo = 0, hi = WB_PHI0_NSTEPS - 1;
    if (q < 0) return 10.0f;
    /* last step whose breakpoint is <= q */
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (wb_phi0_brk[mid] <= q) lo = mid; else hi = mid - 1;
    }
    return wb_phi0_val[lo];
}

void wo_phi0_array(const float *x, float *y, long n)
{
    long i;
    for (i = 0; i < n; i++) y[i] = wo_phi0(x[i]);
}

/* ===================================================================== */
/* sd_to_llr -- reference src/mpdecode_core.c:569-595.                    */
/* All double, two sequential passes; the "L" constants make the EsN0      */
/* expression and the final product x87 long double (80-bit) on x86-64.    */
/* ===================================================================== */

void wo_sd_to_llr(float *llr, const double *sd, int n)
{
    double sum = 0.0, sumsq = 0.0, mean, estvar, estEsN0;
    int i;
    for (i = 0; i < n; i++) sum += fabs(sd[i]);            /* :577-579 */
    mean = sum / n;
    sum = 0.0;
    for (i = 0; i < n; i++) {                              /* :584-589 */
        double sign = (double)((sd[i] > 0.0) - (sd[i] < 0.0));
        double x = sd[i] / mean - sign;
        sum += x;
        sumsq += x * x;
    }
    estvar = (n * sumsq - sum * sum) / (n * (n - 1));      /* :590 */
    estEsN0 = (double)(1.0 / (2.0L * estvar + 1E-3));      /* :593, long double */
    for (i = 0; i < n; i++)
        llr[i] = (float)(4.0L * estEsN0 * sd[i]);          /* :594-595, long double */
}

/* ===================================================================== */
/* LDPC -- reference src/mpdecode_core.c:152-379 (graph), :385-489         */
/* (SumProduct), :494-566 (wrapper), with the parameters drs232_ldpc.c     */
/* passes (:128-138): dec_type 0, H1 = 1, shift = 0.                       */
/*                                                                         */
/* Flat edge list, check-major: check j owns edges coff[j] .. coff[j+1]-1  */
/* in the reference's c_nodes[j].subs[] order: the 12 H1 columns, then     */
/* parity column j-1 (j > 0), then parity column j (:230-246).             */
/* Variable i owns slots voff[i] .. voff[i+1]-1 in the reference's          */
/* v_nodes[i].subs[] order: H_cols order for data columns, checks q, q+1   */
/* for parity column q (:316-324).                                         */
/* ===================================================================== */

#define NEDGE (WB_NPAR * WB_ROWW + 2 * WB_NPAR - 1) /* 7223 */

static int g_init;
static int g_coff[WB_NPAR + 1];
static uint16_t g_evar[NEDGE];      /* variable of edge e */
static int g_voff[WB_NCODE + 1];
static uint16_t g_vedge[NEDGE];     /* edge id of (variable, slot) */

static void ldpc_graph_init(void)
{
    int j, k, i, e = 0, s = 0;
    if (g_init) return;
    for (j = 0; j < WB_NPAR; j++) {
        g_coff[j] = e;
        for (k = 0; k < WB_ROWW; k++) g_evar[e++] = wb_hrows[j * WB_ROWW + k];
        if (j > 0) g_evar[e++] = (uint16_t)(WB_NDATA + j - 1);
        g_evar[e++] = (uint16_t)(WB_NDATA + j);
    }
    g_coff[WB_NPAR] = e; /* == NEDGE */
    for (i = 0; i < WB_NCODE; i++) {
        int deg, c[3];
        g_voff[i] = s;
        if (i < WB_NDATA) {
            deg = WB_COLW;
            for (k = 0; k < deg; k++) c[k] = wb_hcols[i * WB_COLW + k];
        } else {
            int q = i - WB_NDATA;
            deg = (i == WB_NCODE - 1) ? 1 : 2;
            c[0] = q; c[1] = q + 1;
        }
        for (k = 0; k < deg; k++) {
            /* "search the connected c-node for the proper message value" :337-341 */
            int ee, found = -1;
            for (ee = g_coff[c[k]]; ee < g_coff[c[k] + 1]; ee++)
                if (g_evar[ee] == i) { found = ee; break; }
            g_vedge[s++] = (uint16_t)found;
        }
    }
    g_voff[WB_NCODE] = s;
    g_init = 1;
}

int wo_ldpc_decode(const float *llr, uint8_t *bits, int max_iter, int *pcc)
{
    static float q_mag[NEDGE];      /* v->c message, phi domain (v_sub_node.message) */
    static uint8_t q_sgn[NEDGE];    /* v->c sign (v_sub_node.sign) */
    static float r_msg[NEDGE];      /* c->v message (c_sub_node.message) */
    int iter, result, j, i, e, s;

    ldpc_graph_init();
    /* init :343-350 */
    for (i = 0; i < WB_NCODE; i++) {
        float m = wo_phi0((float)fabs(llr[i]));
        for (s = g_voff[i]; s < g_voff[i + 1]; s++) {
            q_mag[g_vedge[s]] = m;
            q_sgn[g_vedge[s]] = (llr[i] < 0);
        }
    }
    for (e = 0; e < NEDGE; e++) r_msg[e] = 0.0f;
    memset(bits, 0, WB_NCODE);

    result = max_iter;
    for (iter = 0; iter < max_iter; iter++) {
        int ssum = 0, nonzero_data = 0;
        memset(bits, 0, WB_NCODE);
        /* check-node pass :412-436 */
        for (j = 0; j < WB_NPAR; j++) {
            int e0 = g_coff[j], e1 = g_coff[j + 1];
            int sign = q_sgn[e0];
            float phi_sum = q_mag[e0];
            for (e = e0 + 1; e < e1; e++) { phi_sum += q_mag[e]; sign ^= q_sgn[e]; }
            if (sign == 0) ssum++;
            for (e = e0; e < e1; e++) {
                float v = wo_phi0(phi_sum - q_mag[e]);
                r_msg[e] = (sign ^ q_sgn[e]) ? -v : v;
            }
        }
        /* variable-node pass :439-464 */
        for (i = 0; i < WB_NCODE; i++) {
            float Qi = llr[i];
            for (s = g_voff[i]; s < g_voff[i + 1]; s++) Qi += r_msg[g_vedge[s]];
            if (Qi < 0) bits[i] = 1;
            for (s = g_voff[i]; s < g_voff[i + 1]; s++) {
                float t = Qi - r_msg[g_vedge[s]];
                q_mag[g_vedge[s]] = wo_phi0((float)fabs(t));
                q_sgn[g_vedge[s]] = (t > 0) ? 0 : 1;
            }
        }
        /* exits :467-483 -- data[] is all zero in run_ldpc_decoder (:541) */
        for (i = 0; i < WB_NDATA; i++) nonzero_data += bits[i];
        if (nonzero_data == 0) { result = iter + 1; break; }
        *pcc = ssum;
        if (ssum == WB_NPAR) { result = iter + 1; break; }
    }
    return result;
}

/* reference src/mpdecode_core.c:72-91 */
void wo_ldpc_encode(const uint8_t *ibits, uint8_t *pbits)
{
    unsigned p, k, prev = 0;
    for (p = 0; p < WB_NPAR; p++) {
        unsigned par = prev;
        for (k = 0; k < WB_ROWW; k++) par += ibits[wb_hrows[p * WB_ROWW + k]];
        prev = par & 1u;
        pbits[p] = (uint8_t)prev;
    }
}

/* reference src/drs232_ldpc.c:91-102: CRC16-CCITT-FALSE, bitwise form */
uint16_t wo_crc16(const uint8_t *data, int n)
{
    uint16_t crc = 0xFFFF;
    int i, b;
    for (i = 0; i < n; i++) {
        crc ^= (uint16_t)(data[i] << 8);
        for (b = 0; b < 8; b++)
            crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}

/* ===================================================================== */
/* Deframer -- reference src/drs232_ldpc.c:176-275 / src/wenet_ldpc.c.     */
/* ===================================================================== */

#define PKT_BYTES 256
#define FRAME_BYTES (256 + 2 + 65)       /* payload + crc + parity = 323 */

struct wo_deframer {
    int mode;              /* 1 = RS232 v1, 2 = scrambled v2 */
    int max_iter;
    int uw_bits, uw_thresh, bits_per_byte, nsym;
    uint8_t uw[40];
    uint8_t window[40];    /* bit_buffer: NOT updated while collecting */
    int collecting, ind;
    double buf[FRAME_BYTES * 10];
    long fed;              /* symbols consumed so far */
    long first_pos;
    uint16_t packets, packet_errors;
};

wo_deframer *wo_deframer_create(int mode, int max_iter)
{
    static const uint8_t uwbytes[4] = {0xAB, 0xCD, 0xEF, 0x01}; /* tx/PacketTX.py:65 */
    wo_deframer *d = (wo_deframer *)calloc(1, sizeof(*d));
    int b, k, n = 0;
    if (!d) return NULL;
    d->mode = mode;
    d->max_iter = max_iter;
    if (mode == 1) {
        /* RS232: start 0, 8 data bits LSB first, stop 1 (drs232_ldpc.c:77-86) */
        for (b = 0; b < 4; b++) {
            d->uw[n++] = 0;
            for (k = 0; k < 8; k++) d->uw[n++] = (uwbytes[b] >> k) & 1;
            d->uw[n++] = 1;
        }
        d->uw_bits = 40; d->uw_thresh = 35; d->bits_per_byte = 10;
    } else {
        for (b = 0; b < 4; b++)
            for (k = 7; k >= 0; k--) d->uw[n++] = (uwbytes[b] >> k) & 1; /* wenet_ldpc.c:77-82 */
        d->uw_bits = 32; d->uw_thresh = 28; d->bits_per_byte = 8;
    }
    d->nsym = FRAME_BYTES * d->bits_per_byte;
    return d;
}

void wo_deframer_destroy(wo_deframer *d) { free(d); }

void wo_deframer_counts(wo_deframer *d, int *packets, int *packet_errors)
{
    *packets = d->packets;
    *packet_errors = d->packet_errors;
}

long wo_deframer_feed(wo_deframer *d, const float *sd, long n,
                      uint8_t *out, long out_cap,
                      float *tap_llr, int *tap_iters, int *tap_pcc,
                      uint8_t *tap_crc_ok, long *tap_pos, uint8_t *tap_bytes,
                      long tap_cap, long *n_cw)
{
    long t, nout = 0, ncw = 0;
    static double cw[FRAME_BYTES * 8];
    static float llr[WB_NCODE + 8];
    static uint8_t bits[WB_NCODE];
    static int pcc_keep; /* the reference's parityCheckCount is an uninitialised local that
                            persists across packets; only reported through the tap */
    for (t = 0; t < n; t++) {
        float symbol = sd[t];
        int was_collecting = d->collecting;
        if (!was_collecting) {
            int i, score = 0;
            uint8_t bit = symbol < 0;
            memmove(d->window, d->window + 1, (size_t)(d->uw_bits - 1));
            d->window[d->uw_bits - 1] = bit;
            for (i = 0; i < d->uw_bits; i++) score += (d->window[i] == d->uw[i]);
            if (score >= d->uw_thresh) { d->ind = 0; d->collecting = 1; d->first_pos = d->fed + t + 1; }
        } else {
            if (d->mode == 2)
                d->buf[d->ind] = symbol * (wb_scramble_neg[d->ind % WB_SCRAMBLE_LEN] ? -1.0 : 1.0);
            else
                d->buf[d->ind] = symbol;
            d->ind++;
            if (d->ind == d->nsym) {
                int i, j, iters;
                uint8_t pkt[PKT_BYTES + 2];
                uint16_t rx, tx;
                if (d->mode == 1) {
                    /* strip start/stop, LSB-first -> MSB-first (drs232_ldpc.c:220-225) */
                    int k = 0;
                    for (i = 0; i < d->nsym; i += 10, k += 8)
                        for (j = 0; j < 8; j++) cw[k + j] = d->buf[i + 8 - j];
                } else {
                    memcpy(cw, d->buf, sizeof(double) * (size_t)d->nsym);
                }
                wo_sd_to_llr(llr, cw, WB_NCODE);
                iters = wo_ldpc_decode(llr, bits, d->max_iter, &pcc_keep);
                for (i = 0; i < PKT_BYTES + 2; i++) {
                    uint8_t a = 0;
                    for (j = 0; j < 8; j++) a |= (uint8_t)(bits[8 * i + j] << (7 - j));
                    pkt[i] = a;
                }
                rx = wo_crc16(pkt, PKT_BYTES);
                tx = (uint16_t)(pkt[PKT_BYTES] + (pkt[PKT_BYTES + 1] << 8));
                d->packets++;
                if (rx == tx) {
                    if (nout + PKT_BYTES <= out_cap) memcpy(out + nout, pkt, PKT_BYTES);
                    nout += PKT_BYTES;
                } else {
                    d->packet_errors++;
                }
                if (ncw < tap_cap) {
                    if (tap_llr) memcpy(tap_llr + ncw * WB_NCODE, llr, sizeof(float) * WB_NCODE);
                    if (tap_iters) tap_iters[ncw] = iters;
                    if (tap_pcc) tap_pcc[ncw] = pcc_keep;
                    if (tap_crc_ok) tap_crc_ok[ncw] = (rx == tx);
                    if (tap_pos) tap_pos[ncw] = d->first_pos;
                    if (tap_bytes) memcpy(tap_bytes + ncw * (PKT_BYTES + 2), pkt, PKT_BYTES + 2);
                }
                ncw++;
                d->collecting = 0;
            }
        }
    }
    d->fed += n;
    if (n_cw) *n_cw = ncw;
    return nout;
}

/* ===================================================================== */
/* FSK demodulator -- reference src/fsk.c.                                 */
/* ===================================================================== */

typedef struct { float r, i; } cpx;

static cpx cmul(cpx a, cpx b)       /* comp_prim.h cmult */
{
    cpx c;
    c.r = a.r * b.r - a.i * b.i;
    c.i = a.r * b.i + a.i * b.r;
    return c;
}
static cpx cconjg(cpx a) { a.i = -a.i; return a; }
static cpx cexpj(float phi) { cpx c; c.r = cosf(phi); c.i = sinf(phi); return c; } /* comp_exp_j */

#define MAXM 4
#define EYE_TR 8
#define EYE_IND 160

struct wo_fsk {
    int Fs, Rs, Ts, P, M, N, Nsym, Nmem, Nbits, Ndft, nstash, nin;
    int est_min, est_max, est_space;
    float *hann;         /* Ndft */
    cpx *tw;             /* Ndft twiddles */
    int fac[64];         /* kiss_fft factor list: p1,m1,p2,m2,... */
    cpx phi_c[MAXM];
    float f_est[MAXM];
    float norm_rx_timing, ppm, EbNodB;
    cpx *samp_old;       /* nstash */
    float *fft_est;      /* Ndft/2 */
    /* stats (struct MODEM_STATS fields the demod writes) */
    float snr_est, rx_timing, foff, clock_offset;
    int neyetr, neyesamp;
    float rx_eye[EYE_TR][EYE_IND];
    /* scratch */
    cpx *fin, *fout, *f_int[MAXM];
    /* modulator, reference struct FSK f1_tx / fs_tx / tx_phase_c */
    int f1_tx, fs_tx;
    cpx tx_phase_c;
};

/* kiss_fft factorisation, reference src/kiss_fft.c:309-330 */
static void fft_factor(int n, int *fac)
{
    int p = 4;
    double fs = floor(sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2; else if (p == 2) p = 3; else p += 2;
            if (p > fs) p = n;
        }
        n /= p;
        *fac++ = p;
        *fac++ = n;
    } while (n > 1);
}

/* radix-2 / radix-4 combine steps, reference src/kiss_fft.c:22-90 (forward) */
static void bfly2(cpx *F, int fstride, const cpx *tw, int m)
{
    int k;
    for (k = 0; k < m; k++) {
        cpx t = cmul(F[m + k], tw[k * fstride]);
        F[m + k].r = F[k].r - t.r; F[m + k].i = F[k].i - t.i;
        F[k].r += t.r; F[k].i += t.i;
    }
}

static void bfly4(cpx *F, int fstride, const cpx *tw, int m)
{
    int k;
    for (k = 0; k < m; k++) {
        cpx s0 = cmul(F[k + m], tw[k * fstride]);
        cpx s1 = cmul(F[k + 2 * m], tw[2 * k * fstride]);
        cpx s2 = cmul(F[k + 3 * m], tw[3 * k * fstride]);
        cpx s5, s3, s4, a = F[k];
        s5.r = a.r - s1.r; s5.i = a.i - s1.i;
        a.r += s1.r; a.i += s1.i;
        s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
        s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
        F[k + 2 * m].r = a.r - s3.r; F[k + 2 * m].i = a.i - s3.i;
        a.r += s3.r; a.i += s3.i;
        F[k] = a;
        F[k + m].r = s5.r + s4.i; F[k + m].i = s5.i - s4.r;
        F[k + 3 * m].r = s5.r - s4.i; F[k + 3 * m].i = s5.i + s4.r;
    }
}

/* decimation-in-time recursion, reference src/kiss_fft.c:238-306 */
static void fft_work(cpx *F, const cpx *f, int fstride, const int *fac, const cpx *tw)
{
    int p = fac[0], m = fac[1], k;
    if (m == 1) {
        for (k = 0; k < p; k++) F[k] = f[k * fstride];
    } else {
        for (k = 0; k < p; k++) fft_work(F + k * m, f + k * fstride, fstride * p, fac + 2, tw);
    }
    if (p == 4) bfly4(F, fstride, tw, m);
    else bfly2(F, fstride, tw, m); /* Ndft is a power of two: only radix 4 and 2 occur */
}

static void eye_geometry(const wo_fsk *f, int *dec, int *nsamp)
{
    /* reference src/fsk.c:1039-1040 / :415-416 */
    *dec = (int)ceil(((float)f->P * 2) / EYE_IND);
    *nsamp = (f->P * 2) / *dec;
}

wo_fsk *wo_fsk_create(int Fs, int Rs, int P, int M)
{
    wo_fsk *f;
    int i, m, Ndft = 0, dec, ns;
    if (Fs <= 0 || Rs <= 0 || P <= 0 || (Fs % Rs) || ((Fs / Rs) % P) || !(M == 2 || M == 4)) return NULL;
    f = (wo_fsk *)calloc(1, sizeof(*f));
    if (!f) return NULL;
    f->Fs = Fs; f->Rs = Rs; f->Ts = Fs / Rs; f->P = P; f->M = M;
    f->Nsym = 48;                               /* fsk.c:134 */
    f->N = f->Ts * f->Nsym;
    f->Nmem = f->N + 2 * f->Ts;
    f->nin = f->N;
    f->Nbits = (M == 2) ? f->Nsym : 2 * f->Nsym;
    for (i = 1; i > 0 && i <= f->N; i <<= 1) if (f->N & i) Ndft = i;   /* highest set bit, fsk.c:169-173 */
    f->Ndft = Ndft;
    f->est_min = Rs / 4;
    f->est_max = Fs / 2 - Rs / 4;
    f->est_space = Rs - Rs / 5;
    f->nstash = 4 * f->Ts;
    f->samp_old = (cpx *)calloc((size_t)f->nstash, sizeof(cpx));
    f->fft_est = (float *)calloc((size_t)Ndft / 2, sizeof(float));
    f->hann = (float *)calloc((size_t)Ndft, sizeof(float));
    f->tw = (cpx *)calloc((size_t)Ndft, sizeof(cpx));
    f->fin = (cpx *)calloc((size_t)Ndft, sizeof(cpx));
    f->fout = (cpx *)calloc((size_t)Ndft, sizeof(cpx));
    for (m = 0; m < M; m++) {
        f->f_int[m] = (cpx *)calloc((size_t)(f->Nsym + 1) * P, sizeof(cpx));
        f->phi_c[m].r = 1.0f; f->phi_c[m].i = 0.0f;   /* comp_exp_j(0) */
    }
    /* Hann table by oscillator recurrence, fsk.c:94-111 */
    {
        cpx dphi = cexpj((float)((2 * M_PI) / ((float)Ndft - 1)));
        cpx rphi; rphi.r = .5f; rphi.i = 0;
        rphi = cmul(cconjg(dphi), rphi);
        for (i = 0; i < Ndft; i++) {
            rphi = cmul(dphi, rphi);
            f->hann[i] = (float)(.5 - rphi.r);
        }
    }
    /* kiss_fft twiddles, kiss_fft.c:357-363 */
    for (i = 0; i < Ndft; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / Ndft;
        f->tw[i].r = cosf((float)phase);
        f->tw[i].i = sinf((float)phase);
    }
    fft_factor(Ndft, f->fac);
    eye_geometry(f, &dec, &ns);
    f->neyesamp = ns;
    f->neyetr = M * (EYE_TR / M);
    return f;
}

void wo_fsk_destroy(wo_fsk *f)
{
    int m;
    if (!f) return;
    for (m = 0; m < f->M; m++) free(f->f_int[m]);
    free(f->samp_old); free(f->fft_est); free(f->hann); free(f->tw); free(f->fin); free(f->fout);
    free(f);
}

void wo_fsk_set_est_limits(wo_fsk *f, int lo, int hi)
{
    f->est_min = lo < 0 ? 0 : lo;
    f->est_max = hi;
}

int wo_fsk_nin(wo_fsk *f) { return f->nin; }
int wo_fsk_nbits(wo_fsk *f) { return f->Nbits; }

/* reference src/fsk.c:540-677 */
static void freq_est(wo_fsk *f, const cpx *in, float *freqs)
{
    int Ndft = f->Ndft, Fs = f->Fs, nin = f->nin, M = f->M;
    int f_min = (f->est_min * Ndft) / Fs;
    int f_max = (f->est_max * Ndft) / Fs;
    int f_zero = (f->est_space * Ndft) / Fs;
    float tc = (float)(0.95 * Ndft / Fs);
    int loops = nin / Ndft, j, i, m;
    int freqi[MAXM];
    cpx *fin = f->fin, *fout = f->fout;

    for (j = 0; j < loops; j++) {
        int samps = nin - (j + 1) * Ndft;
        int nwin = samps >= Ndft ? Ndft : samps;
        for (i = 0; i < nwin; i++) {
            fin[i].r = f->hann[i] * in[i + Ndft * j].r;
            fin[i].i = f->hann[i] * in[i + Ndft * j].i;
        }
        for (; i < Ndft; i++) { fin[i].r = 0; fin[i].i = 0; }
        fft_work(fout, fin, 1, f->fac, f->tw);
        for (i = 0; i < Ndft / 2; i++) fout[i].r = fout[i].r * fout[i].r + fout[i].i * fout[i].i;
        for (i = 0; i < f_min; i++) fout[i].r = 0;
        for (i = f_max - 1; i < Ndft / 2; i++) fout[i].r = 0;
        for (i = 0; i < Ndft / 2; i++) {
            f->fft_est[i] = (f->fft_est[i] * (1 - tc)) + (sqrtf(fout[i].r) * tc);
            fout[i].i = f->fft_est[i];
        }
    }
    for (m = 0; m < M; m++) {
        int imax = 0, lo, hi;
        float mx = 0;
        for (j = 0; j < Ndft / 2; j++)
            if (fout[j].i > mx) { mx = fout[j].i; imax = j; }
        lo = imax - f_zero; if (lo < 0) lo = 0;
        hi = imax + f_zero; if (hi > Ndft) hi = Ndft;
        for (j = lo; j < hi; j++) fout[j].i = 0;
        freqi[m] = imax;
    }
    /* ascending order (the reference's gnome sort, :658-667) */
    for (i = 1; i < M; i++) {
        int v = freqi[i];
        for (j = i; j > 0 && freqi[j - 1] > v; j--) freqi[j] = freqi[j - 1];
        freqi[j] = v;
    }
    for (m = 0; m < M; m++) freqs[m] = (float)freqi[m] * ((float)Fs / (float)Ndft);
}

/* reference src/fsk.c:679-1111 */
void wo_fsk_demod(wo_fsk *f, float *sd, uint8_t *bits, const float *in_f)
{
    const cpx *in = (const cpx *)in_f;
    int Ts = f->Ts, P = f->P, M = f->M, N = f->N, nsym = f->Nsym, nin = f->nin;
    int Nmem = f->Nmem, Fs = f->Fs, Rs = f->Rs, nstash = f->nstash;
    int nold = Nmem - nin, step = Ts / P, nint = (nsym + 1) * P;
    int m, i, j;
    float f_est[MAXM];
    cpx phi_c[MAXM], dphi[MAXM];
    cpx ring[256];

    for (m = 0; m < M; m++) phi_c[m] = f->phi_c[m];
    freq_est(f, in, f_est);
    if (f->f_est[0] < 1)                                   /* :729-732 first run */
        for (m = 0; m < M; m++) f->f_est[m] = f_est[m];

    for (m = 0; m < M; m++) {
        /* back the phase off over the re-integrated old samples, :756-759 */
        cpx back = cexpj((float)(-2 * (Nmem - nin - step) * M_PI * ((f->f_est[m]) / (float)Fs)));
        phi_c[m] = cmul(back, phi_c[m]);
        dphi[m] = cexpj((float)(2 * M_PI * ((f->f_est[m]) / (float)Fs)));
    }

    for (m = 0; m < M; m++) {
        const cpx *src = &f->samp_old[nstash - nold];
        cpx d = dphi[m], ph = phi_c[m];
        int dc = 0, old = 1, cb;
        cpx *out = f->f_int[m];
        /* one sample step of the mixer, with the old->new switch (:775-799, :805-826) */
#define MIX_STEP(slot)                                                              \
        do {                                                                        \
            if (dc >= nold && old) {                                                \
                float av;                                                           \
                src = in; dc = 0; old = 0;                                          \
                av = sqrtf(ph.r * ph.r + ph.i * ph.i);    /* comp_normalize */      \
                ph.r = ph.r / av; ph.i = ph.i / av;                                 \
                d = cexpj((float)(2 * M_PI * ((f_est[m]) / (float)Fs)));            \
            }                                                                       \
            ring[slot] = cmul(src[dc], cconjg(ph));                                 \
            ph = cmul(ph, d);                                                       \
        } while (0)
        for (dc = 0; dc < Ts - step; dc++) MIX_STEP(dc);
        cb = dc;
        for (i = 0; i < nint; i++) {
            float sr = 0, si = 0;
            for (j = 0; j < step; j++, dc++) MIX_STEP(cb + j);
            cb += step;
            if (cb >= Ts) cb = 0;
            for (j = 0; j < Ts; j++) { sr += ring[j].r; si += ring[j].i; }
            out[i].r = sr; out[i].i = si;
        }
#undef MIX_STEP
        phi_c[m] = ph;
    }

    for (m = 0; m < M; m++) { f->phi_c[m] = phi_c[m]; f->f_est[m] = f_est[m]; }
    memcpy(f->samp_old, &in[nin - nstash], sizeof(cpx) * (size_t)nstash);

    /* fine timing :853-884 */
    {
        cpx dft = cexpj((float)(2 * M_PI * ((float)Rs / (float)(P * Rs))));
        cpx pft, tc;
        float norm, rx_timing, old_norm, dn, fract;
        int low, high;
        float meanebno = 0, stdebno = 0;
        pft.r = 1; pft.i = 0; tc.r = 0; tc.i = 0;
        for (i = 0; i < nint; i++) {
            float e = 0;
            for (m = 0; m < M; m++)
                e += (f->f_int[m][i].r * f->f_int[m][i].r) + (f->f_int[m][i].i * f->f_int[m][i].i);
            tc.r = tc.r + e * pft.r;
            tc.i = tc.i + e * pft.i;
            pft = cmul(pft, dft);
        }
        if (isnan(tc.r) || isnan(tc.i)) return;            /* :878-880 */
        norm = (float)(atan2f(tc.i, tc.r) / (2 * M_PI));
        rx_timing = norm * (float)P;
        old_norm = f->norm_rx_timing;
        f->norm_rx_timing = norm;
        dn = norm - old_norm;
        if (fabsf(dn) < .2) {
            float appm = (float)(1e6 * dn / (float)nsym);
            f->ppm = (float)(.9 * f->ppm + .1 * appm);
        }
        if (norm > 0.25) f->nin = N + Ts / 2;
        else if (norm < -0.25) f->nin = N - Ts / 2;
        else f->nin = N;

        low = (int)floorf(rx_timing);
        fract = rx_timing - (float)low;
        high = (int)ceilf(rx_timing);

        /* symbol decisions :927-993 */
        for (i = 0; i < nsym; i++) {
            int st = (i + 1) * P, sym = 0;
            float tm[MAXM] = {0, 0, 0, 0}, mx, mn;
            for (m = 0; m < M; m++) {
                cpx a = f->f_int[m][st + low], b = f->f_int[m][st + high], t;
                t.r = (1 - fract) * a.r; t.i = (1 - fract) * a.i;
                t.r = t.r + fract * b.r; t.i = t.i + fract * b.i;
                tm[m] = (t.r * t.r) + (t.i * t.i);
            }
            mx = tm[0]; mn = tm[0];
            for (m = 0; m < M; m++) {
                if (tm[m] > mx) { mx = tm[m]; sym = m; }
                if (tm[m] < mn) mn = tm[m];
            }
            if (bits) {
                if (M == 2) bits[i] = (sym == 1);
                else { bits[2 * i + 1] = sym & 1; bits[2 * i] = (sym & 2) >> 1; }
            }
            if (sd) {
                for (m = 0; m < M; m++) tm[m] = sqrtf(tm[m]);
                if (M == 2) {
                    sd[i] = tm[0] - tm[1];
                } else {
                    float lo_b = -tm[0], hi_b = -tm[0];    /* sd[2i+1], sd[2i] */
                    lo_b += tm[1]; hi_b += -tm[1];
                    lo_b += -tm[2]; hi_b += tm[2];
                    lo_b += tm[3]; hi_b += tm[3];
                    sd[2 * i + 1] = lo_b; sd[2 * i] = hi_b;
                }
            }
            stdebno += mx;
            meanebno += sqrtf(mx);
        }
        /* Eb/N0 estimate :997-1010 */
        meanebno = meanebno / (float)nsym;
        stdebno = (stdebno / (float)nsym) - (meanebno * meanebno);
        if (stdebno > 0.0) stdebno = (float)sqrt(stdebno); else stdebno = 0.0;
        f->EbNodB = -6 + (20 * log10f((float)((1e-6 + meanebno) / (1e-6 + stdebno))));

        /* stats :1018-1080 */
        f->clock_offset = f->ppm;
        f->snr_est = (float)(.5 * f->snr_est + .5 * f->EbNodB);
        f->rx_timing = rx_timing;
        f->foff = (float)((1200 + 1200 + 400) / 2) - (f_est[0] + f_est[1]) / 2;
        {
            int dec, ns, tr, ntr = EYE_TR / M, off = high + 1;
            float emax = 0;
            eye_geometry(f, &dec, &ns);
            f->neyesamp = ns;
            f->neyetr = M * ntr;
            for (tr = 0; tr < ntr; tr++)
                for (m = 0; m < M; m++)
                    for (j = 0; j < ns; j++) {
                        cpx v = f->f_int[m][2 * P * tr + off + j * dec];
                        f->rx_eye[tr * M + m][j] = sqrtf(v.r * v.r + v.i * v.i);
                    }
            for (i = 0; i < M * ntr; i++)
                for (j = 0; j < ns; j++)
                    if (fabsf(f->rx_eye[i][j]) > emax) emax = fabsf(f->rx_eye[i][j]);
            for (i = 0; i < M * ntr; i++)
                for (j = 0; j < ns; j++) f->rx_eye[i][j] = f->rx_eye[i][j] / emax;
        }
    }
}

void wo_fsk_state(wo_fsk *f, float *out)
{
    int i;
    for (i = 0; i < 4; i++) { out[2 * i] = f->phi_c[i].r; out[2 * i + 1] = f->phi_c[i].i; }
    for (i = 0; i < 4; i++) out[8 + i] = f->f_est[i];
    out[12] = f->norm_rx_timing; out[13] = f->ppm; out[14] = f->EbNodB;
    out[15] = f->snr_est; out[16] = f->rx_timing; out[17] = f->foff; out[18] = f->clock_offset;
}

void wo_fsk_fft_est(wo_fsk *f, float *out) { memcpy(out, f->fft_est, sizeof(float) * (size_t)(f->Ndft / 2)); }

void wo_fsk_eye(wo_fsk *f, int *neyetr, int *neyesamp, float *out)
{
    int i, j;
    *neyetr = f->neyetr; *neyesamp = f->neyesamp;
    for (i = 0; i < f->neyetr; i++)
        for (j = 0; j < f->neyesamp; j++) out[i * f->neyesamp + j] = f->rx_eye[i][j];
}

/* sample conversion of one frame, reference src/fsk_demod.c:273-296 */
static void load_frame(float *mod, int fmt, const void *raw, long pos, int nin)
{
    int i;
    if (fmt == 0) {
        memcpy(mod, (const float *)raw + 2 * pos, sizeof(float) * 2 * (size_t)nin);
    } else if (fmt == 1) {
        const uint8_t *p = (const uint8_t *)raw + 2 * pos;
        for (i = 0; i < 2 * nin; i++) mod[i] = (float)(((float)p[i] - 127.0) / 128.0);
    } else if (fmt == 2) {
        const int16_t *p = (const int16_t *)raw + 2 * pos;
        for (i = 0; i < 2 * nin; i++) mod[i] = ((float)p[i]) / 1000;
    } else {
        const int16_t *p = (const int16_t *)raw + pos;
        for (i = 0; i < nin; i++) { mod[2 * i] = ((float)p[i]) / 1000; mod[2 * i + 1] = 0.0f; }
    }
}

/* frame loop + input conversion, reference src/fsk_demod.c:270-299, :403-412 */
long wo_fsk_run(wo_fsk *f, int fmt, const void *raw, long nsamp,
                float *sd_out, long sd_cap, long *n_sd,
                float *frame_log, long log_cap, long *consumed)
{
    long pos = 0, frames = 0, nsd = 0;
    int nmax = f->N + 2 * f->Ts;
    float *mod = (float *)malloc(sizeof(float) * 2 * (size_t)nmax);
    float *sdbuf = (float *)calloc((size_t)f->Nbits, sizeof(float));
    while (pos + f->nin <= nsamp) {
        int nin = f->nin;
        load_frame(mod, fmt, raw, pos, nin);
        wo_fsk_demod(f, sdbuf, NULL, mod);
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
    free(mod); free(sdbuf);
    *n_sd = nsd; *consumed = pos;
    return frames;
}

/* the same loop with the hard bits of fsk_demod() (fsk_demod without -s: src/fsk_demod.c:301, :405), one byte each */
long wo_fsk_run_bits(wo_fsk *f, int fmt, const void *raw, long nsamp, uint8_t *bits_out, long cap, long *n_bits)
{
    long pos = 0, frames = 0, nb = 0;
    float *mod = (float *)malloc(sizeof(float) * 2 * (size_t)(f->N + 2 * f->Ts));
    uint8_t *bitbuf = (uint8_t *)calloc((size_t)f->Nbits, 1);
    while (pos + f->nin <= nsamp) {
        int nin = f->nin;
        load_frame(mod, fmt, raw, pos, nin);
        wo_fsk_demod(f, NULL, bitbuf, mod);
        pos += nin;
        if (nb + f->Nbits <= cap) { memcpy(bits_out + nb, bitbuf, (size_t)f->Nbits); nb += f->Nbits; }
        frames++;
    }
    free(mod); free(bitbuf);
    *n_bits = nb;
    return frames;
}

/* ===================================================================== */
/* Transmit side: what the receive path above decodes.   */
/* ===================================================================== */

/* One on-air frame as 0/1 bytes in transmit order.
 *   reference tx/PacketTX.py:123-137 (frame_packet): payload padded with 0x55 to 256 bytes, CRC16 little endian,
 *     parity = ldpc_encode(payload + crc) packed MSB first into 65 bytes (tx/ldpc_encoder.py:42-52),
 *     preamble (16 x 0x55, tx/PacketTX.py:65) + unique word AB CD EF 01 (:66) + scramble(body);
 *   mode 1 (RS232): no scrambling (tx/radio_wrappers.py:502-503); the UART sends every byte as start 0, 8 data bits
 *     LSB first, stop 1 -- the framing src/drs232_ldpc.c:211-225 undoes;
 *   mode 2: body XORed with the 125-byte table of tx/radio_wrappers.py:385-404 (= the signs of
 *     src/wenet_scramble.h:22 packed MSB first), bits MSB first (tx/radio_wrappers.py:410-417).
 * bits must hold 3430 (mode 1) / 2744 (mode 2) bytes; returns the count. */
int wo_tx_frame_bits(const uint8_t *payload, int payload_len, int mode, uint8_t *bits)
{
    static const uint8_t head[20] = {0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55, 0x55,
                                     0xAB, 0xCD, 0xEF, 0x01};
    uint8_t raw[20 + FRAME_BYTES], ib[WB_NDATA], pb[WB_NPAR + 4];
    uint16_t crc;
    int i, k, n = 0;
    memcpy(raw, head, 20);
    for (i = 0; i < PKT_BYTES; i++) raw[20 + i] = (i < payload_len) ? payload[i] : 0x55;
    crc = wo_crc16(raw + 20, PKT_BYTES);
    raw[20 + PKT_BYTES] = (uint8_t)(crc & 0xFF);               /* struct.pack("<H", crc) */
    raw[20 + PKT_BYTES + 1] = (uint8_t)(crc >> 8);
    for (i = 0; i < WB_NDATA; i++) ib[i] = (raw[20 + i / 8] >> (7 - i % 8)) & 1;   /* np.unpackbits: MSB first */
    wo_ldpc_encode(ib, pb);
    for (i = 0; i < 4; i++) pb[WB_NPAR + i] = 0;                /* np.packbits pads with zeros */
    for (i = 0; i < 65; i++) {
        uint8_t b = 0;
        for (k = 0; k < 8; k++) b = (uint8_t)((b << 1) | pb[8 * i + k]);
        raw[20 + PKT_BYTES + 2 + i] = b;
    }
    if (mode == 2) {
        for (i = 0; i < FRAME_BYTES; i++) {
            uint8_t sc = 0;
            for (k = 0; k < 8; k++) sc = (uint8_t)((sc << 1) | wb_scramble_neg[(8 * (i % 125) + k) % WB_SCRAMBLE_LEN]);
            raw[20 + i] ^= sc;
        }
        for (i = 0; i < 20 + FRAME_BYTES; i++)
            for (k = 7; k >= 0; k--) bits[n++] = (raw[i] >> k) & 1;
    } else {
        for (i = 0; i < 20 + FRAME_BYTES; i++) {
            bits[n++] = 0;
            for (k = 0; k < 8; k++) bits[n++] = (raw[i] >> k) & 1;
            bits[n++] = 1;
        }
    }
    return n;
}

/* reference src/fsk.c:1162-1204 (fsk_mod_c): Nsym symbols -> Nsym*Ts complex samples of amplitude 2; the phase
 * carries over between calls (fsk->tx_phase_c, normalised at the end of every call).  tx_bits holds Nbits 0/1 bytes. */
void wo_fsk_mod_c(wo_fsk *f, float *out, const uint8_t *tx_bits)
{
    cpx ph = f->tx_phase_c, dosc[MAXM];
    int m, i, j, bit_i = 0;
    for (m = 0; m < f->M; m++)
        dosc[m] = cexpj((float)(2 * M_PI * ((float)(f->f1_tx + (f->fs_tx * m)) / (float)(f->Fs))));
    for (i = 0; i < f->Nsym; i++) {
        int sym = 0;
        for (m = f->M; m >>= 1;) { sym = (sym << 1) | (tx_bits[bit_i] == 1 ? 1 : 0); bit_i++; }
        for (j = 0; j < f->Ts; j++) {
            ph = cmul(ph, dosc[sym]);
            out[2 * (i * f->Ts + j)] = 2 * ph.r;                /* fcmult(2, tx_phase_c) */
            out[2 * (i * f->Ts + j) + 1] = 2 * ph.i;
        }
    }
    {   /* comp_normalize, src/comp_prim.h:133-139 */
        float av = sqrtf(ph.r * ph.r + ph.i * ph.i);
        ph.r = ph.r / av; ph.i = ph.i / av;
    }
    f->tx_phase_c = ph;
}

void wo_fsk_set_tx(wo_fsk *f, int f1_tx, int fs_tx)
{
    f->f1_tx = f1_tx; f->fs_tx = fs_tx;
    f->tx_phase_c.r = 1.0f; f->tx_phase_c.i = 0.0f;             /* comp_exp_j(0), src/fsk.c:237 */
}

