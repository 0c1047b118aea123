/*
 * wb_engine.cu -- host side of libwenet_b200.so: the C ABI of include/wenet_b200.h.
 *
 * Owns the HBM layout described in wb_internal.h, builds the constant tables the kernels use
 * (with the host libm, i.e. the cosf/sinf the reference itself would call), and sequences the
 * kernels on one CUDA stream:
 *
 *   wb_process:  K1 wb_fsk_kernel -> K2 wb_deframe_kernel -> K3 wb_llr_stats_kernel
 *                -> K4 wb_ldpc_kernel -> wb_carry_kernel
 *
 * There is NO CPU fallback: without a CUDA device wb_create fails with WB_ENODEV.
 */
#include <cuda_runtime.h>

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "wb_internal.h"
#include "wb_math.h"
#include "wb_phi0.h"
#include "wb_fsk_kernel.cuh"
#include "wb_deframe_kernel.cuh"
#include "wb_ldpc_kernel.cuh"
#include "wb_tx_kernel.cuh"
#include "wb_vperm.h"

#define WB_HEADROOM 512u    /* samples in front of every input row for the parked remainder (>= WB_MAX_NIN) */

static thread_local char g_err[512] = "";

static int wb_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return wb_fail(WB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

struct wb_engine {
    wb_config cfg;
    wb_fsk_params fp;
    wb_deframe_params dp;
    int max_iter;
    cudaStream_t stream;
    cudaEvent_t ev[8];
    cudaEvent_t tev0, tev1;
    /* device buffers */
    unsigned char *d_in;
    unsigned long long in_stride, in_cap;          /* bytes per row; samples per row (headroom + chunk) */
    wb_stream_state *d_state;
    wb_cursor *d_cursor;
    float *d_sd;
    unsigned char *d_hard;           /* WB_FLAG_HARD_BITS */
    bool hard_valid;                 /* the last process call ran the demodulator (wb_process, not wb_process_soft) */
    unsigned long long sd_stride;
    unsigned sd_cap;
    unsigned *d_jobs;
    double *d_c4;
    wb_codeword *d_cw, *d_cw_packed;
    float *d_llr, *d_llr_packed;
    unsigned *d_gather;
    unsigned long long *d_fill;
    int job_cap;
    float *d_frame_log;
    int log_cap;
    /* tables */
    void *d_tables;
    ushort4 *d_vedge;
    uint16_t *d_crc_tab;
    unsigned crc0;
    uint8_t *d_scramble;
    wb_phi0_flag *d_lut;
    /* resident LDPC benchmark */
    float *d_bench_llr;
    uint8_t *d_bench_bits;
    int *d_bench_iters, *d_bench_pcc;
    size_t bench_n;
    /* host mirrors */
    std::vector<wb_cursor> cursor;                 /* as of the last wb_sync */
    std::vector<unsigned long long> fill;          /* samples resident per stream (host's view) */
    std::vector<int> nin;
    /* payload bytes and, packet for packet, the codeword sequence number (wb_codeword.seq) they came from */
    struct pktq { std::vector<uint8_t> buf; std::vector<uint32_t> seqs; size_t rd = 0; size_t size() const { return buf.size() - rd; }
                  void clear() { buf.clear(); seqs.clear(); rd = 0; } };
    std::vector<pktq> packets;                     /* CRC-valid payloads not yet drained, per stream */
    std::vector<wb_codeword> last_cw;              /* codewords of the last wb_process, (stream, seq) order */
    std::vector<float> last_llr;
    bool uniform_fill;                             /* every stream has the same fill (strided feed possible) */
    bool pending;                                  /* a wb_process has not been collected yet */
    bool resident_mode;                            /* wb_dev_set_fill: do not compact */
    uint8_t *d_tx_bits;                            /* wb_tx_synthesize: on-air bits, one byte each */
    unsigned long long tx_bits_stride, tx_nbits;
    uint16_t *d_hrows;
    uint64_t launches;
    uint64_t last_codewords, last_samples;
    float kernel_ms[4];
    size_t fsk_smem;
    int spb;
};

/* ---- table construction ------------------------------------------------- */

typedef struct { float r, i; } hcpx;
static hcpx hcmul(hcpx a, hcpx b)
{
    hcpx c;
    c.r = a.r * b.r - a.i * b.i;
    c.i = a.r * b.i + a.i * b.r;
    return c;
}
static hcpx hcexpj(float phi) { hcpx c; c.r = cosf(phi); c.i = sinf(phi); return c; }

/* kiss_fft factorisation, reference src/kiss_fft.c:309-330 */
static int fft_factor(int n, int *fac)
{
    int p = 4, cnt = 0;
    double fs = floor(sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2; else if (p == 2) p = 3; else p += 2;
            if (p > fs) p = n;
        }
        n /= p;
        fac[2 * cnt] = p; fac[2 * cnt + 1] = n; cnt++;
    } while (n > 1);
    return cnt;
}

static void fft_perm(uint16_t *perm, int out_base, int in_base, int fstride, const int *fac)
{
    int p = fac[0], m = fac[1], k;
    if (m == 1) {
        for (k = 0; k < p; k++) perm[out_base + k] = (uint16_t)(in_base + k * fstride);
    } else {
        for (k = 0; k < p; k++) fft_perm(perm, out_base + k * m, in_base + k * fstride, fstride * p, fac + 2);
    }
}

static uint16_t crc16_ccitt(const uint8_t *d, int n, uint16_t crc)
{
    for (int i = 0; i < n; i++) {
        crc ^= (uint16_t)(d[i] << 8);
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}

static int build_fsk_params(const wb_config *cfg, wb_fsk_params *fp)
{
    int Fs = cfg->Fs, Rs = cfg->Rs, M = cfg->M, P = cfg->P;
    memset(fp, 0, sizeof(*fp));
    if (Fs <= 0 || Rs <= 0 || Rs > Fs) return wb_fail(WB_EINVAL, "invalid Fs/Rs");
    if (!(M == 2 || M == 4)) return wb_fail(WB_EINVAL, "M must be 2 or 4");
    if (Fs % Rs) return wb_fail(WB_EINVAL, "Fs must be an integer multiple of Rs (reference src/fsk.c:137)");
    if (P <= 0) P = Fs / Rs;                                   /* reference src/fsk_demod.c:186-188 */
    int Ts = Fs / Rs;
    if (Ts % P) return wb_fail(WB_EINVAL, "Fs/Rs must be a multiple of P (reference src/fsk.c:139)");
    if (Ts > WB_MAX_TS || Ts < 4) return wb_fail(WB_EINVAL, "Fs/Rs = %d unsupported (4..%d samples per symbol)", Ts, WB_MAX_TS);
    fp->Fs = Fs; fp->Rs = Rs; fp->Ts = Ts; fp->P = P; fp->M = M;
    fp->Nsym = WB_FRAME_SYMS;
    fp->N = Ts * fp->Nsym;
    fp->Nmem = fp->N + 2 * Ts;
    fp->Nbits = (M == 2) ? fp->Nsym : 2 * fp->Nsym;
    int Ndft = 0;
    for (int i = 1; i > 0 && i <= fp->N; i <<= 1) if (fp->N & i) Ndft = i;   /* reference src/fsk.c:169-173 */
    fp->Ndft = Ndft;
    fp->nstash = 4 * Ts;
    fp->nst = 2 * Ts + Ts / 2;
    fp->step = Ts / P;
    fp->nint = (fp->Nsym + 1) * P;
    fp->nsteps = fp->Nmem - fp->step;
    fp->nmax = fp->N + Ts / 2;
    if (Ndft > WB_MAX_NDFT || Ndft < 128) return wb_fail(WB_EINVAL, "Ndft = %d unsupported", Ndft);
    if (fp->nint > WB_MAX_NINT) return wb_fail(WB_EINVAL, "too many integrator outputs per frame");
    if (fp->N - Ts / 2 < Ndft || fp->nmax >= 2 * Ndft || fp->nmax > WB_MAX_NIN)
        return wb_fail(WB_EINVAL, "frame length %d vs Ndft %d unsupported (needs exactly one estimator FFT per frame)", fp->N, Ndft);
    if (Fs / Ndft < 1) return wb_fail(WB_EINVAL, "Fs too low");
    int est_min = Rs / 4, est_max = Fs / 2 - Rs / 4, est_space = Rs - Rs / 5;   /* reference src/fsk.c:175-180 */
    if (cfg->est_lo > 0 || cfg->est_hi > 0) {                  /* fsk_set_est_limits, reference src/fsk.c:522-535 */
        est_min = cfg->est_lo < 0 ? 0 : cfg->est_lo;
        est_max = cfg->est_hi;
    }
    fp->f_min = (int)(((long long)est_min * Ndft) / Fs);
    fp->f_max = (int)(((long long)est_max * Ndft) / Fs);
    fp->f_zero = (int)(((long long)est_space * Ndft) / Fs);
    fp->tc = (float)(0.95 * Ndft / Fs);
    fp->in_fmt = cfg->in_fmt;
    fp->stats = (cfg->flags & WB_FLAG_STATS) ? 1 : 0;
    switch (cfg->in_fmt) {
    case WB_FMT_CF32: fp->in_bps = 8; break;
    case WB_FMT_CS16: fp->in_bps = 4; break;
    case WB_FMT_CU8: case WB_FMT_S16: fp->in_bps = 2; break;
    default: return wb_fail(WB_EINVAL, "unknown in_fmt %d", cfg->in_fmt);
    }
    if (Ndft != WB_MAX_NDFT) return wb_fail(WB_EINVAL, "Ndft = %d: the kernels are built for %d", Ndft, WB_MAX_NDFT);
    /* shared-memory geometry: the formulas the kernel uses (wb_internal.h) */
    fp->xlen = wb_geom_xlen(Ts);
    fp->ylen = wb_geom_ylen(Ts, fp->step);
    fp->blen = wb_geom_blen(M, fp->ylen);
    fp->sreg = wb_geom_sreg(fp->xlen, fp->blen, wb_geom_efl(P));
    return WB_OK;
}

static int upload_tables(wb_engine *e)
{
    wb_fsk_params &fp = e->fp;
    const int Ndft = fp.Ndft, nh = Ndft / 2;
    /* one allocation, 256-byte aligned sections */
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_hann = take(sizeof(float) * Ndft), o_tw = take(sizeof(float2) * Ndft), o_perm = take(sizeof(uint16_t) * Ndft);
    size_t o_pft = take(sizeof(float2) * fp.nint), o_dphi = take(sizeof(float2) * nh), o_back = take(sizeof(float2) * 3 * nh);
    std::vector<unsigned char> h(off, 0);
    float *hann = (float *)&h[o_hann];
    hcpx *tw = (hcpx *)&h[o_tw], *pft = (hcpx *)&h[o_pft], *dphi = (hcpx *)&h[o_dphi], *back = (hcpx *)&h[o_back];
    uint16_t *perm = (uint16_t *)&h[o_perm];
    /* Hann table by oscillator recurrence, reference src/fsk.c:94-111 */
    {
        hcpx d = hcexpj((float)((2 * M_PI) / ((float)Ndft - 1)));
        hcpx r; r.r = .5f; r.i = 0;
        hcpx dc = d; dc.i = -dc.i;
        r = hcmul(dc, r);
        for (int i = 0; i < Ndft; i++) { r = hcmul(d, r); hann[i] = (float)(.5 - r.r); }
    }
    /* kiss_fft twiddles, reference src/kiss_fft.c:357-363 */
    for (int i = 0; i < Ndft; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / Ndft;
        tw[i].r = cosf((float)phase); tw[i].i = sinf((float)phase);
    }
    int fac[64];
    int nl = fft_factor(Ndft, fac);
    if (nl > WB_MAX_LEVELS) return wb_fail(WB_EINVAL, "FFT too deep");
    for (int l = 0; l < nl; l++) if (fac[2 * l] != 4 && fac[2 * l] != 2) return wb_fail(WB_EINVAL, "unsupported FFT radix");
    for (int l = 0; l < nl - 1; l++) if (fac[2 * l] != 4) return wb_fail(WB_EINVAL, "unsupported FFT factorisation");
    fft_perm(perm, 0, 0, 1, fac);
    /* the kernel's FFT schedule is compiled in: 256 = 4 x 4 x 4 x 4, leaf first (kiss_fft factors 4s first) */
    if (nl != 4) return wb_fail(WB_EINVAL, "FFT schedule of %d levels: the kernels are built for 4 x 4 x 4 x 4", nl);
    for (int l = 0; l < nl; l++) if (fac[2 * l] != 4) return wb_fail(WB_EINVAL, "FFT radix %d at level %d: the kernels are built for radix 4", fac[2 * l], l);
    /* fine-timing oscillator, reference src/fsk.c:858-873 */
    {
        hcpx d = hcexpj((float)(2 * M_PI * ((float)fp.Rs / (float)(fp.P * fp.Rs))));
        hcpx ph; ph.r = 1; ph.i = 0;
        for (int i = 0; i < fp.nint; i++) { pft[i] = ph; ((float *)fp.pftc4[0])[i] = ph.r; ((float *)fp.pftc4[1])[i] = ph.i; ph = hcmul(ph, d); }
        /* the float recurrence falls into an exactly periodic orbit (period P) after a few steps: the blocked kernel
           keeps one period of multipliers in registers when that holds from block WB_PFT_NT on (it does for P = 8, 10) */
        fp.pft_steady = (fp.step == 1);
        for (int i = (WB_PFT_NT + 1) * fp.P; i < fp.nint && fp.pft_steady; i++)
            if (memcmp(&pft[i], &pft[i - fp.P], sizeof(hcpx)) != 0) fp.pft_steady = 0;
        if (getenv("WB_FSK_PFT_GENERIC")) fp.pft_steady = 0;
    }
    /* per-bin tone oscillators, reference src/fsk.c:756-764, :671 */
    for (int b = 0; b < nh; b++) {
        float f = (float)b * ((float)fp.Fs / (float)Ndft);
        dphi[b] = hcexpj((float)(2 * M_PI * ((f) / (float)fp.Fs)));
        for (int k = 0; k < 3; k++) {
            int nin = fp.N + (k - 1) * (fp.Ts / 2);
            back[k * nh + b] = hcexpj((float)(-2 * (fp.Nmem - nin - fp.step) * M_PI * ((f) / (float)fp.Fs)));
        }
    }
    if (cudaMalloc(&e->d_tables, off) != cudaSuccess) return wb_fail(WB_ENOMEM, "cudaMalloc tables");
    CU(cudaMemcpy(e->d_tables, h.data(), off, cudaMemcpyHostToDevice));
    unsigned char *base = (unsigned char *)e->d_tables;
    fp.hann = (const float *)(base + o_hann); fp.tw = (const float2 *)(base + o_tw);
    fp.perm = (const uint16_t *)(base + o_perm);
    fp.dphi = (const float2 *)(base + o_dphi); fp.back = (const float2 *)(base + o_back);

    /* LDPC edge table: message words of (data column i, k-th check in H_cols order), k = 0..2, and i itself, listed
       by variable-pass slot: slot s works on column wb_vperm[s] (an order with few shared-memory bank conflicts,
       tools/gen_vperm.py; any permutation is correct) */
    {
        std::vector<char> seen(WB_NDATA, 0);
        for (int s = 0; s < WB_NDATA; s++) {
            if (wb_vperm[s] >= WB_NDATA || seen[wb_vperm[s]]) return wb_fail(WB_EINVAL, "wb_vperm is not a permutation");
            seen[wb_vperm[s]] = 1;
        }
    }
    std::vector<ushort4> vedge(WB_NDATA);
    for (int s = 0; s < WB_NDATA; s++) {
        const int i = wb_vperm[s];
        unsigned short w[3];
        for (int k = 0; k < WB_COLW; k++) {
            int c = wb_hcols[i * WB_COLW + k], found = -1;
            for (int t = 0; t < WB_ROWW; t++) if (wb_hrows[c * WB_ROWW + t] == i) { found = t; break; }
            if (found < 0) return wb_fail(WB_EINVAL, "H tables inconsistent");
            w[k] = (unsigned short)(found * WB_NPAR + c);
        }
        vedge[s] = make_ushort4(w[0], w[1], w[2], (unsigned short)i);
    }
    CU(cudaMalloc(&e->d_vedge, sizeof(ushort4) * WB_NDATA));
    CU(cudaMemcpy(e->d_vedge, vedge.data(), sizeof(ushort4) * WB_NDATA, cudaMemcpyHostToDevice));
    /* CRC as an affine map: crc(msg) = crc(0..0) ^ XOR_{set bits} T[bit] */
    std::vector<uint16_t> crc_tab(2048);
    {
        uint8_t msg[256];
        memset(msg, 0, sizeof(msg));
        e->crc0 = crc16_ccitt(msg, 256, 0xFFFF);
        for (int i = 0; i < 2048; i++) {
            msg[i >> 3] = (uint8_t)(0x80u >> (i & 7));
            crc_tab[i] = crc16_ccitt(msg, 256, 0x0000);
            msg[i >> 3] = 0;
        }
    }
    CU(cudaMalloc(&e->d_crc_tab, sizeof(uint16_t) * 2048));
    CU(cudaMemcpy(e->d_crc_tab, crc_tab.data(), sizeof(uint16_t) * 2048, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&e->d_scramble, WB_SCRAMBLE_LEN));
    CU(cudaMemcpy(e->d_scramble, wb_scramble_neg, WB_SCRAMBLE_LEN, cudaMemcpyHostToDevice));
    static wb_phi0_flag lut;
    if (wb_phi0_build_flag(&lut) != 0) return wb_fail(WB_EINVAL, "phi0 table: the step function does not fit the flag form");
    CU(cudaMalloc(&e->d_lut, sizeof(lut)));
    CU(cudaMemcpy(e->d_lut, &lut, sizeof(lut), cudaMemcpyHostToDevice));
    return WB_OK;
}

/* ---- lifecycle ---------------------------------------------------------- */

extern "C" int wb_abi_version(void) { return WB_ABI_VERSION; }
extern "C" const char *wb_last_error(void) { return g_err; }

extern "C" void wb_destroy(wb_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_in); cudaFree(e->d_state); cudaFree(e->d_cursor); cudaFree(e->d_sd); cudaFree(e->d_hard); cudaFree(e->d_jobs);
    cudaFree(e->d_c4); cudaFree(e->d_cw); cudaFree(e->d_cw_packed); cudaFree(e->d_llr); cudaFree(e->d_llr_packed);
    cudaFree(e->d_gather); cudaFree(e->d_fill); cudaFree(e->d_frame_log); cudaFree(e->d_tables); cudaFree(e->d_vedge);
    cudaFree(e->d_crc_tab); cudaFree(e->d_scramble); cudaFree(e->d_lut);
    cudaFree(e->d_tx_bits); cudaFree(e->d_hrows);
    cudaFree(e->d_bench_llr); cudaFree(e->d_bench_bits); cudaFree(e->d_bench_iters); cudaFree(e->d_bench_pcc);
    for (int i = 0; i < 8; i++) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    if (e->tev0) cudaEventDestroy(e->tev0);
    if (e->tev1) cudaEventDestroy(e->tev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int init_states(wb_engine *e)
{
    const int n = e->cfg.n_streams;
    std::vector<wb_stream_state> h(n);
    memset(h.data(), 0, sizeof(wb_stream_state) * n);
    for (int s = 0; s < n; s++) {
        for (int m = 0; m < WB_MAXM; m++) { h[s].phi_c[m].x = 1.0f; h[s].phi_c[m].y = 0.0f; }   /* comp_exp_j(0) */
        h[s].nin = e->fp.N;
        h[s].in_pos = h[s].in_fill = WB_HEADROOM;
    }
    CU(cudaMemcpy(e->d_state, h.data(), sizeof(wb_stream_state) * n, cudaMemcpyHostToDevice));
    CU(cudaMemset(e->d_cursor, 0, sizeof(wb_cursor) * n));
    for (int s = 0; s < n; s++) {
        memset(&e->cursor[s], 0, sizeof(wb_cursor));
        e->cursor[s].nin = e->fp.N;
        e->fill[s] = 0;
        e->nin[s] = e->fp.N;
        e->packets[s].clear();
    }
    e->uniform_fill = true;
    e->pending = false;
    e->resident_mode = false;
    return WB_OK;
}

template <int M, int TS, bool CF32, bool BLK, bool HARD>
static cudaError_t fsk_set_attr3(size_t smem)
{
    cudaError_t ce = cudaFuncSetAttribute(wb_fsk_kernel<M, TS, CF32, BLK, HARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    return cudaFuncSetAttribute(wb_fsk_kernel<M, TS, CF32, BLK, HARD>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}
template <int M, int TS, bool CF32, bool BLK>
static cudaError_t fsk_set_attr2(size_t smem)
{
    cudaError_t ce = fsk_set_attr3<M, TS, CF32, BLK, false>(smem);
    return ce != cudaSuccess ? ce : fsk_set_attr3<M, TS, CF32, BLK, true>(smem);
}
template <int M, int TS, bool CF32>
static cudaError_t fsk_set_attr(size_t smem, bool blk)
{
    return blk ? fsk_set_attr2<M, TS, CF32, true>(smem) : fsk_set_attr2<M, TS, CF32, false>(smem);
}

extern "C" int wb_create(const wb_config *cfg, wb_engine **out)
{
    if (!cfg || !out) return wb_fail(WB_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(wb_config)) return wb_fail(WB_EINVAL, "wb_config.struct_size mismatch (ABI %d)", WB_ABI_VERSION);
    if (cfg->n_streams <= 0 || cfg->n_streams > 65535 * 16) return wb_fail(WB_EINVAL, "n_streams out of range");
    if (cfg->framing < WB_FRAMING_NONE || cfg->framing > WB_FRAMING_V2) return wb_fail(WB_EINVAL, "unknown framing");
    if (cfg->framing != WB_FRAMING_NONE && cfg->M != 2) return wb_fail(WB_EINVAL, "framing needs 2-FSK (the reference has no 4-FSK deframer)");
    if (cfg->chunk_samples < 1024) return wb_fail(WB_EINVAL, "chunk_samples too small");
    if (cfg->chunk_samples > 0xF0000000ull) return wb_fail(WB_EINVAL, "chunk_samples too large (row positions are 32-bit)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return wb_fail(WB_ENODEV, "no CUDA device visible: libwenet_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return wb_fail(WB_EINVAL, "device %d of %d", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));

    wb_engine *e = new wb_engine();
    memset(&e->cfg, 0, sizeof(e->cfg));
    e->cfg = *cfg;
    e->stream = nullptr; e->tev0 = e->tev1 = nullptr;
    for (int i = 0; i < 8; i++) e->ev[i] = nullptr;
    e->d_in = nullptr; e->d_state = nullptr; e->d_cursor = nullptr; e->d_sd = nullptr; e->d_hard = nullptr; e->hard_valid = false; e->d_jobs = nullptr; e->d_c4 = nullptr;
    e->d_cw = e->d_cw_packed = nullptr; e->d_llr = e->d_llr_packed = nullptr; e->d_gather = nullptr; e->d_fill = nullptr; e->d_frame_log = nullptr;
    e->d_tx_bits = nullptr; e->d_hrows = nullptr; e->tx_bits_stride = e->tx_nbits = 0;
    e->d_tables = nullptr; e->d_vedge = nullptr; e->d_crc_tab = nullptr; e->d_scramble = nullptr; e->d_lut = nullptr;
    e->d_bench_llr = nullptr; e->d_bench_bits = nullptr; e->d_bench_iters = e->d_bench_pcc = nullptr; e->bench_n = 0;
    e->launches = 0; e->last_codewords = e->last_samples = 0; e->log_cap = 0;
    memset(e->kernel_ms, 0, sizeof(e->kernel_ms));
    int rc = build_fsk_params(cfg, &e->fp);
    if (rc) { delete e; return rc; }
    e->cfg.P = e->fp.P;
    e->max_iter = cfg->ldpc_max_iter > 0 ? cfg->ldpc_max_iter : WB_LDPC_MAX_ITER;
    /* deframer constants: reference src/drs232_ldpc.c:65-86 / src/wenet_ldpc.c:65-82 */
    {
        static const uint8_t uwb[4] = {0xAB, 0xCD, 0xEF, 0x01};
        uint8_t uw[40]; int nb = 0;
        wb_deframe_params &dp = e->dp;
        memset(&dp, 0, sizeof(dp));
        dp.mode = cfg->framing;
        if (cfg->framing == WB_FRAMING_V2) {
            for (int b = 0; b < 4; b++) for (int k = 7; k >= 0; k--) uw[nb++] = (uwb[b] >> k) & 1;
            dp.uw_thresh = 28; dp.nsym = WB_PKT_BODY_BYTES * 8;
        } else {
            for (int b = 0; b < 4; b++) { uw[nb++] = 0; for (int k = 0; k < 8; k++) uw[nb++] = (uwb[b] >> k) & 1; uw[nb++] = 1; }
            dp.uw_thresh = 35; dp.nsym = WB_PKT_BODY_BYTES * 10;
        }
        dp.uw_bits = nb;
        dp.uw = 0;
        for (int i = 0; i < nb; i++) dp.uw |= (unsigned long long)uw[i] << (nb - 1 - i);   /* oldest bit highest */
        dp.uw_mask = (nb == 64) ? ~0ULL : ((1ULL << nb) - 1);
    }
    const int n = cfg->n_streams;
    e->cursor.resize(n); e->fill.resize(n); e->nin.resize(n); e->packets.resize(n);

#define CRE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        wb_fail(e_ == cudaErrorMemoryAllocation ? WB_ENOMEM : WB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
        wb_destroy(e); return e_ == cudaErrorMemoryAllocation ? WB_ENOMEM : WB_ECUDA; } } while (0)

    CRE(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) CRE(cudaEventCreate(&e->ev[i]));
    CRE(cudaEventCreate(&e->tev0)); CRE(cudaEventCreate(&e->tev1));
    e->in_cap = WB_HEADROOM + cfg->chunk_samples;
    e->in_stride = ((e->in_cap * e->fp.in_bps + 64) + 255) & ~255ULL;
    unsigned long long max_frames = e->in_cap / (unsigned)(e->fp.N - e->fp.Ts / 2) + 2;
    e->sd_cap = (unsigned)(max_frames * e->fp.Nbits);
    e->sd_stride = (WB_CARRY_CAP + (unsigned long long)e->sd_cap + 63) & ~63ULL;
    e->job_cap = (cfg->framing == WB_FRAMING_NONE) ? 1 : (int)((WB_CARRY_CAP + e->sd_cap) / e->dp.nsym + 2);
    CRE(cudaMalloc(&e->d_in, e->in_stride * n));
    CRE(cudaMalloc(&e->d_state, sizeof(wb_stream_state) * n));
    CRE(cudaMalloc(&e->d_cursor, sizeof(wb_cursor) * n));
    CRE(cudaMalloc(&e->d_sd, sizeof(float) * e->sd_stride * n));
    CRE(cudaMemsetAsync(e->d_sd, 0, sizeof(float) * e->sd_stride * n, e->stream));
    if (cfg->flags & WB_FLAG_HARD_BITS) {
        CRE(cudaMalloc(&e->d_hard, (WB_HARD_PRE + (size_t)e->sd_cap) * n));
        CRE(cudaMemsetAsync(e->d_hard, 0, (WB_HARD_PRE + (size_t)e->sd_cap) * n, e->stream));
    }
    size_t nslots = (size_t)n * e->job_cap;
    CRE(cudaMalloc(&e->d_jobs, sizeof(unsigned) * nslots));
    CRE(cudaMalloc(&e->d_c4, sizeof(double) * nslots));
    CRE(cudaMalloc(&e->d_cw, sizeof(wb_codeword) * nslots));
    CRE(cudaMalloc(&e->d_cw_packed, sizeof(wb_codeword) * nslots));
    CRE(cudaMalloc(&e->d_gather, sizeof(unsigned) * (nslots + (size_t)n + 1)));
    CRE(cudaMalloc(&e->d_fill, sizeof(unsigned long long) * n));
    if (cfg->flags & WB_FLAG_KEEP_LLR) {
        CRE(cudaMalloc(&e->d_llr, sizeof(float) * WB_NCODE * nslots));
        CRE(cudaMalloc(&e->d_llr_packed, sizeof(float) * WB_NCODE * nslots));
    }
    rc = upload_tables(e);
    if (rc) { wb_destroy(e); return rc; }
    rc = init_states(e);
    if (rc) { wb_destroy(e); return rc; }
    /* kernel attributes: streams per CTA so that two CTAs share an SM (one in a sequential phase while the
       other runs a parallel one) */
    {
        const int max_spb = (e->fp.M == 2) ? 14 : 8;           /* launch bounds of wb_fsk_kernel */
        int dev_smem_sm = 0, dev_smem_blk = 0;
        CRE(cudaDeviceGetAttribute(&dev_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cfg->device));
        CRE(cudaDeviceGetAttribute(&dev_smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
        auto smem_for = [&](int spb) { return (size_t)wb_geom_head(spb, (int)sizeof(wb_fsk_sc)) + (size_t)spb * e->fp.sreg; };
        int spb = 0;
        for (int ctas = 2; ctas >= 1 && spb == 0; ctas--)
            for (int t = max_spb; t >= (ctas == 2 ? 4 : 2); t--)      /* >= 2: the re and im fine-timing chains run in warps 0 and 1 */
                if (smem_for(t) <= (size_t)dev_smem_blk && ctas * (smem_for(t) + 1024) <= (size_t)dev_smem_sm) { spb = t; break; }
        if (spb == 0) { wb_destroy(e); return wb_fail(WB_EINVAL, "FSK kernel does not fit shared memory"); }
        /* fewer streams than one full CTA per SM: spread them, one smaller CTA on every SM instead of full CTAs on some
           of them (1024 streams on a B200: 74 CTAs x 14 streams 34.8 ms per 1 Mi-sample chunk, 147 x 7 28.2 ms; 512
           streams: 37 x 14 34.7 ms, 128 x 4 26.7 ms; from one CTA per SM up the full CTA wins: 2048 streams 35.1 vs
           36.9 ms; tools/sweep_spb.sh) */
        {
            int n_sm = 0;
            CRE(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
            if (n_sm > 0 && (n + spb - 1) / spb < n_sm) spb = std::min(spb, std::max(4, (n + n_sm - 1) / n_sm));
        }
        if (getenv("WB_FSK_SPB")) spb = std::max(2, std::min(max_spb, atoi(getenv("WB_FSK_SPB"))));
        e->spb = spb;
        /* segments of the sequential mixer phase (see wb_fsk_kernel.cuh, B1): multiples of 8 steps, shrinking
           geometrically by the cost ratio (bare recurrence step) / (recurrence + mix step) */
        {
            int W = std::min(spb, 4);
            if (getenv("WB_FSK_B1W")) W = std::max(1, std::min(std::min(spb, WB_MAX_B1W), atoi(getenv("WB_FSK_B1W"))));
            double r = getenv("WB_FSK_B1R") ? atof(getenv("WB_FSK_B1R")) : 0.7;
            const int nsteps = e->fp.nsteps;
            double tot = 0, wgt = 1;
            for (int j = 0; j < W; j++) { tot += wgt; wgt *= r; }
            int acc = 0; wgt = 1;
            e->fp.b1_seg[0] = 0;
            for (int j = 0; j < W; j++) {
                int len = (int)(nsteps * wgt / tot / 8.0 + 0.5) * 8;
                if (len < 8) len = 8;
                acc += len; wgt *= r;
                if (acc > nsteps || j == W - 1) acc = nsteps;
                e->fp.b1_seg[j + 1] = acc;
                if (acc == nsteps) { W = j + 1; break; }
            }
            e->fp.b1_w = W;
            /* experiments: explicit segment ends, e.g. WB_FSK_B1SEG=160,272,352 (multiples of 8, ascending, < nsteps) */
            if (const char *sg = getenv("WB_FSK_B1SEG")) {
                int w = 0, prev = 0;
                bool ok = true;
                while (*sg && w < WB_MAX_B1W - 1) {
                    char *end = nullptr;
                    const long v = strtol(sg, &end, 10);
                    if (end == sg) break;
                    if (v <= prev || v >= nsteps || v % 8) { ok = false; break; }
                    e->fp.b1_seg[++w] = (int)v; prev = (int)v;
                    sg = (*end == ',') ? end + 1 : end;
                }
                if (ok && w + 1 <= spb) { e->fp.b1_seg[w + 1] = nsteps; e->fp.b1_w = w + 1; }
                else { wb_destroy(e); return wb_fail(WB_EINVAL, "WB_FSK_B1SEG: ascending multiples of 8 below the frame's step count, at most one segment per stream of a CTA"); }
            }
        }
        e->fsk_smem = smem_for(spb);
        const bool cf32 = e->fp.in_fmt == WB_FMT_CF32, blk = e->fp.step == 1;
        cudaError_t ce = cudaErrorInvalidValue;
        if (e->fp.M == 2 && e->fp.Ts == 8) ce = cf32 ? fsk_set_attr<2, 8, true>(e->fsk_smem, blk) : fsk_set_attr<2, 8, false>(e->fsk_smem, blk);
        else if (e->fp.M == 2 && e->fp.Ts == 10) ce = cf32 ? fsk_set_attr<2, 10, true>(e->fsk_smem, blk) : fsk_set_attr<2, 10, false>(e->fsk_smem, blk);
        else if (e->fp.M == 4 && e->fp.Ts == 8) ce = cf32 ? fsk_set_attr<4, 8, true>(e->fsk_smem, blk) : fsk_set_attr<4, 8, false>(e->fsk_smem, blk);
        else if (e->fp.M == 4 && e->fp.Ts == 10) ce = cf32 ? fsk_set_attr<4, 10, true>(e->fsk_smem, blk) : fsk_set_attr<4, 10, false>(e->fsk_smem, blk);
        else { wb_destroy(e); return wb_fail(WB_EINVAL, "Fs/Rs = %d: only 8 and 10 samples per symbol are built", e->fp.Ts); }
        if (ce != cudaSuccess) { wb_fail(WB_ECUDA, "fsk kernel smem %zu: %s", e->fsk_smem, cudaGetErrorString(ce)); wb_destroy(e); return WB_ECUDA; }
        CRE(cudaFuncSetAttribute(wb_ldpc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(wb_ldpc_smem)));
    }
    CRE(cudaStreamSynchronize(e->stream));
#undef CRE
    *out = e;
    return WB_OK;
}

/* ---- streaming path ----------------------------------------------------- */

static int wb_collect(wb_engine *e);

extern "C" int wb_nin(wb_engine *e, uint32_t *nin)
{
    if (!e || !nin) return wb_fail(WB_EINVAL, "null argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    for (int s = 0; s < e->cfg.n_streams; s++) nin[s] = (uint32_t)e->nin[s];
    return WB_OK;
}

extern "C" int wb_feed(wb_engine *e, const void *const *iq, const uint64_t *nsamp)
{
    if (!e || !nsamp) return wb_fail(WB_EINVAL, "null argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    if (e->resident_mode) return wb_fail(WB_EINVAL, "engine is in resident (wb_dev_set_fill) mode");
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams, bps = e->fp.in_bps;
    for (int s = 0; s < n; s++) {
        if (nsamp[s] == 0) continue;
        if (!iq || !iq[s]) return wb_fail(WB_EINVAL, "stream %d: null buffer", s);
        if (WB_HEADROOM + e->fill[s] + nsamp[s] > e->in_cap) return wb_fail(WB_ERANGE, "stream %d: chunk capacity exceeded", s);
    }
    /* the remainder of stream s occupies [HEADROOM - rem, HEADROOM) after a compacting wb_process, or
       [HEADROOM, HEADROOM + fill) if nothing was processed since the last feed: fill[] counts from the
       parked remainder's start in both cases */
    for (int s = 0; s < n; s++) {
        if (nsamp[s] == 0) continue;
        unsigned long long rem = e->cursor[s].in_fill;       /* parked remainder (0 before the first process) */
        unsigned long long off = WB_HEADROOM + (e->fill[s] - rem);
        CU(cudaMemcpyAsync(e->d_in + (size_t)s * e->in_stride + off * bps, iq[s], nsamp[s] * bps, cudaMemcpyHostToDevice, e->stream));
        e->fill[s] += nsamp[s];
    }
    for (int s = 1; s < n && e->uniform_fill; s++)
        if (e->fill[s] - e->cursor[s].in_fill != e->fill[0] - e->cursor[0].in_fill) e->uniform_fill = false;
    return WB_OK;
}

extern "C" int wb_feed_strided(wb_engine *e, const void *base, uint64_t stride_bytes, uint64_t nsamp)
{
    if (!e || !base) return wb_fail(WB_EINVAL, "null argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    if (e->resident_mode) return wb_fail(WB_EINVAL, "engine is in resident (wb_dev_set_fill) mode");
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams, bps = e->fp.in_bps;
    if (nsamp == 0) return WB_OK;
    /* fresh samples of every stream start at the same row offset iff (fill - parked remainder) is uniform */
    unsigned long long fresh0 = e->fill[0] - e->cursor[0].in_fill;
    bool uniform = true;
    for (int s = 0; s < n; s++) {
        if (WB_HEADROOM + e->fill[s] + nsamp > e->in_cap) return wb_fail(WB_ERANGE, "stream %d: chunk capacity exceeded", s);
        if (e->fill[s] - e->cursor[s].in_fill != fresh0) uniform = false;
    }
    if (uniform) {
        CU(cudaMemcpy2DAsync(e->d_in + (WB_HEADROOM + fresh0) * bps, e->in_stride, base, stride_bytes, nsamp * bps, n,
                             cudaMemcpyHostToDevice, e->stream));
        for (int s = 0; s < n; s++) e->fill[s] += nsamp;
    } else {
        for (int s = 0; s < n; s++) {
            unsigned long long off = WB_HEADROOM + (e->fill[s] - e->cursor[s].in_fill);
            CU(cudaMemcpyAsync(e->d_in + (size_t)s * e->in_stride + off * bps, (const unsigned char *)base + (size_t)s * stride_bytes,
                               nsamp * bps, cudaMemcpyHostToDevice, e->stream));
            e->fill[s] += nsamp;
        }
    }
    return WB_OK;
}

__global__ void wb_set_fill_kernel(wb_stream_state *st, const unsigned long long *fill, int n)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    /* in_pos stays where the last process left it (headroom - parked remainder) */
    st[s].in_fill = st[s].in_pos + fill[s];
}

__global__ void wb_gather_kernel(const wb_codeword *cw, const float *llr, const unsigned *offs, const wb_cursor *cur,
                                 wb_codeword *cw_out, float *llr_out, int job_cap, int n_streams)
{
    /* one block per stream: copy its n_jobs records (and LLR rows) to the packed arrays */
    int s = blockIdx.x;
    unsigned nj = cur[s].n_jobs, o = offs[s];
    const unsigned *src = reinterpret_cast<const unsigned *>(cw + (size_t)s * job_cap);
    unsigned *dst = reinterpret_cast<unsigned *>(cw_out + o);
    const unsigned words = nj * (unsigned)(sizeof(wb_codeword) / 4);
    for (unsigned i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    if (llr && llr_out) {
        const float *ls = llr + (size_t)s * job_cap * WB_NCODE;
        float *ld = llr_out + (size_t)o * WB_NCODE;
        for (unsigned i = threadIdx.x; i < nj * WB_NCODE; i += blockDim.x) ld[i] = ls[i];
    }
}

template <int M, int TS, bool HARD>
static void launch_fsk2(wb_engine *e, const wb_fsk_args &a)
{
    const int grid = (e->cfg.n_streams + e->spb - 1) / e->spb;
    const bool cf32 = e->fp.in_fmt == WB_FMT_CF32, blk = e->fp.step == 1;
    if (cf32 && blk) wb_fsk_kernel<M, TS, true, true, HARD><<<grid, e->spb * 32, e->fsk_smem, e->stream>>>(e->fp, a);
    else if (cf32) wb_fsk_kernel<M, TS, true, false, HARD><<<grid, e->spb * 32, e->fsk_smem, e->stream>>>(e->fp, a);
    else if (blk) wb_fsk_kernel<M, TS, false, true, HARD><<<grid, e->spb * 32, e->fsk_smem, e->stream>>>(e->fp, a);
    else wb_fsk_kernel<M, TS, false, false, HARD><<<grid, e->spb * 32, e->fsk_smem, e->stream>>>(e->fp, a);
}
/* WB_FLAG_HARD_BITS picks the instantiation that also writes rx_bits */
template <int M, int TS>
static void launch_fsk(wb_engine *e, const wb_fsk_args &a)
{
    if (a.hard) launch_fsk2<M, TS, true>(e, a);
    else launch_fsk2<M, TS, false>(e, a);
}

/* K2 -> K3 -> K4 -> carry over whatever soft decisions the rows hold (cursor[s].n_sd of them per stream) */
static int launch_decode(wb_engine *e)
{
    const int n = e->cfg.n_streams;
    if (e->cfg.framing != WB_FRAMING_NONE) {
        wb_deframe_kernel<<<(n + 3) / 4, 128, 0, e->stream>>>(e->dp, e->d_state, e->d_cursor, e->d_sd, e->sd_stride,
                                                               e->d_jobs, e->job_cap, n);
        CU(cudaEventRecord(e->ev[2], e->stream));
        size_t nslots = (size_t)n * e->job_cap;
        wb_llr_stats_kernel<<<(unsigned)((nslots + 127) / 128), 128, 0, e->stream>>>(e->d_sd, e->sd_stride, e->d_jobs, e->d_c4,
                                                                                    e->d_cursor, e->job_cap, n, e->cfg.framing, e->d_scramble);
        CU(cudaEventRecord(e->ev[3], e->stream));
        wb_ldpc_args la;
        memset(&la, 0, sizeof(la));
        la.sd = e->d_sd; la.sd_stride = e->sd_stride; la.jobs = e->d_jobs; la.c4 = e->d_c4; la.cur = e->d_cursor;
        la.st = e->d_state; la.job_cap = e->job_cap; la.framing = e->cfg.framing; la.llr_in = nullptr;
        la.cw = e->d_cw; la.llr_out = e->d_llr; la.max_iter = e->max_iter;
        la.vedge = e->d_vedge; la.crc_tab = e->d_crc_tab; la.crc0 = e->crc0; la.scramble = e->d_scramble; la.lut = e->d_lut;
        wb_ldpc_kernel<<<(unsigned)nslots, WB_LDPC_THREADS, sizeof(wb_ldpc_smem), e->stream>>>(la, 0);
        CU(cudaEventRecord(e->ev[4], e->stream));
        wb_carry_kernel<<<(n + 3) / 4, 128, 0, e->stream>>>(e->d_state, e->d_cursor, e->d_sd, e->sd_stride, n);
        e->launches += 4;
    } else {
        CU(cudaEventRecord(e->ev[2], e->stream));
        CU(cudaEventRecord(e->ev[3], e->stream));
        CU(cudaEventRecord(e->ev[4], e->stream));
    }
    return WB_OK;
}

extern "C" int wb_process(wb_engine *e)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    /* resident (benchmark) mode queues passes back to back: only the last one's records are collected */
    int rc = e->resident_mode ? WB_OK : wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams;
    if (!e->resident_mode) {
        /* tell the device how many samples each stream now holds (fill counts from the parked remainder);
           e->fill is pageable, so the copy is staged before the call returns */
        CU(cudaMemcpyAsync(e->d_fill, e->fill.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, e->stream));
        wb_set_fill_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(e->d_state, e->d_fill, n);
        e->launches++;
    }
    wb_fsk_args a;
    a.state = e->d_state; a.cursor = e->d_cursor; a.in = e->d_in; a.in_stride = e->in_stride;
    a.sd = e->d_sd; a.sd_stride = e->sd_stride; a.sd_cap = e->sd_cap; a.n_streams = n;
    a.compact = e->resident_mode ? 0 : 1; a.headroom = WB_HEADROOM; a.spb = e->spb;
    a.frame_log = e->d_frame_log; a.log_cap = e->log_cap; a.hard = e->d_hard;
    CU(cudaEventRecord(e->ev[0], e->stream));
    if (e->fp.M == 2 && e->fp.Ts == 8) launch_fsk<2, 8>(e, a);
    else if (e->fp.M == 2 && e->fp.Ts == 10) launch_fsk<2, 10>(e, a);
    else if (e->fp.M == 4 && e->fp.Ts == 8) launch_fsk<4, 8>(e, a);
    else launch_fsk<4, 10>(e, a);
    e->launches++;
    e->hard_valid = true;
    CU(cudaEventRecord(e->ev[1], e->stream));
    int rc2 = launch_decode(e);
    if (rc2) return rc2;
    CU(cudaGetLastError());
    e->pending = true;
    return WB_OK;
}


__global__ void wb_set_nsd_kernel(wb_cursor *cur, const unsigned long long *nsd, const wb_stream_state *st, int n)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    cur[s].n_sd = (unsigned)nsd[s];
    cur[s].consumed = 0;
    cur[s].n_jobs = 0;
    cur[s].nin = st[s].nin;
    /* in_fill (the parked IQ remainder) is left as it is */
}

/* the fread(&symbol) loop of drs232_ldpc.c:176 / wenet_ldpc.c:171 without a demodulator in front */
extern "C" int wb_process_soft(wb_engine *e, const float *const *sd, const uint64_t *nsym)
{
    if (!e || !sd || !nsym) return wb_fail(WB_EINVAL, "null argument");
    if (e->cfg.framing == WB_FRAMING_NONE) return wb_fail(WB_EINVAL, "engine was created without a framing");
    if (e->resident_mode) return wb_fail(WB_EINVAL, "engine is in resident (wb_dev_set_fill) mode");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams;
    std::vector<unsigned long long> cnt(n);
    for (int s = 0; s < n; s++) {
        if (nsym[s] > e->sd_cap) return wb_fail(WB_ERANGE, "stream %d: more than %u soft symbols per call", s, e->sd_cap);
        if (nsym[s] && !sd[s]) return wb_fail(WB_EINVAL, "stream %d: null buffer", s);
        cnt[s] = nsym[s];
    }
    for (int s = 0; s < n; s++)
        if (nsym[s])
            CU(cudaMemcpyAsync(e->d_sd + (size_t)s * e->sd_stride + WB_CARRY_CAP, sd[s], sizeof(float) * nsym[s],
                               cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_fill, cnt.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, e->stream));
    wb_set_nsd_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(e->d_cursor, e->d_fill, e->d_state, n);
    e->launches++;
    e->hard_valid = false;
    CU(cudaEventRecord(e->ev[0], e->stream));
    CU(cudaEventRecord(e->ev[1], e->stream));
    rc = launch_decode(e);
    if (rc) return rc;
    CU(cudaGetLastError());
    e->pending = true;
    return WB_OK;
}

/* wait for the last wb_process and bring its bookkeeping (cursors, codeword records) to the host */
static int wb_collect(wb_engine *e)
{
    if (!e->pending) return WB_OK;
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams;
    CU(cudaMemcpyAsync(e->cursor.data(), e->d_cursor, sizeof(wb_cursor) * n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->pending = false;
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&e->kernel_ms[i], e->ev[i], e->ev[i + 1]);
    uint64_t total = 0, samples = 0;
    std::vector<unsigned> offs(n + 1);
    for (int s = 0; s < n; s++) {
        offs[s] = (unsigned)total;
        if (e->cfg.framing == WB_FRAMING_NONE) e->cursor[s].n_jobs = 0;
        total += e->cursor[s].n_jobs;
        samples += e->cursor[s].consumed;
        /* only a pass that ran the demodulator consumed IQ samples: after wb_process_soft the samples fed since the last
           wb_process are still waiting in the row (cursor.in_fill is the OLD parked remainder there) */
        if (!e->resident_mode && e->hard_valid) e->fill[s] = e->cursor[s].in_fill;
        e->nin[s] = e->cursor[s].nin;
    }
    offs[n] = (unsigned)total;
    e->last_codewords = total;
    e->last_samples = samples;
    e->last_cw.resize(total);
    e->last_llr.clear();
    if (total) {
        CU(cudaMemcpyAsync(e->d_gather, offs.data(), sizeof(unsigned) * (n + 1), cudaMemcpyHostToDevice, e->stream));
        wb_gather_kernel<<<n, 256, 0, e->stream>>>(e->d_cw, e->d_llr, e->d_gather, e->d_cursor, e->d_cw_packed, e->d_llr_packed,
                                                    e->job_cap, n);
        e->launches++;
        CU(cudaMemcpyAsync(e->last_cw.data(), e->d_cw_packed, sizeof(wb_codeword) * total, cudaMemcpyDeviceToHost, e->stream));
        if (e->d_llr) {
            e->last_llr.resize((size_t)total * WB_NCODE);
            CU(cudaMemcpyAsync(e->last_llr.data(), e->d_llr_packed, sizeof(float) * WB_NCODE * total, cudaMemcpyDeviceToHost, e->stream));
        }
        CU(cudaStreamSynchronize(e->stream));
        for (size_t i = 0; i < total; i++) {
            const wb_codeword &c = e->last_cw[i];
            if (c.crc_ok) {
                auto &q = e->packets[c.stream];
                q.buf.insert(q.buf.end(), c.bytes, c.bytes + WB_PACKET_BYTES);
                q.seqs.push_back(c.seq);
            }
        }
    }
    return WB_OK;
}

extern "C" int wb_sync(wb_engine *e)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaStreamSynchronize(e->stream));
    return WB_OK;
}

extern "C" int wb_drain_packets(wb_engine *e, int stream, uint8_t *buf, size_t cap, size_t *nbytes)
{
    if (!e || !nbytes) return wb_fail(WB_EINVAL, "null argument");
    if (stream < 0 || stream >= e->cfg.n_streams) return wb_fail(WB_EINVAL, "stream %d out of range", stream);
    int rc = wb_collect(e);
    if (rc) return rc;
    wb_engine::pktq &q = e->packets[stream];
    size_t n = std::min(q.size(), cap - cap % WB_PACKET_BYTES);
    if (n && !buf) return wb_fail(WB_EINVAL, "null buffer");
    if (n) memcpy(buf, q.buf.data() + q.rd, n);
    q.rd += n;
    if (q.rd == q.buf.size()) q.clear();
    *nbytes = n;
    return WB_OK;
}

extern "C" int wb_drain_all_packets(wb_engine *e, uint8_t *buf, size_t cap, size_t *nbytes, uint64_t *npackets)
{
    if (!e || !nbytes) return wb_fail(WB_EINVAL, "null argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    const size_t rec = 8 + WB_PACKET_BYTES;
    size_t need = 0;
    for (auto &q : e->packets) need += q.size() / WB_PACKET_BYTES * rec;
    if (need > cap) { *nbytes = need; return wb_fail(WB_ERANGE, "need %zu bytes", need); }
    if (need && !buf) return wb_fail(WB_EINVAL, "null buffer");
    size_t o = 0; uint64_t np = 0;
    for (int s = 0; s < e->cfg.n_streams; s++) {
        wb_engine::pktq &q = e->packets[s];
        while (q.size() >= WB_PACKET_BYTES) {
            const int32_t ss = s;
            const uint32_t k = q.seqs[q.rd / WB_PACKET_BYTES];     /* = wb_codeword.seq of the codeword it came from */
            memcpy(buf + o, &ss, 4); memcpy(buf + o + 4, &k, 4);
            memcpy(buf + o + 8, q.buf.data() + q.rd, WB_PACKET_BYTES);
            q.rd += WB_PACKET_BYTES;
            o += rec; np++;
        }
        q.clear();
    }
    *nbytes = o;
    if (npackets) *npackets = np;
    return WB_OK;
}

extern "C" int wb_drain_soft(wb_engine *e, int stream, float *buf, size_t cap_floats, size_t *nout)
{
    if (!e || !nout) return wb_fail(WB_EINVAL, "null argument");
    if (stream < 0 || stream >= e->cfg.n_streams) return wb_fail(WB_EINVAL, "stream %d out of range", stream);
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    size_t n = e->cursor[stream].n_sd;
    *nout = n;
    if (n > cap_floats) return wb_fail(WB_ERANGE, "need %zu floats", n);
    if (n) {
        CU(cudaMemcpyAsync(buf, e->d_sd + (size_t)stream * e->sd_stride + WB_CARRY_CAP, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return WB_OK;
}

extern "C" int wb_drain_hard(wb_engine *e, int stream, uint8_t *buf, size_t cap, size_t *nout)
{
    if (!e || !nout) return wb_fail(WB_EINVAL, "null argument");
    if (stream < 0 || stream >= e->cfg.n_streams) return wb_fail(WB_EINVAL, "stream %d out of range", stream);
    if (!e->d_hard) return wb_fail(WB_EINVAL, "hard bits need WB_FLAG_HARD_BITS");
    if (!e->hard_valid) return wb_fail(WB_EINVAL, "no demodulator ran in the last process call");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    size_t n = e->cursor[stream].n_sd;
    *nout = n;
    if (n > cap) return wb_fail(WB_ERANGE, "need %zu bytes", n);
    if (n && !buf) return wb_fail(WB_EINVAL, "null buffer");
    if (n) {
        CU(cudaMemcpyAsync(buf, e->d_hard + (size_t)stream * (WB_HARD_PRE + (size_t)e->sd_cap) + WB_HARD_PRE, n, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return WB_OK;
}

extern "C" int wb_drain_codewords(wb_engine *e, wb_codeword *cw, float *llr, size_t cap, size_t *nout)
{
    if (!e || !nout) return wb_fail(WB_EINVAL, "null argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    size_t n = e->last_cw.size();
    *nout = n;
    if (n > cap) return wb_fail(WB_ERANGE, "need %zu records", n);
    if (n && cw) memcpy(cw, e->last_cw.data(), sizeof(wb_codeword) * n);
    if (n && llr) {
        if (e->last_llr.empty()) return wb_fail(WB_EINVAL, "LLRs need WB_FLAG_KEEP_LLR");
        memcpy(llr, e->last_llr.data(), sizeof(float) * WB_NCODE * n);
    }
    return WB_OK;
}

extern "C" int wb_get_stats(wb_engine *e, int stream, wb_stats *out)
{
    if (!e || !out) return wb_fail(WB_EINVAL, "null argument");
    if (stream < 0 || stream >= e->cfg.n_streams) return wb_fail(WB_EINVAL, "stream %d out of range", stream);
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    wb_stream_state st;
    CU(cudaMemcpyAsync(&st, e->d_state + stream, sizeof(st), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memset(out, 0, sizeof(*out));
    const float binw = (float)e->fp.Fs / (float)e->fp.Ndft;
    const int M = e->fp.M, P = e->fp.P;
    out->ppm = st.ppm;
    for (int m = 0; m < M; m++) out->f_est[m] = (float)st.fbin[m] * binw;   /* reference src/fsk.c:671 */
    out->rx_timing = st.rx_timing;
    out->norm_rx_timing = st.norm_rx_timing;
    {   /* reference src/fsk.c:1026-1029 with f1_tx = 1200, fs_tx = 400 (src/fsk_demod.c:214) */
        float fc_avg = (out->f_est[0] + out->f_est[1]) / 2;
        float fc_tx = (float)((1200 + 1200 + 400) / 2);
        out->foff = fc_tx - fc_avg;
    }
    out->nin = st.nin;
    out->nfft = e->fp.Ndft / 2;
    memcpy(out->samp_fft, st.fft_est, sizeof(float) * (e->fp.Ndft / 2));
    if (e->fp.stats && st.eb_count) {
        /* EbNodB and the snr_est IIR (reference src/fsk.c:1010, :1024), replayed over the last <= 32 frames with the
           host libm; older frames weigh less than 2^-32 */
        unsigned n = st.eb_count < 32 ? st.eb_count : 32;
        float snr = 0.0f;
        for (unsigned k = 0; k < n; k++) {
            float a = st.eb_arg[(st.eb_count - n + k) & 31];
            float ebnodb = -6 + (20 * log10f(a));
            snr = .5 * snr + .5 * ebnodb;
        }
        out->EbNodB = snr;
        /* eye diagram, reference src/fsk.c:1037-1077.  (The reference indexes f_int with high_sample + 1 + ..., which
           is negative for rx_timing < -1 and reads outside its array there; those samples are 0 here.) */
        int dec = (int)ceil(((float)P * 2) / 160);
        int neyesamp = (P * 2) / dec, traces = 8 / M, off = st.eye_high + 1;
        out->neyesamp = neyesamp; out->neyetr = M * traces;
        float emax = 0;
        for (int i = 0; i < traces; i++)
            for (int m = 0; m < M; m++)
                for (int j = 0; j < neyesamp; j++) {
                    int ind = 2 * P * i + off + j * dec;
                    float v = 0.0f;
                    if (ind >= 0 && ind < WB_EYE_KEEP) v = sqrtf(powf(st.eye_fint[m][ind].x, 2.0) + powf(st.eye_fint[m][ind].y, 2.0));
                    out->rx_eye[i * M + m][j] = v;
                    if (fabsf(v) > emax) emax = fabsf(v);
                }
        for (int i = 0; i < M * traces; i++)
            for (int j = 0; j < neyesamp; j++) out->rx_eye[i][j] = out->rx_eye[i][j] / emax;
    }
    out->frames = st.frames;
    out->packets = st.packets;
    out->packet_errors = st.packet_errors;
    return WB_OK;
}

extern "C" int wb_clear_estimators(wb_engine *e)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    int rc = wb_collect(e);
    if (rc) return rc;
    /* reference src/fsk.c:505-520: zero fft_est and re-arm the first-frame f_est latch */
    const int n = e->cfg.n_streams;
    std::vector<wb_stream_state> h(n);
    CU(cudaMemcpy(h.data(), e->d_state, sizeof(wb_stream_state) * n, cudaMemcpyDeviceToHost));
    for (int s = 0; s < n; s++) {
        memset(h[s].fft_est, 0, sizeof(h[s].fft_est));
        for (int m = 0; m < WB_MAXM; m++) h[s].fbin[m] = 0;
    }
    CU(cudaMemcpy(e->d_state, h.data(), sizeof(wb_stream_state) * n, cudaMemcpyHostToDevice));
    return WB_OK;
}

/* ---- stage-level entry points ------------------------------------------- */

extern "C" int wb_ldpc_decode_batch(wb_engine *e, const float *llr, size_t n, int max_iter, uint8_t *bits_packed,
                                    int32_t *iters, int32_t *parity_ok)
{
    if (!e || !llr || !bits_packed || !iters || !parity_ok) return wb_fail(WB_EINVAL, "null argument");
    if (n == 0) return WB_OK;
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    float *d_llr = nullptr; uint8_t *d_bits = nullptr; int *d_it = nullptr, *d_pc = nullptr;
    cudaError_t ce;
    if ((ce = cudaMalloc(&d_llr, sizeof(float) * WB_NCODE * n)) != cudaSuccess ||
        (ce = cudaMalloc(&d_bits, 323 * n)) != cudaSuccess || (ce = cudaMalloc(&d_it, 4 * n)) != cudaSuccess ||
        (ce = cudaMalloc(&d_pc, 4 * n)) != cudaSuccess) {
        cudaFree(d_llr); cudaFree(d_bits); cudaFree(d_it); cudaFree(d_pc);
        return wb_fail(WB_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(ce));
    }
    rc = WB_OK;
    do {
        if ((ce = cudaMemcpyAsync(d_llr, llr, sizeof(float) * WB_NCODE * n, cudaMemcpyHostToDevice, e->stream)) != cudaSuccess) break;
        wb_ldpc_args la;
        memset(&la, 0, sizeof(la));
        la.llr_in = d_llr; la.bits_out = d_bits; la.iters_out = d_it; la.pcc_out = d_pc;
        la.max_iter = max_iter > 0 ? max_iter : e->max_iter;
        la.vedge = e->d_vedge; la.crc_tab = e->d_crc_tab; la.crc0 = e->crc0; la.scramble = e->d_scramble; la.lut = e->d_lut;
        wb_ldpc_kernel<<<(unsigned)n, WB_LDPC_THREADS, sizeof(wb_ldpc_smem), e->stream>>>(la, (long long)n);
        e->launches++;
        if ((ce = cudaGetLastError()) != cudaSuccess) break;
        if ((ce = cudaMemcpyAsync(bits_packed, d_bits, 323 * n, cudaMemcpyDeviceToHost, e->stream)) != cudaSuccess) break;
        if ((ce = cudaMemcpyAsync(iters, d_it, 4 * n, cudaMemcpyDeviceToHost, e->stream)) != cudaSuccess) break;
        if ((ce = cudaMemcpyAsync(parity_ok, d_pc, 4 * n, cudaMemcpyDeviceToHost, e->stream)) != cudaSuccess) break;
        ce = cudaStreamSynchronize(e->stream);
    } while (0);
    cudaFree(d_llr); cudaFree(d_bits); cudaFree(d_it); cudaFree(d_pc);
    if (ce != cudaSuccess) return wb_fail(WB_ECUDA, "wb_ldpc_decode_batch: %s", cudaGetErrorString(ce));
    return rc;
}

extern "C" int wb_sd_to_llr_batch(wb_engine *e, const float *sd, size_t n, float *llr)
{
    if (!e || !sd || !llr) return wb_fail(WB_EINVAL, "null argument");
    if (n == 0) return WB_OK;
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    float *d_sd = nullptr, *d_llr = nullptr; double *d_c4 = nullptr;
    cudaError_t ce;
    if ((ce = cudaMalloc(&d_sd, sizeof(float) * WB_NCODE * n)) != cudaSuccess ||
        (ce = cudaMalloc(&d_llr, sizeof(float) * WB_NCODE * n)) != cudaSuccess ||
        (ce = cudaMalloc(&d_c4, sizeof(double) * n)) != cudaSuccess) {
        cudaFree(d_sd); cudaFree(d_llr); cudaFree(d_c4);
        return wb_fail(WB_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(ce));
    }
    do {
        if ((ce = cudaMemcpyAsync(d_sd, sd, sizeof(float) * WB_NCODE * n, cudaMemcpyHostToDevice, e->stream)) != cudaSuccess) break;
        wb_llr_stats_plain_kernel<<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(d_sd, d_c4, (long long)n);
        wb_llr_scale_kernel<<<(unsigned)((n * WB_NCODE + 255) / 256), 256, 0, e->stream>>>(d_sd, d_c4, d_llr, (long long)n);
        e->launches += 2;
        if ((ce = cudaGetLastError()) != cudaSuccess) break;
        if ((ce = cudaMemcpyAsync(llr, d_llr, sizeof(float) * WB_NCODE * n, cudaMemcpyDeviceToHost, e->stream)) != cudaSuccess) break;
        ce = cudaStreamSynchronize(e->stream);
    } while (0);
    cudaFree(d_sd); cudaFree(d_llr); cudaFree(d_c4);
    if (ce != cudaSuccess) return wb_fail(WB_ECUDA, "wb_sd_to_llr_batch: %s", cudaGetErrorString(ce));
    return WB_OK;
}

/* ---- HBM-resident benchmarking helpers ---------------------------------- */

extern "C" int wb_dev_input(wb_engine *e, void **dptr, uint64_t *stride_bytes, uint64_t *capacity_samples)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    if (dptr) *dptr = e->d_in + (size_t)WB_HEADROOM * e->fp.in_bps;
    if (stride_bytes) *stride_bytes = e->in_stride;
    if (capacity_samples) *capacity_samples = e->cfg.chunk_samples;
    return WB_OK;
}

__global__ void wb_rewind_kernel(wb_stream_state *st, int n, unsigned headroom, unsigned long long nsamp)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    st[s].in_pos = headroom;
    st[s].in_fill = headroom + nsamp;
}

extern "C" int wb_dev_set_fill(wb_engine *e, uint64_t nsamp)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    int rc = e->resident_mode ? WB_OK : wb_collect(e);
    if (rc) return rc;
    if (nsamp > e->cfg.chunk_samples) return wb_fail(WB_ERANGE, "nsamp exceeds chunk_samples");
    CU(cudaSetDevice(e->cfg.device));
    e->resident_mode = true;
    wb_rewind_kernel<<<(e->cfg.n_streams + 127) / 128, 128, 0, e->stream>>>(e->d_state, e->cfg.n_streams, WB_HEADROOM, nsamp);
    e->launches++;
    CU(cudaGetLastError());
    return WB_OK;
}

__global__ void wb_replicate_kernel(unsigned char *in, unsigned long long stride, int n_src, int n_streams,
                                    unsigned long long nbytes, unsigned long long rot_bytes, unsigned headroom_bytes)
{
    /* stream s >= n_src := source (s % n_src) rotated left by (s / n_src) * rot bytes; 16-byte granules */
    int s = n_src + blockIdx.y;
    if (s >= n_streams) return;
    const unsigned char *src = in + (size_t)(s % n_src) * stride + headroom_bytes;
    unsigned char *dst = in + (size_t)s * stride + headroom_bytes;
    unsigned long long r = ((unsigned long long)(s / n_src) * rot_bytes) % nbytes;
    unsigned long long ng = nbytes / 16;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long so = (g * 16 + r) % nbytes;
        uint4 v;
        if (so + 16 <= nbytes && (so & 15) == 0) v = *reinterpret_cast<const uint4 *>(src + so);
        else {
            unsigned char t[16];
            for (int b = 0; b < 16; b++) t[b] = src[(so + b) % nbytes];
            memcpy(&v, t, 16);
        }
        *reinterpret_cast<uint4 *>(dst + g * 16) = v;
    }
}

extern "C" int wb_dev_replicate(wb_engine *e, int n_src, uint64_t nsamp, uint64_t rot)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    if (n_src <= 0 || n_src > e->cfg.n_streams) return wb_fail(WB_EINVAL, "n_src out of range");
    if (nsamp > e->cfg.chunk_samples) return wb_fail(WB_ERANGE, "nsamp exceeds chunk_samples");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    const int bps = e->fp.in_bps;
    unsigned long long nbytes = nsamp * bps;
    if (nbytes % 16) return wb_fail(WB_EINVAL, "nsamp * bytes-per-sample must be a multiple of 16");
    if ((rot * bps) % 16) return wb_fail(WB_EINVAL, "rot * bytes-per-sample must be a multiple of 16");
    int rest = e->cfg.n_streams - n_src;
    if (rest > 0) {
        dim3 grid(64, rest);
        wb_replicate_kernel<<<grid, 256, 0, e->stream>>>(e->d_in, e->in_stride, n_src, e->cfg.n_streams, nbytes, rot * bps,
                                                        WB_HEADROOM * bps);
        e->launches++;
        CU(cudaGetLastError());
    }
    return WB_OK;
}

__global__ void wb_replicate_llr_kernel(float *llr, size_t n_src, size_t n)
{
    size_t total = n * WB_NCODE;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + n_src * WB_NCODE; i < total; i += (size_t)gridDim.x * blockDim.x)
        llr[i] = llr[i % (n_src * WB_NCODE)];
}

extern "C" int wb_dev_ldpc_setup(wb_engine *e, const float *llr, size_t n_src, size_t n)
{
    if (!e || !llr || n_src == 0 || n < n_src) return wb_fail(WB_EINVAL, "bad argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    cudaFree(e->d_bench_llr); cudaFree(e->d_bench_bits); cudaFree(e->d_bench_iters); cudaFree(e->d_bench_pcc);
    e->d_bench_llr = nullptr; e->d_bench_bits = nullptr; e->d_bench_iters = e->d_bench_pcc = nullptr; e->bench_n = 0;
    if (cudaMalloc(&e->d_bench_llr, sizeof(float) * WB_NCODE * n) != cudaSuccess || cudaMalloc(&e->d_bench_bits, 323 * n) != cudaSuccess ||
        cudaMalloc(&e->d_bench_iters, 4 * n) != cudaSuccess || cudaMalloc(&e->d_bench_pcc, 4 * n) != cudaSuccess)
        return wb_fail(WB_ENOMEM, "cudaMalloc LDPC bench buffers");
    CU(cudaMemcpyAsync(e->d_bench_llr, llr, sizeof(float) * WB_NCODE * n_src, cudaMemcpyHostToDevice, e->stream));
    if (n > n_src) {
        wb_replicate_llr_kernel<<<1184, 256, 0, e->stream>>>(e->d_bench_llr, n_src, n);
        e->launches++;
    }
    CU(cudaStreamSynchronize(e->stream));
    e->bench_n = n;
    return WB_OK;
}

extern "C" int wb_dev_ldpc_run(wb_engine *e, int max_iter)
{
    if (!e || !e->bench_n) return wb_fail(WB_EINVAL, "wb_dev_ldpc_setup first");
    CU(cudaSetDevice(e->cfg.device));
    wb_ldpc_args la;
    memset(&la, 0, sizeof(la));
    la.llr_in = e->d_bench_llr; la.bits_out = e->d_bench_bits; la.iters_out = e->d_bench_iters; la.pcc_out = e->d_bench_pcc;
    la.max_iter = max_iter > 0 ? max_iter : e->max_iter;
    la.vedge = e->d_vedge; la.crc_tab = e->d_crc_tab; la.crc0 = e->crc0; la.scramble = e->d_scramble; la.lut = e->d_lut;
    wb_ldpc_kernel<<<(unsigned)e->bench_n, WB_LDPC_THREADS, sizeof(wb_ldpc_smem), e->stream>>>(la, (long long)e->bench_n);
    e->launches++;
    CU(cudaGetLastError());
    return WB_OK;
}

extern "C" int wb_dev_ldpc_result(wb_engine *e, size_t first, size_t n, uint8_t *bits_packed, int32_t *iters, int32_t *parity_ok)
{
    if (!e || first + n > e->bench_n) return wb_fail(WB_EINVAL, "range");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    if (bits_packed) CU(cudaMemcpy(bits_packed, e->d_bench_bits + first * 323, 323 * n, cudaMemcpyDeviceToHost));
    if (iters) CU(cudaMemcpy(iters, e->d_bench_iters + first, 4 * n, cudaMemcpyDeviceToHost));
    if (parity_ok) CU(cudaMemcpy(parity_ok, e->d_bench_pcc + first, 4 * n, cudaMemcpyDeviceToHost));
    return WB_OK;
}

extern "C" int wb_timer_start(wb_engine *e)
{
    if (!e) return wb_fail(WB_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventRecord(e->tev0, e->stream));
    return WB_OK;
}

extern "C" int wb_timer_stop(wb_engine *e, float *ms)
{
    if (!e || !ms) return wb_fail(WB_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventRecord(e->tev1, e->stream));
    CU(cudaEventSynchronize(e->tev1));
    CU(cudaEventElapsedTime(ms, e->tev0, e->tev1));
    return WB_OK;
}

extern "C" int wb_last_kernel_ms(wb_engine *e, float *ms)
{
    if (!e || !ms) return wb_fail(WB_EINVAL, "null argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    for (int i = 0; i < 4; i++) ms[i] = e->kernel_ms[i];
    return WB_OK;
}

extern "C" uint64_t wb_launch_count(wb_engine *e) { return e ? e->launches : 0; }
extern "C" uint64_t wb_last_codewords(wb_engine *e) { if (!e || wb_collect(e)) return 0; return e->last_codewords; }
extern "C" uint64_t wb_last_samples(wb_engine *e) { if (!e || wb_collect(e)) return 0; return e->last_samples; }

/* test tap: per-frame log of the next wb_process calls (8 floats per frame: nin, bins[4], norm_rx_timing, ppm, rx_timing) */
extern "C" int wb_enable_frame_log(wb_engine *e, int frames_per_stream)
{
    if (!e || frames_per_stream < 0) return wb_fail(WB_EINVAL, "bad argument");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    cudaFree(e->d_frame_log); e->d_frame_log = nullptr; e->log_cap = 0;
    if (frames_per_stream) {
        size_t bytes = sizeof(float) * 8 * (size_t)frames_per_stream * e->cfg.n_streams;
        if (cudaMalloc(&e->d_frame_log, bytes) != cudaSuccess) return wb_fail(WB_ENOMEM, "cudaMalloc frame log");
        CU(cudaMemset(e->d_frame_log, 0, bytes));
        e->log_cap = frames_per_stream;
    }
    return WB_OK;
}

extern "C" int wb_read_frame_log(wb_engine *e, int stream, float *buf, size_t cap_frames)
{
    if (!e || !buf || !e->d_frame_log) return wb_fail(WB_EINVAL, "frame log not enabled");
    int rc = wb_collect(e);
    if (rc) return rc;
    CU(cudaSetDevice(e->cfg.device));
    size_t nfr = std::min((size_t)e->log_cap, cap_frames);
    CU(cudaMemcpy(buf, e->d_frame_log + (size_t)stream * e->log_cap * 8, sizeof(float) * 8 * nfr, cudaMemcpyDeviceToHost));
    return WB_OK;
}

/* pinned host memory for the caller's staging buffers (the e2e path copies from these) */
extern "C" void *wb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { wb_fail(WB_ENOMEM, "cudaHostAlloc %zu", bytes); return nullptr; }
    return p;
}
extern "C" void wb_host_free(void *p) { if (p) cudaFreeHost(p); }

/* ---- host <-> device copy probe ------------------------------------------
   The host-buffer path (wb_feed / wb_feed_strided from pinned memory) can go no faster than the PCIe / host-bridge leg;
   this measures that leg alone: `bytes` of pinned host memory and of device memory, flat cudaMemcpyAsync, CUDA events. */
struct wb_copy_probe { int device; size_t bytes; unsigned char *h, *d; cudaStream_t s; cudaEvent_t e0, e1; };

extern "C" void wb_copy_probe_destroy(wb_copy_probe *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->s) cudaStreamSynchronize(p->s);
    if (p->h) cudaFreeHost(p->h);
    if (p->d) cudaFree(p->d);
    if (p->e0) cudaEventDestroy(p->e0);
    if (p->e1) cudaEventDestroy(p->e1);
    if (p->s) cudaStreamDestroy(p->s);
    delete p;
}

extern "C" int wb_copy_probe_create(int device, size_t bytes, wb_copy_probe **out)
{
    if (!out || bytes == 0) return wb_fail(WB_EINVAL, "bad argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return wb_fail(WB_ENODEV, "no CUDA device visible");
    if (device < 0 || device >= ndev) return wb_fail(WB_EINVAL, "device %d of %d", device, ndev);
    CU(cudaSetDevice(device));
    wb_copy_probe *p = new wb_copy_probe();
    memset(p, 0, sizeof(*p));
    p->device = device; p->bytes = bytes;
    if (cudaHostAlloc(&p->h, bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&p->d, bytes) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->s, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&p->e0) != cudaSuccess ||
        cudaEventCreate(&p->e1) != cudaSuccess) {
        wb_copy_probe_destroy(p);
        return wb_fail(WB_ENOMEM, "copy probe: allocation of %zu bytes failed", bytes);
    }
    memset(p->h, 1, bytes);                                        /* touch every page */
    if (cudaMemcpyAsync(p->d, p->h, bytes, cudaMemcpyHostToDevice, p->s) != cudaSuccess || cudaStreamSynchronize(p->s) != cudaSuccess) {
        wb_copy_probe_destroy(p);
        return wb_fail(WB_ECUDA, "copy probe: warm-up copy failed");
    }
    *out = p;
    return WB_OK;
}

extern "C" int wb_copy_probe_run(wb_copy_probe *p, int iters, int d2h, float *ms)
{
    if (!p || !ms || iters <= 0) return wb_fail(WB_EINVAL, "bad argument");
    CU(cudaSetDevice(p->device));
    CU(cudaEventRecord(p->e0, p->s));
    for (int i = 0; i < iters; i++) {
        if (d2h) CU(cudaMemcpyAsync(p->h, p->d, p->bytes, cudaMemcpyDeviceToHost, p->s));
        else CU(cudaMemcpyAsync(p->d, p->h, p->bytes, cudaMemcpyHostToDevice, p->s));
    }
    CU(cudaEventRecord(p->e1, p->s));
    CU(cudaEventSynchronize(p->e1));
    CU(cudaEventElapsedTime(ms, p->e0, p->e1));
    return WB_OK;
}

/* geometry the host wrapper needs */
extern "C" int wb_geometry(wb_engine *e, int32_t *out, int n)
{
    if (!e || !out) return wb_fail(WB_EINVAL, "null argument");
    int32_t g[14] = {e->fp.N, e->fp.Nbits, e->fp.Ts, e->fp.P, e->fp.Ndft, e->fp.nmax, e->job_cap, (int32_t)e->sd_cap,
                     e->fp.Nsym, e->fp.M, e->max_iter, e->dp.nsym, e->spb, (int32_t)e->fsk_smem};
    for (int i = 0; i < n && i < 14; i++) out[i] = g[i];
    return WB_OK;
}

/* ---- transmit side on the device ----------------------- */

extern "C" int wb_tx_synthesize(wb_engine *e, const uint8_t *payloads, const wb_tx_config *cfg, uint64_t *nsamp_per_stream)
{
    if (!e || !payloads || !cfg) return wb_fail(WB_EINVAL, "null argument");
    if (cfg->struct_size != sizeof(wb_tx_config)) return wb_fail(WB_EINVAL, "wb_tx_config.struct_size mismatch");
    if (cfg->n_packets <= 0 || cfg->lead_in_bits < 0 || cfg->gap_bits < 0 || cfg->tail_bits < 0) return wb_fail(WB_EINVAL, "bad wb_tx_config");
    int rc = wb_collect(e);
    if (rc) return rc;
    if (e->resident_mode) return wb_fail(WB_EINVAL, "engine is in resident (wb_dev_set_fill) mode");
    CU(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams, M = e->fp.M, Ts = e->fp.Ts, Nsym = e->fp.Nsym, framing = e->cfg.framing;
    const int bps = (M == 2) ? 1 : 2;
    const int frame_bits = framing == WB_FRAMING_V1 ? WB_TX_RAW_BYTES * 10 : (framing == WB_FRAMING_V2 ? WB_TX_RAW_BYTES * 8 : WB_PACKET_BYTES * 8);
    unsigned long long nbits = (unsigned long long)cfg->lead_in_bits + (unsigned long long)cfg->n_packets * (frame_bits + cfg->gap_bits) + cfg->tail_bits;
    const unsigned long long per_call = (unsigned long long)Nsym * bps;
    const unsigned long long n_calls = (nbits + per_call - 1) / per_call;
    nbits = n_calls * per_call;
    const unsigned long long nsamp = n_calls * Nsym * Ts;
    for (int s = 0; s < n; s++)
        if (WB_HEADROOM + e->fill[s] + nsamp > e->in_cap) return wb_fail(WB_ERANGE, "stream %d: chunk capacity exceeded (%llu samples)", s, nsamp);

    const bool noise = cfg->ebno_db_per_stream != nullptr || !std::isnan(cfg->ebno_db);
    const bool cf32 = e->fp.in_fmt == WB_FMT_CF32;
    uint8_t *d_pl = nullptr;
    float *d_work = nullptr, *d_peak = nullptr, *d_sigma = nullptr;
    unsigned long long *d_off = nullptr;
    auto cleanup = [&]() { cudaFree(d_pl); cudaFree(d_work); cudaFree(d_peak); cudaFree(d_off); cudaFree(d_sigma); };
#define CT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
        return wb_fail(e_ == cudaErrorMemoryAllocation ? WB_ENOMEM : WB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    const size_t pl_bytes = (size_t)n * cfg->n_packets * WB_PACKET_BYTES;
    CT(cudaMalloc(&d_pl, pl_bytes));
    CT(cudaMemcpyAsync(d_pl, payloads, pl_bytes, cudaMemcpyHostToDevice, e->stream));
    /* the bit rows stay with the engine until the next synthesis (test tap) */
    cudaFree(e->d_tx_bits); e->d_tx_bits = nullptr;
    e->tx_bits_stride = (nbits + 63) & ~63ULL; e->tx_nbits = nbits;
    CT(cudaMalloc(&e->d_tx_bits, e->tx_bits_stride * n));
    CT(cudaMemsetAsync(e->d_tx_bits, 1, e->tx_bits_stride * n, e->stream));          /* idle = '1' */
    if (!e->d_hrows) {
        CT(cudaMalloc(&e->d_hrows, sizeof(wb_hrows)));
        CT(cudaMemcpyAsync(e->d_hrows, wb_hrows, sizeof(wb_hrows), cudaMemcpyHostToDevice, e->stream));
    }
    std::vector<unsigned long long> off(n);
    for (int s = 0; s < n; s++) off[s] = WB_HEADROOM + (e->fill[s] - e->cursor[s].in_fill);
    CT(cudaMalloc(&d_off, sizeof(unsigned long long) * n));
    CT(cudaMemcpyAsync(d_off, off.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, e->stream));
    CT(cudaMalloc(&d_peak, sizeof(float) * n));
    CT(cudaMemsetAsync(d_peak, 0, sizeof(float) * n, e->stream));
    if (!cf32) CT(cudaMalloc(&d_work, sizeof(float2) * nsamp * n));

    wb_tx_args ta;
    memset(&ta, 0, sizeof(ta));
    ta.payloads = d_pl; ta.bits = e->d_tx_bits; ta.bits_stride = e->tx_bits_stride; ta.n_streams = n; ta.n_packets = cfg->n_packets;
    ta.framing = framing; ta.lead_in = cfg->lead_in_bits; ta.gap = cfg->gap_bits; ta.frame_bits = frame_bits;
    ta.hrows = e->d_hrows; ta.scramble = e->d_scramble;
    const long long npk = (long long)n * cfg->n_packets;
    if (framing == WB_FRAMING_NONE) {
        wb_tx_raw_bits_kernel<<<(unsigned)((npk * WB_PACKET_BYTES * 8 + 255) / 256), 256, 0, e->stream>>>(ta);
    } else {
        wb_tx_frame_kernel<<<(unsigned)((npk + 3) / 4), 128, 0, e->stream>>>(ta);
    }
    wb_mod_args ma;
    memset(&ma, 0, sizeof(ma));
    ma.bits = e->d_tx_bits; ma.bits_stride = e->tx_bits_stride; ma.peak = d_peak; ma.n_streams = n; ma.M = M; ma.Ts = Ts; ma.Nsym = Nsym;
    ma.n_calls = n_calls; ma.seed = cfg->seed; ma.unit = noise ? 1 : 0;
    if (cf32) { ma.work = reinterpret_cast<float *>(e->d_in); ma.work_stride = e->in_stride / sizeof(float2); ma.work_off = d_off; }
    else { ma.work = d_work; ma.work_stride = nsamp; ma.work_off = nullptr; }
    for (int m = 0; m < M; m++) {                  /* dosc_f[m], src/fsk.c:1174-1176 */
        hcpx d = hcexpj((float)(2 * M_PI * ((float)(cfg->f1_tx + (cfg->fs_tx * m)) / (float)(e->fp.Fs))));
        ma.dosc[m] = make_float2(d.r, d.i);
    }
    std::vector<float> sig;
    if (noise) {                                   /* generate_lowsnr.py:70-77 with unit signal variance */
        auto sigma_of = [&](double db) { return (float)sqrt((double)e->fp.Fs / ((double)e->fp.Rs * pow(10.0, db / 10.0) * bps) / 2.0); };
        if (cfg->ebno_db_per_stream) {
            sig.resize(n);
            for (int s = 0; s < n; s++) {
                if (std::isnan(cfg->ebno_db_per_stream[s])) { cleanup(); return wb_fail(WB_EINVAL, "ebno_db_per_stream[%d] is NaN", s); }
                sig[s] = sigma_of(cfg->ebno_db_per_stream[s]);
            }
            CT(cudaMalloc(&d_sigma, sizeof(float) * n));
            CT(cudaMemcpyAsync(d_sigma, sig.data(), sizeof(float) * n, cudaMemcpyHostToDevice, e->stream));
            ma.sigma_per_stream = d_sigma;
        } else {
            ma.sigma = sigma_of(cfg->ebno_db);
        }
    }
    wb_tx_mod_kernel<<<(n + 31) / 32, 32, 0, e->stream>>>(ma);
    if (noise || !cf32) {
        dim3 grid(std::max(1u, std::min(64u, (unsigned)((nsamp + 255) / 256))), (unsigned)n);
        wb_tx_scale_kernel<<<grid, 256, 0, e->stream>>>(ma.work, ma.work_stride, ma.work_off, d_peak, noise ? 1 : 0, e->d_in, e->in_stride,
                                                        d_off, e->fp.in_fmt, nsamp, n);
    }
    e->launches += 3;
    CT(cudaStreamSynchronize(e->stream));
    CT(cudaGetLastError());
#undef CT
    cleanup();
    for (int s = 0; s < n; s++) e->fill[s] += nsamp;
    for (int s = 1; s < n && e->uniform_fill; s++)
        if (e->fill[s] - e->cursor[s].in_fill != e->fill[0] - e->cursor[0].in_fill) e->uniform_fill = false;
    if (nsamp_per_stream) *nsamp_per_stream = nsamp;
    return WB_OK;
}

extern "C" int wb_dev_read_input(wb_engine *e, int stream, uint64_t first, uint64_t nsamp, void *out)
{
    if (!e || !out) return wb_fail(WB_EINVAL, "null argument");
    if (stream < 0 || stream >= e->cfg.n_streams) return wb_fail(WB_EINVAL, "stream %d out of range", stream);
    if (WB_HEADROOM + first + nsamp > e->in_cap) return wb_fail(WB_ERANGE, "beyond the input row");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(out, e->d_in + (size_t)stream * e->in_stride + (WB_HEADROOM + first) * e->fp.in_bps, nsamp * e->fp.in_bps,
                  cudaMemcpyDeviceToHost));
    return WB_OK;
}

extern "C" int wb_tx_read_bits(wb_engine *e, int stream, uint8_t *bits, size_t cap, size_t *nout)
{
    if (!e || !bits || !nout) return wb_fail(WB_EINVAL, "null argument");
    if (stream < 0 || stream >= e->cfg.n_streams) return wb_fail(WB_EINVAL, "stream %d out of range", stream);
    if (!e->d_tx_bits) return wb_fail(WB_EINVAL, "no wb_tx_synthesize yet");
    if (cap < e->tx_nbits) return wb_fail(WB_ERANGE, "need room for %llu bits", e->tx_nbits);
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaMemcpy(bits, e->d_tx_bits + (size_t)stream * e->tx_bits_stride, e->tx_nbits, cudaMemcpyDeviceToHost));
    *nout = (size_t)e->tx_nbits;
    return WB_OK;
}

#ifdef WB_PHASE_CLK
/* debug build only (make dbg -> libwenet_b200_dbg.so): read and clear the twelve cycle counters of wb_fsk_kernel
   (0-3, 6: phases; 4, 5, 7: the last oscillator warp's way to its segment; 8-11: when each oscillator warp is done) */
extern "C" int wb_debug_phase_clk(unsigned long long *out12)
{
    unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out12, wb_phase_clk, sizeof(z)) != cudaSuccess) return WB_ECUDA;
    if (cudaMemcpyToSymbol(wb_phase_clk, z, sizeof(z)) != cudaSuccess) return WB_ECUDA;
    return WB_OK;
}
#endif
