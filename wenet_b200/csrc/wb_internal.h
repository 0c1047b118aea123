/*
 * wb_internal.h -- device-visible parameter blocks and per-stream state.
 *
 * HBM layout (one engine = one GPU):
 *   d_in      [n_streams][in_stride]  raw input samples in cfg.in_fmt (cf32: 8 B, cs16: 4 B, cu8/s16: 2 B)
 *   d_state   [n_streams]             wb_stream_state (persistent modem + deframer state)
 *   d_sd      [n_streams][sd_stride]  float soft decisions: [carry of a half-collected packet | this chunk]
 *                                     the chunk starts at float index WB_CARRY_CAP of the row
 *   d_hard    [n_streams][WB_HARD_PRE + sd_cap]  uint8 hard bits, only with WB_FLAG_HARD_BITS: [.. last frame of the
 *                                     previous chunk | this chunk], the chunk starts at byte WB_HARD_PRE of the row
 *   d_cursor  [n_streams]             wb_cursor (what the host reads back after every wb_process)
 *   d_jobs    [n_streams][job_cap]    unsigned: row offset of the first collected symbol of each codeword
 *   d_c4      [n_streams][job_cap]    double: 4 * estEsN0 of each codeword (sd_to_llr statistics)
 *   d_cw      [n_streams][job_cap]    wb_codeword records (LDPC + CRC output)
 *   d_llr     [n_streams][job_cap][2580] float LLRs, only with WB_FLAG_KEEP_LLR
 */
#ifndef WB_INTERNAL_H
#define WB_INTERNAL_H

#include <stdint.h>
#include "wenet_b200.h"
#include "wb_tables.h"

#define WB_MAXM 4
#define WB_MAX_NDFT 256                /* estimator FFT size the kernels are laid out for */
#define WB_MAX_TS 10                   /* samples per symbol (v1: 8, v2: 10) */
#define WB_MAX_NSTASH (4 * WB_MAX_TS)
#define WB_MAX_NIN 512                 /* N + Ts/2 must not exceed this */
#define WB_MAX_LEVELS 8
#define WB_MAX_B1W 14
#define WB_EYE_KEEP 96                 /* integrator outputs kept for the eye diagram: 8/M traces x 2P + offset */
#define WB_MAX_NINT 496                /* (Nsym + 1) * P <= 49 * 10, rounded */
#define WB_PFT_NT 2                    /* blocks of the fine-timing oscillator table before it is exactly periodic (checked by the host) */
#define WB_FRAME_SYMS 48               /* nsyms, reference src/fsk.c:134 */

#define WB_PKT_BODY_BYTES 323          /* 256 payload + 2 crc + 65 parity */
#define WB_HARD_PRE 128                /* >= 2 * WB_FRAME_SYMS: room for the previous chunk's last frame of hard bits */
#define WB_CARRY_CAP 3264              /* >= 3230 symbols of a half-collected v1 packet, 64-float aligned */

#define WB_LDPC_THREADS 288            /* 9 warps: 516 checks = 2 rounds, 2580 variables = 9 rounds */
#define WB_LDPC_SLOTS 14               /* edge slots per check: 12 H1 + parity j-1 + parity j */
#define WB_LDPC_NMSG (WB_LDPC_SLOTS * WB_NPAR)   /* 7224 message words (slot 12 of check 0 is unused) */

/* shared-memory geometry of wb_fsk_kernel per stream, one formula for host (build_fsk_params) and kernel:
   xlen = float2 of X (old + new samples), ylen = float2 of one other-tone buffer, blen = float2 of all of them
   (>= Ndft: the FFT work buffer), efl = floats of E, sreg = bytes per stream region (== 8 mod 16, so the
   lanes of a sequential-phase warp, one stream each, hit distinct banks).  wb_blk_* : P == Ts. */
#ifdef __CUDACC__
#define WB_HD __host__ __device__
#else
#define WB_HD
#endif
WB_HD constexpr int wb_geom_xlen(int Ts) { return ((2 * Ts + Ts / 2) + (Ts * WB_FRAME_SYMS + Ts / 2) + 1) & ~1; }
WB_HD constexpr int wb_geom_ylen(int Ts, int step) { return ((Ts * WB_FRAME_SYMS + 2 * Ts - step) + 1) & ~1; }
WB_HD constexpr int wb_geom_blen(int M, int ylen) { return (M - 1) * ylen > WB_MAX_NDFT ? (M - 1) * ylen : WB_MAX_NDFT; }
WB_HD constexpr int wb_geom_efl(int P) { return (WB_FRAME_SYMS + 1) * P; }          /* one e_i per integrator output */
WB_HD constexpr int wb_geom_sreg(int xlen, int blen, int efl)
{
    const int bytes = (xlen + blen) * 8 + ((efl * 4 + 7) & ~7);
    const int sreg = (bytes & ~15) + 8;
    return sreg < bytes ? sreg + 16 : sreg;
}
/* bytes in front of the stream regions: the frame scalars (wb_fsk_sc, `sc_bytes` each), one 8-byte mbarrier per stream
   (the TMA frame fetch), rounded to 128, then the FFT twiddles */
WB_HD constexpr int wb_geom_head(int spb, int sc_bytes) { return ((sc_bytes + 8) * spb + 127) / 128 * 128 + 3 * (WB_MAX_NDFT / 4) * 8; }
WB_HD constexpr int wb_blk_ylen(int Ts) { return wb_geom_ylen(Ts, 1); }
WB_HD constexpr int wb_blk_blen(int M, int Ts) { return wb_geom_blen(M, wb_blk_ylen(Ts)); }
WB_HD constexpr int wb_blk_sreg(int M, int Ts) { return wb_geom_sreg(wb_geom_xlen(Ts), wb_blk_blen(M, Ts), wb_geom_efl(Ts)); }

struct wb_fsk_params {
    int Fs, Rs, Ts, P, M, N, Nsym, Nmem, Nbits, Ndft, nstash;
    int nst;                  /* 2*Ts + Ts/2: old samples the mixer can reach back to (<= nstash = 4*Ts) */
    int step;                 /* Ts / P */
    int nint;                 /* (Nsym + 1) * P integrator outputs per frame */
    int nsteps;               /* Nmem - step mixer steps per frame */
    int nmax;                 /* N + Ts/2: longest frame */
    int f_min, f_max, f_zero; /* estimator bin limits, reference src/fsk.c:568-570 */
    float tc;                 /* 0.95 * Ndft / Fs, reference src/fsk.c:573 */
    int in_fmt, in_bps;       /* bytes per input sample */
    int stats;                /* WB_FLAG_STATS: per-frame Eb/N0 terms and eye-diagram tap */
    int b1_w;                 /* warps sharing the sequential mixer phase, and their frame segments */
    int b1_seg[WB_MAX_B1W + 1];
    int xlen, ylen, blen, sreg; /* smem geometry per stream: float2 of X, of one other-tone buffer, of all of them
                                  (>= Ndft: FFT work buffer); bytes per stream region (== 8 mod 16) */
    /* host-built constant tables (glibc cosf/sinf on the host = what the reference would use) */
    const float  *hann;       /* [Ndft]       reference src/fsk.c:94-111 */
    const float2 *tw;         /* [Ndft]       kiss_fft twiddles, reference src/kiss_fft.c:357-363 */
    const uint16_t *perm;     /* [Ndft]       leaf load order of the DIT recursion */
    int pft_steady;           /* P == Ts and pft[i] == pft[i - P] for every i >= (WB_PFT_NT + 1) * P: see wb_b3_chain */
    const float2 *dphi;       /* [Ndft/2]     comp_exp_j(2 pi f/Fs), f = bin*Fs/Ndft, src/fsk.c:763 */
    const float2 *back;       /* [3][Ndft/2]  phase back-off for nin = N-Ts/2, N, N+Ts/2, src/fsk.c:758 */
    /* the fine-timing oscillator again, by value: kernel parameters live in the constant bank, which is the
       right home for a table every lane reads at the same index */
    float4 pftc4[2][WB_MAX_NINT / 4];   /* [0] = real parts, [1] = imaginary parts, element i = ((float *)pftc4[c])[i] */
};

struct wb_stream_state {
    /* demodulator, reference struct FSK src/fsk.h:43-90 */
    float2 phi_c[WB_MAXM];
    int    fbin[WB_MAXM];     /* f_est as estimator bins (f_est = bin * Fs/Ndft) */
    float  norm_rx_timing, ppm;
    int    nin;
    float  rx_timing;
    unsigned long long frames;
    float  fft_est[WB_MAX_NDFT / 2];
    float2 samp_old[WB_MAX_NSTASH];
    float  sd_last[2 * WB_FRAME_SYMS];   /* last frame's soft decisions: re-emitted when the NaN guard trips, src/fsk.c:878 */
    /* statistics taps (WB_FLAG_STATS): reference src/fsk.c:995-1080.  eb_arg = (1e-6+mean)/(1e-6+std) of the last 32
       frames (the host finishes EbNodB = -6 + 20 log10f(.) and the 0.5/0.5 snr_est IIR with its own libm);
       eye_fint = the first integrator outputs of the last frame, enough for the eye traces */
    float  eb_arg[32];
    unsigned int eb_count;
    int    eye_high;
    float2 eye_fint[WB_MAXM][WB_EYE_KEEP];
    /* deframer, reference locals of main() src/drs232_ldpc.c:106-118 */
    unsigned long long window;  /* bit_buffer, newest bit = bit 0 */
    int    collecting, ind;
    unsigned int seq;           /* codewords found so far (= packets) */
    unsigned int packets, packet_errors;
    unsigned int pad0;
    /* buffer cursors */
    unsigned long long in_pos, in_fill;   /* samples */
};

/* read back by the host after each wb_process */
struct wb_cursor {
    unsigned long long in_fill;   /* samples left resident (after compaction) */
    unsigned long long consumed;  /* samples demodulated by this wb_process */
    unsigned int n_sd;            /* soft decisions produced by this wb_process */
    unsigned int n_jobs;          /* codewords found by this wb_process */
    unsigned int seq0;            /* sequence number of the first of them */
    int nin;                      /* fsk_nin() for the next frame */
};

struct wb_deframe_params {
    int mode;                 /* WB_FRAMING_V1 / V2 */
    int uw_bits, uw_thresh;
    int nsym;                 /* symbols collected per packet: 3230 (v1) / 2584 (v2) */
    unsigned long long uw, uw_mask;
};

#endif
