/*
 * wb_fsk_kernel.cuh -- K1: batched 2/4-FSK demodulator, every frame of every resident stream.
 *
 * Replaces, per stream and per modem frame:
 *   reference src/fsk_demod.c:273-296   sample-format conversion
 *   reference src/fsk.c:540-677         fsk_demod_freq_est (windowed, zero-padded kiss_fft, IIR'd
 *                                       magnitude spectrum, M peaks with blanking)
 *   reference src/fsk.c:755-848         per-tone down-mix + sliding integrate
 *   reference src/fsk.c:853-907         fine timing, ppm, next nin
 *   reference src/fsk.c:912-993         resample, soft decisions
 * with the reference's float operation order (no FMA, IEEE div/sqrt, its atan2f), so soft decisions,
 * the nin sequence and the estimator state are bit-identical to the CPU pipe.
 *
 * Mapping.  Frames of one stream are strictly sequential (nin of frame k+1 comes out of frame k),
 * and inside a frame two recurrences are sequential as well: the tone oscillators
 * (phi_c *= dphi, Nmem-1 dependent complex products per tone) and the fine-timing accumulator
 * (nint dependent additions).  Everything else is parallel over the sample index.  So:
 *
 *   CTA = spb streams (runtime; 14 for 2-FSK at Ts = 8 so that two CTAs = 28 streams fit one SM and
 *   4096 streams are all resident on 148 SMs), one "stream warp" per stream, frames in lock step.
 *
 *   A  (stream warps)  land the prefetched frame in shared memory, window + 256-point FFT in the
 *                      reference's butterfly order, spectrum IIR (registers), warp-argmax peak
 *                      picking, then issue the global loads of the NEXT frame into registers
 *                      (consumed a whole frame later: HBM latency is off the critical path).
 *   B1 (warp 0, lane = (tone, stream))  ONLY the sequential part of the mixer: oscillator
 *                      recurrence + down-mix product, written in place over the samples.
 *   B2 (stream warps, lanes = integrator outputs)  Ts-tap sums in the reference's ring-buffer slot
 *                      order, in place; |.|^2 summed over tones -> e[i].
 *   B3 (warp 0, lane = (re/im, stream))  the sequential fine-timing accumulation over e[i], then
 *                      atan2 / ppm / nin / resampling offsets on the re-lanes.
 *   C  (stream warps, lanes = symbols)  linear-interpolated resampling and soft decisions,
 *                      48 (96) floats per frame written coalesced.
 *
 * The sequential phases cost one warp's issue slots for ALL streams of the CTA (every lane carries
 * a different dependent chain); while one CTA of an SM is in B1/B3 the other one runs A/B2/C.
 *
 * Shared memory per stream: X[nstash + nmax] float2 (old + new samples -> tone 0 mixer products ->
 * tone 0 integrator outputs, all in place), Y[(M-1) * ylen] float2 for the other tones (the FFT work
 * buffer in phase A) and E[nint] floats.  Stream regions are an odd multiple of 8 bytes mod 128
 * apart so the lanes of warp 0 (one stream each) hit distinct banks.
 * HBM traffic: every input sample is read once (8 B as cf32), 4 B x Nbits/N written.
 */
#ifndef WB_FSK_KERNEL_CUH
#define WB_FSK_KERNEL_CUH

#include "wb_internal.h"
#include "wb_math.h"

#define WB_NST ((WB_MAX_NSTASH + 31) / 32)
#define WB_NEQ (WB_MAX_NDFT / 2 / 32)

struct wb_fsk_sc {                 /* per-stream frame scalars in shared memory (88 bytes) */
    float2 phi_c[WB_MAXM];
    short pb[WB_MAXM];             /* estimator bins in force before this frame (fsk->f_est) */
    short nb[WB_MAXM];             /* bins estimated from this frame */
    int nin, flags, nin_next;      /* flags: bit 0 = active this frame, bit 1 = NaN guard tripped */
    int low, high;
    float fract, norm, ppm, rx_timing;
};

struct wb_fsk_args {
    wb_stream_state *state;
    wb_cursor *cursor;
    const unsigned char *in;       /* [n_streams][in_stride bytes] */
    unsigned long long in_stride;
    float *sd;                     /* [n_streams][sd_stride] */
    unsigned long long sd_stride;
    unsigned sd_cap;               /* floats available after WB_CARRY_CAP */
    int n_streams;
    int spb;                       /* streams per CTA = warps per CTA */
    int compact;                   /* park the unconsumed remainder right before the headroom mark */
    unsigned headroom;             /* samples; wb_feed appends at this row offset after a compacting process */
    float *frame_log;              /* optional test tap [n_streams][log_cap][8] */
    int log_cap;
};

__device__ __forceinline__ float2 wb_cmul2(float2 a, float2 b)   /* reference src/comp_prim.h cmult */
{
    float2 c;
    c.x = __fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    c.y = __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
    return c;
}

/* raw sample -> COMP, reference src/fsk_demod.c:273-296 */
template <bool CF32>
__device__ __forceinline__ float2 wb_convert(int fmt, unsigned lo, unsigned hi)
{
    float2 v;
    if (CF32) {
        v.x = __uint_as_float(lo); v.y = __uint_as_float(hi);
    } else if (fmt == WB_FMT_CU8) {
        /* ((float)u8 - 127.0) / 128.0 in double, then to float: exact, so float arithmetic gives the same */
        v.x = __fmul_rn(__fsub_rn((float)(lo & 0xffu), 127.0f), 0.0078125f);
        v.y = __fmul_rn(__fsub_rn((float)((lo >> 8) & 0xffu), 127.0f), 0.0078125f);
    } else if (fmt == WB_FMT_CS16) {
        v.x = __fdiv_rn((float)(short)(lo & 0xffffu), 1000.0f);
        v.y = __fdiv_rn((float)(short)(lo >> 16), 1000.0f);
    } else {
        v.x = __fdiv_rn((float)(short)(lo & 0xffffu), 1000.0f);
        v.y = 0.0f;
    }
    return v;
}

template <bool CF32>
__device__ __forceinline__ void wb_load_raw(int fmt, const unsigned char *p, unsigned long long idx, unsigned &lo, unsigned &hi)
{
    if (CF32) {
        uint2 r = __ldg(reinterpret_cast<const uint2 *>(p) + idx);
        lo = r.x; hi = r.y;
    } else if (fmt == WB_FMT_CS16) {
        lo = __ldg(reinterpret_cast<const unsigned *>(p) + idx); hi = 0u;
    } else {
        lo = __ldg(reinterpret_cast<const unsigned short *>(p) + idx); hi = 0u;
    }
}

/* one mixer step: product with the conjugated oscillator (reference src/fsk.c:794-798) */
#define WB_MIX_STEP(SRC, DST, N)                                                            \
    do {                                                                                    \
        const float2 x_ = (SRC)[N];                                                         \
        float2 o_;                                                                          \
        o_.x = __fadd_rn(__fmul_rn(x_.x, ph.x), __fmul_rn(x_.y, ph.y));                     \
        o_.y = __fsub_rn(__fmul_rn(x_.y, ph.x), __fmul_rn(x_.x, ph.y));                     \
        (DST)[N] = o_;                                                                      \
        ph = wb_cmul2(ph, d);                                                               \
    } while (0)

template <int M, int TS, bool CF32>
__global__ void __launch_bounds__(M == 2 ? 448 : 256, 2)
wb_fsk_kernel(wb_fsk_params p, wb_fsk_args a)
{
    constexpr int NPRE = (TS * WB_FRAME_SYMS + TS / 2 + 31) / 32;     /* lanes x NPRE >= nmax */
    extern __shared__ __align__(16) unsigned char wb_fsk_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int spb = a.spb;
    const int sg = blockIdx.x * spb + warp;
    const bool have = sg < a.n_streams;

    wb_fsk_sc *sc = reinterpret_cast<wb_fsk_sc *>(wb_fsk_raw);
    unsigned char *regions = wb_fsk_raw + ((sizeof(wb_fsk_sc) * spb + 127) / 128) * 128;
    float2 *X = reinterpret_cast<float2 *>(regions + (size_t)warp * p.sreg);
    float2 *Y = X + p.xlen;
    float *E = reinterpret_cast<float *>(Y + p.blen);
    const int Ndft = p.Ndft, nh = Ndft >> 1, nstash = p.nstash, fmt = p.in_fmt;

    /* ---- per-stream state -> registers / shared memory ---- */
    wb_stream_state *st = have ? a.state + sg : nullptr;
    const unsigned char *in = a.in + (size_t)(have ? sg : 0) * a.in_stride;
    float *sdrow = a.sd + (size_t)(have ? sg : 0) * a.sd_stride + WB_CARRY_CAP;
    unsigned long long pos = 0, fill = 0, pos0 = 0;
    unsigned frames = 0;
    int nin = p.N;
    unsigned n_out = 0;
    float est[WB_NEQ];
    float2 stash[WB_NST];
    unsigned pre_lo[NPRE], pre_hi[NPRE];
#pragma unroll
    for (int q = 0; q < WB_NEQ; q++) est[q] = 0.0f;
#pragma unroll
    for (int q = 0; q < WB_NST; q++) stash[q] = make_float2(0.0f, 0.0f);
    if (have) {
        pos = pos0 = st->in_pos; fill = st->in_fill; nin = st->nin;
#pragma unroll
        for (int q = 0; q < WB_NEQ; q++)
            if (lane + 32 * q < nh) est[q] = st->fft_est[lane + 32 * q];
#pragma unroll
        for (int q = 0; q < WB_NST; q++)
            if (lane + 32 * q < nstash) X[lane + 32 * q] = st->samp_old[lane + 32 * q];
        if (lane == 0) {
            wb_fsk_sc &c = sc[warp];
            for (int m = 0; m < M; m++) { c.phi_c[m] = st->phi_c[m]; c.pb[m] = (short)st->fbin[m]; c.nb[m] = 0; }
            c.norm = st->norm_rx_timing; c.ppm = st->ppm; c.rx_timing = st->rx_timing;
            c.nin = nin; c.nin_next = nin; c.flags = 0; c.low = c.high = 0; c.fract = 0.0f;
        }
    } else if (lane == 0) {
        wb_fsk_sc &c = sc[warp];
        for (int m = 0; m < M; m++) { c.phi_c[m] = make_float2(1.0f, 0.0f); c.pb[m] = 0; c.nb[m] = 0; }
        c.nin = p.N; c.flags = 0; c.nin_next = p.N; c.norm = 0.0f; c.ppm = 0.0f; c.rx_timing = 0.0f;
        c.low = c.high = 0; c.fract = 0.0f;
    }
    /* prefetch the first frame */
#pragma unroll
    for (int q = 0; q < NPRE; q++) {
        const int n = lane + 32 * q;
        pre_lo[q] = pre_hi[q] = 0u;
        if (have && n < p.nmax && pos + n < fill) wb_load_raw<CF32>(fmt, in, pos + n, pre_lo[q], pre_hi[q]);
    }

    const float omt = __fsub_rn(1.0f, p.tc);

    for (;;) {
        const bool active = have && (pos + (unsigned long long)nin <= fill) && (n_out + (unsigned)p.Nbits <= a.sd_cap);
        if (!__syncthreads_or(active)) break;
        const unsigned long long pos_next = pos + nin;
        const int xo = nstash - (p.Nmem - nin);      /* X index of the first mixer sample */

        /* ================= A: stream warps ================= */
        if (active) {
#pragma unroll
            for (int q = 0; q < NPRE; q++) {
                const int n = lane + 32 * q;
                if (n < p.nmax) X[nstash + n] = wb_convert<CF32>(fmt, pre_lo[q], pre_hi[q]);
            }
            __syncwarp();
            /* window the first nin - Ndft samples, zero-pad, in the leaf order of the DIT recursion
               (reference src/fsk.c:583-603, src/kiss_fft.c:238-306) */
            const int nwin = min(nin - Ndft, Ndft);
            float2 *F = Y;
            for (int o = lane; o < Ndft; o += 32) {
                const int idx = __ldg(&p.perm[o]);
                float2 v = make_float2(0.0f, 0.0f);
                if (idx < nwin) {
                    const float h = __ldg(&p.hann[idx]);
                    const float2 x = X[nstash + idx];
                    v.x = __fmul_rn(h, x.x); v.y = __fmul_rn(h, x.y);
                }
                F[o] = v;
            }
            __syncwarp();
            for (int L = 0; L < p.n_levels; L++) {
                const int pp = p.lev_p[L], mm = p.lev_m[L], fs = p.lev_fstride[L];
                const int nbf = Ndft / pp;
                for (int t = lane; t < nbf; t += 32) {
                    const int blk = t / mm, k = t - blk * mm;
                    const int base = blk * pp * mm + k;
                    if (pp == 4) {      /* reference src/kiss_fft.c:44-90, forward */
                        const float2 f0 = F[base], f1 = F[base + mm], f2 = F[base + 2 * mm], f3 = F[base + 3 * mm];
                        const float2 s0 = wb_cmul2(f1, __ldg(&p.tw[k * fs]));
                        const float2 s1 = wb_cmul2(f2, __ldg(&p.tw[2 * k * fs]));
                        const float2 s2 = wb_cmul2(f3, __ldg(&p.tw[3 * k * fs]));
                        const float2 s5 = make_float2(__fsub_rn(f0.x, s1.x), __fsub_rn(f0.y, s1.y));
                        const float2 aa = make_float2(__fadd_rn(f0.x, s1.x), __fadd_rn(f0.y, s1.y));
                        const float2 s3 = make_float2(__fadd_rn(s0.x, s2.x), __fadd_rn(s0.y, s2.y));
                        const float2 s4 = make_float2(__fsub_rn(s0.x, s2.x), __fsub_rn(s0.y, s2.y));
                        F[base + 2 * mm] = make_float2(__fsub_rn(aa.x, s3.x), __fsub_rn(aa.y, s3.y));
                        F[base] = make_float2(__fadd_rn(aa.x, s3.x), __fadd_rn(aa.y, s3.y));
                        F[base + mm] = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
                        F[base + 3 * mm] = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
                    } else {            /* reference src/kiss_fft.c:22-42 */
                        const float2 f0 = F[base], f1 = F[base + mm];
                        const float2 tt = wb_cmul2(f1, __ldg(&p.tw[k * fs]));
                        F[base + mm] = make_float2(__fsub_rn(f0.x, tt.x), __fsub_rn(f0.y, tt.y));
                        F[base] = make_float2(__fadd_rn(f0.x, tt.x), __fadd_rn(f0.y, tt.y));
                    }
                }
                __syncwarp();
            }
            /* magnitude spectrum, band limits, IIR (reference src/fsk.c:610-628) */
            float v[WB_NEQ];
#pragma unroll
            for (int q = 0; q < WB_NEQ; q++) {
                const int i = lane + 32 * q;
                v[q] = 0.0f;
                if (i < nh) {
                    const float2 Xf = F[i];
                    float pw = __fadd_rn(__fmul_rn(Xf.x, Xf.x), __fmul_rn(Xf.y, Xf.y));
                    if (i < p.f_min || i >= p.f_max - 1) pw = 0.0f;
                    est[q] = __fadd_rn(__fmul_rn(est[q], omt), __fmul_rn(__fsqrt_rn(pw), p.tc));
                    v[q] = est[q];
                }
            }
            /* M maxima with +-f_zero blanking (reference src/fsk.c:635-654), then ascending order */
            int freqi[M];
#pragma unroll
            for (int m = 0; m < M; m++) {
                float bv = 0.0f; int bi = 0;
#pragma unroll
                for (int q = 0; q < WB_NEQ; q++)
                    if (lane + 32 * q < nh && v[q] > bv) { bv = v[q]; bi = lane + 32 * q; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                const int lo = max(bi - p.f_zero, 0), hi = min(bi + p.f_zero, Ndft);
#pragma unroll
                for (int q = 0; q < WB_NEQ; q++) {
                    const int i = lane + 32 * q;
                    if (i >= lo && i < hi) v[q] = 0.0f;
                }
                freqi[m] = bi;
            }
#pragma unroll
            for (int i = 1; i < M; i++) {       /* insertion sort, M <= 4 */
#pragma unroll
                for (int j = i; j > 0; j--)
                    if (freqi[j - 1] > freqi[j]) { const int t = freqi[j]; freqi[j] = freqi[j - 1]; freqi[j - 1] = t; }
            }
            /* the samples to stash for the next frame (reference src/fsk.c:851) sit where the mixer
               products are about to land: lift them into registers */
#pragma unroll
            for (int q = 0; q < WB_NST; q++)
                if (lane + 32 * q < nstash) stash[q] = X[nin + lane + 32 * q];
            if (lane == 0) {
                wb_fsk_sc &c = sc[warp];
                const bool first = c.pb[0] == 0;         /* fsk->f_est[0] < 1, reference src/fsk.c:729 */
#pragma unroll
                for (int m = 0; m < M; m++) { c.nb[m] = (short)freqi[m]; if (first) c.pb[m] = (short)freqi[m]; }
                c.nin = nin; c.flags = 1;
            }
            /* next frame: global loads now, consumed at the top of the next iteration */
#pragma unroll
            for (int q = 0; q < NPRE; q++) {
                const int n = lane + 32 * q;
                pre_lo[q] = pre_hi[q] = 0u;
                if (n < p.nmax && pos_next + n < fill) wb_load_raw<CF32>(fmt, in, pos_next + n, pre_lo[q], pre_hi[q]);
            }
        } else if (lane == 0) {
            sc[warp].flags = 0;
        }
        __syncthreads();

        /* ================= B1: warp 0, lane = (tone, stream): oscillator + down-mix ================= */
        if (warp == 0 && lane < M * spb) {
            const int m = lane / spb, s = lane - m * spb;
            wb_fsk_sc &c = sc[s];
            if (c.flags & 1) {
                float2 *Xs = reinterpret_cast<float2 *>(regions + (size_t)s * p.sreg);
                const int fnin = c.nin, nold = p.Nmem - fnin;
                const int nin_idx = (fnin < p.N) ? 0 : (fnin == p.N ? 1 : 2);
                const int pb = c.pb[m], nbn = c.nb[m];
                float2 ph = c.phi_c[m];
                ph = wb_cmul2(__ldg(&p.back[nin_idx * nh + pb]), ph);     /* reference src/fsk.c:756-759 */
                float2 d = __ldg(&p.dphi[pb]);
                const float2 dnew = __ldg(&p.dphi[nbn]);
                const float2 *src = Xs + (nstash - nold);
                float2 *dst = (m == 0) ? Xs + (nstash - nold) : Xs + p.xlen + (m - 1) * p.ylen;
                const int nsteps = p.nsteps;
                const int nslow = 2 * p.Ts + p.Ts / 2 + 1;                /* > every possible nold */
                int n = 0;
#pragma unroll 1
                for (; n < nslow; n++) {
                    if (n == nold) {        /* old -> new samples: comp_normalize + new tone, src/fsk.c:787-788 */
                        const float av = __fsqrt_rn(__fadd_rn(__fmul_rn(ph.x, ph.x), __fmul_rn(ph.y, ph.y)));
                        ph.x = __fdiv_rn(ph.x, av); ph.y = __fdiv_rn(ph.y, av);
                        d = dnew;
                    }
                    WB_MIX_STEP(src, dst, n);
                }
#pragma unroll 1
                for (; n + 8 <= nsteps; n += 8) {
#pragma unroll
                    for (int j = 0; j < 8; j++) WB_MIX_STEP(src + n, dst + n, j);
                }
#pragma unroll 1
                for (; n < nsteps; n++) WB_MIX_STEP(src, dst, n);
                c.phi_c[m] = ph;
            }
        }
        __syncthreads();

        /* ================= B2: stream warps, lanes = integrator outputs ================= */
        if (active) {
            /* f_int[m][i] = sum of the Ts ring-buffer slots after mixer step i*step + Ts - 1, added in slot
               order (reference src/fsk.c:835-838): slot j holds the product of the step n in
               [i*step, i*step + Ts) with n mod Ts == j.  In place: output i overwrites product i. */
            const int step = p.step;
            int r = (lane * step) % TS;                  /* (i * step) mod Ts for this lane's output */
            const int rinc = (32 * step) % TS;
            for (int i0 = 0; i0 < p.nint; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < p.nint;
                float2 f[M];
                float e = 0.0f;
                if (valid) {
                    const int n0 = i * step;
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        const float2 *P = (m == 0) ? X + xo + n0 : Y + (m - 1) * p.ylen + n0;
                        float sr = 0.0f, si = 0.0f;
#pragma unroll
                        for (int j = 0; j < TS; j++) {
                            int off = j - r;
                            if (off < 0) off += TS;
                            const float2 vv = P[off];
                            if (j == 0) { sr = vv.x; si = vv.y; }
                            else { sr = __fadd_rn(sr, vv.x); si = __fadd_rn(si, vv.y); }
                        }
                        f[m] = make_float2(sr, si);
                        const float pw = __fadd_rn(__fmul_rn(sr, sr), __fmul_rn(si, si));
                        e = (m == 0) ? pw : __fadd_rn(e, pw);        /* reference src/fsk.c:864-867 */
                    }
                }
                __syncwarp();
                if (valid) {
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        float2 *P = (m == 0) ? X + xo : Y + (m - 1) * p.ylen;
                        P[i] = f[m];
                    }
                    E[i] = e;
                }
                __syncwarp();
                r += rinc;
                if (r >= TS) r -= TS;
            }
        }
        __syncthreads();

        /* ================= B3: warp 0, lane = (re/im, stream): fine-timing accumulation ================= */
        if (warp == 0) {
            const int cidx = (lane < spb) ? 0 : 1;
            const int s = min(lane - cidx * spb, spb - 1);
            float acc = 0.0f;
            if (lane < 2 * spb) {
                const float *Es = reinterpret_cast<const float *>(
                    reinterpret_cast<const float2 *>(regions + (size_t)s * p.sreg) + p.xlen + p.blen);
                const float *pf = reinterpret_cast<const float *>(p.pft) + cidx;
                int i = 0;
#pragma unroll 1
                for (; i + 8 <= p.nint; i += 8) {
                    float ev[8], pv[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) { ev[j] = Es[i + j]; pv[j] = __ldg(pf + 2 * (i + j)); }
#pragma unroll
                    for (int j = 0; j < 8; j++) acc = __fadd_rn(acc, __fmul_rn(ev[j], pv[j]));   /* src/fsk.c:870 */
                }
                for (; i < p.nint; i++) acc = __fadd_rn(acc, __fmul_rn(Es[i], __ldg(pf + 2 * i)));
            }
            const float tcr = __shfl_sync(0xffffffffu, acc, s);
            const float tci = __shfl_sync(0xffffffffu, acc, spb + s);
            if (lane < spb) {
                wb_fsk_sc &c = sc[s];
                if (c.flags & 1) {
                    const bool nan = isnan(tcr) || isnan(tci);     /* reference src/fsk.c:878-880 */
                    if (!nan) {
                        const float norm = (float)((double)wb_atan2f(tci, tcr) / 6.283185307179586);
                        const float rx_timing = __fmul_rn(norm, (float)p.P);
                        const float dn = __fsub_rn(norm, c.norm);
                        c.norm = norm;
                        if ((double)fabsf(dn) < .2) {
                            const float appm = (float)(1e6 * (double)dn / (double)(float)p.Nsym);
                            c.ppm = (float)(.9 * (double)c.ppm + .1 * (double)appm);
                        }
                        if ((double)norm > 0.25) c.nin_next = p.N + p.Ts / 2;
                        else if ((double)norm < -0.25) c.nin_next = p.N - p.Ts / 2;
                        else c.nin_next = p.N;
                        const int low = (int)floorf(rx_timing);
                        c.low = low;
                        c.fract = __fsub_rn(rx_timing, (float)low);
                        c.high = (int)ceilf(rx_timing);
                        c.rx_timing = rx_timing;
                    } else {
                        c.flags = 3;
                        c.nin_next = c.nin;
                    }
                }
            }
        }
        __syncthreads();

        /* ================= C: stream warps, lanes = symbols ================= */
        if (active) {
            wb_fsk_sc &c = sc[warp];
            float *out = sdrow + n_out;
            if (!(c.flags & 2)) {
                const int low = c.low, high = c.high;
                const float fract = c.fract, omf = __fsub_rn(1.0f, fract);
                for (int i = lane; i < p.Nsym; i += 32) {
                    const int stt = (i + 1) * p.P;
                    float tm[M];
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        const float2 *fi = (m == 0) ? X + xo : Y + (m - 1) * p.ylen;
                        const float2 lo = fi[stt + low], hi = fi[stt + high];
                        const float tr = __fadd_rn(__fmul_rn(omf, lo.x), __fmul_rn(fract, hi.x));
                        const float ti = __fadd_rn(__fmul_rn(omf, lo.y), __fmul_rn(fract, hi.y));
                        tm[m] = __fsqrt_rn(__fadd_rn(__fmul_rn(tr, tr), __fmul_rn(ti, ti)));
                    }
                    if (M == 2) {
                        out[i] = __fsub_rn(tm[0], tm[1]);                    /* reference src/fsk.c:966 */
                    } else {                                                  /* reference src/fsk.c:969-979 */
                        float b1 = -tm[0], b0 = -tm[0];
                        b1 = __fadd_rn(b1, tm[1]);  b0 = __fadd_rn(b0, -tm[1]);
                        b1 = __fadd_rn(b1, -tm[2]); b0 = __fadd_rn(b0, tm[2]);
                        b1 = __fadd_rn(b1, tm[3]);  b0 = __fadd_rn(b0, tm[3]);
                        out[2 * i + 1] = b1; out[2 * i] = b0;
                    }
                }
            } else {
                /* NaN guard: the reference returns before writing, so the caller's buffer still holds
                   the previous frame's values (zeros before the first frame) */
                for (int i = lane; i < p.Nbits; i += 32)
                    out[i] = (n_out >= (unsigned)p.Nbits) ? out[i - p.Nbits] : 0.0f;
            }
            __syncwarp();
            /* samp_old for the next frame */
#pragma unroll
            for (int q = 0; q < WB_NST; q++)
                if (lane + 32 * q < nstash) X[lane + 32 * q] = stash[q];
            if (lane == 0) {
#pragma unroll
                for (int m = 0; m < M; m++) c.pb[m] = c.nb[m];              /* fsk->f_est = this frame's, :846 */
                if (a.frame_log && frames < (unsigned)a.log_cap) {
                    float *l = a.frame_log + ((size_t)sg * a.log_cap + frames) * 8;
                    l[0] = (float)nin;
                    for (int m = 0; m < 4; m++) l[1 + m] = m < M ? (float)c.nb[m] : 0.0f;
                    l[5] = c.norm; l[6] = c.ppm; l[7] = c.rx_timing;
                }
            }
            n_out += p.Nbits;
            pos = pos_next;
            nin = c.nin_next;
            frames++;
            __syncwarp();
        }
    }

    /* ---- write the state back ---- */
    if (have) {
        const wb_fsk_sc &c = sc[warp];
#pragma unroll
        for (int q = 0; q < WB_NEQ; q++)
            if (lane + 32 * q < nh) st->fft_est[lane + 32 * q] = est[q];
#pragma unroll
        for (int q = 0; q < WB_NST; q++)
            if (lane + 32 * q < nstash) st->samp_old[lane + 32 * q] = X[lane + 32 * q];
        const unsigned long long rem = fill - pos;
        const unsigned long long dstpos = (unsigned long long)a.headroom - rem;   /* remainder ends at the headroom mark */
        if (a.compact && rem > 0 && pos > dstpos) {
            /* less than one frame is left over: park it right before the headroom mark so that every
               stream's next wb_feed lands at the same row offset (one strided copy for all streams).
               dst < src, ascending copy through registers. */
            const int bps = p.in_bps;
            unsigned char *row = const_cast<unsigned char *>(in);
            const unsigned long long nbytes = rem * bps, sb = pos * bps, db = dstpos * bps;
            for (unsigned long long off = 0; off < nbytes; off += 32 * 8) {
                unsigned char tmp[8];
                const unsigned long long o = off + (unsigned long long)lane * 8;
#pragma unroll
                for (int b = 0; b < 8; b++) tmp[b] = (o + b < nbytes) ? row[sb + o + b] : 0;
                __syncwarp();
#pragma unroll
                for (int b = 0; b < 8; b++) if (o + b < nbytes) row[db + o + b] = tmp[b];
                __syncwarp();
            }
        }
        if (lane == 0) {
            for (int m = 0; m < M; m++) { st->phi_c[m] = c.phi_c[m]; st->fbin[m] = c.pb[m]; }
            st->norm_rx_timing = c.norm; st->ppm = c.ppm; st->rx_timing = c.rx_timing;
            st->nin = nin; st->frames += frames;
            if (a.compact) { st->in_pos = dstpos; st->in_fill = a.headroom; }
            else { st->in_pos = pos; st->in_fill = fill; }
            wb_cursor &cu = a.cursor[sg];
            cu.in_fill = rem;
            cu.consumed = pos - pos0;
            cu.n_sd = n_out;
            cu.nin = nin;
        }
    }
}

#endif /* WB_FSK_KERNEL_CUH */
