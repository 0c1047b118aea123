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
 * and inside a frame the tone oscillators (phi_c *= dphi, 399 dependent complex products per tone)
 * and the fine-timing accumulator (392 dependent additions) are sequential too.  So:
 *
 *   CTA = SPB = 32/M streams, one "stream warp" per stream, frames in lock step.
 *   phase A (stream warps, lanes = samples / butterflies / bins):
 *           land the prefetched frame in shared memory, issue the prefetch of the next one
 *           (global loads into registers, consumed a whole frame later: HBM latency is off the
 *           critical path), window + 256-point FFT in the reference's butterfly order, spectrum IIR
 *           (kept in registers), warp-argmax peak picking.
 *   phase B (warp 0, lane = (tone, stream)): ALL the sequential work of the CTA's streams at once:
 *           oscillator recurrence, down-mix, Ts-tap ring buffer in registers, integrator output,
 *           |.|^2 summed over tones by shuffle, fine-timing accumulation (real part on the tone-0
 *           lane, imaginary part on the tone-1 lane), then atan2 / ppm / nin on the tone-0 lanes.
 *           32 dependent chains share every issue slot instead of one chain idling 31 lanes.
 *   phase C (stream warps, lanes = symbols): linear-interpolated resampling and soft decisions,
 *           48 (96) floats per frame written coalesced.
 *
 * Shared memory per stream: x[nstash + nmax] float2 (old + new samples; overwritten in place by tone
 * 0's integrator outputs, which trail the read pointer) and (M-1) * nint float2 for the other tones
 * (doubling as the FFT work buffer in phase A).  Stream regions are 8 bytes mod 128 apart so the 16
 * streams read by a half-warp in phase B fall in distinct banks.
 * HBM traffic: every input sample is read once (8 B as cf32), 4 B x Nbits/N written.
 */
#ifndef WB_FSK_KERNEL_CUH
#define WB_FSK_KERNEL_CUH

#include "wb_internal.h"
#include "wb_math.h"

#define WB_NPRE (WB_MAX_NIN / 32)
#define WB_NST ((WB_MAX_NSTASH + 31) / 32)
#define WB_NEQ (WB_MAX_NDFT / 2 / 32)

struct wb_fsk_sc {                 /* per-stream frame scalars in shared memory */
    float2 phi_c[WB_MAXM];
    int pb[WB_MAXM];               /* estimator bins in force before this frame (fsk->f_est) */
    int nb[WB_MAXM];               /* bins estimated from this frame */
    int nin, active, nin_next, nanflag;
    int low, high;
    float fract, norm, ppm, rx_timing;
    float ebno_db, snr_est;
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
__device__ __forceinline__ float2 wb_convert(int fmt, uint2 raw)
{
    float2 v;
    if (fmt == WB_FMT_CF32) {
        v.x = __uint_as_float(raw.x); v.y = __uint_as_float(raw.y);
    } else if (fmt == WB_FMT_CU8) {
        /* ((float)u8 - 127.0) / 128.0 in double, then to float: exact, so float arithmetic gives the same */
        v.x = __fmul_rn(__fsub_rn((float)(raw.x & 0xffu), 127.0f), 0.0078125f);
        v.y = __fmul_rn(__fsub_rn((float)((raw.x >> 8) & 0xffu), 127.0f), 0.0078125f);
    } else if (fmt == WB_FMT_CS16) {
        v.x = __fdiv_rn((float)(short)(raw.x & 0xffffu), 1000.0f);
        v.y = __fdiv_rn((float)(short)(raw.x >> 16), 1000.0f);
    } else {
        v.x = __fdiv_rn((float)(short)(raw.x & 0xffffu), 1000.0f);
        v.y = 0.0f;
    }
    return v;
}

__device__ __forceinline__ uint2 wb_load_raw(int fmt, const unsigned char *p, unsigned long long idx)
{
    uint2 r = make_uint2(0u, 0u);
    if (fmt == WB_FMT_CF32) {
        r = __ldg(reinterpret_cast<const uint2 *>(p) + idx);
    } else if (fmt == WB_FMT_CS16) {
        r.x = __ldg(reinterpret_cast<const unsigned *>(p) + idx);
    } else {
        r.x = __ldg(reinterpret_cast<const unsigned short *>(p) + idx);
    }
    return r;
}

template <int M, int TS>
__global__ void __launch_bounds__(32 / M * 32)
wb_fsk_kernel(wb_fsk_params p, wb_fsk_args a)
{
    constexpr int SPB = 32 / M;
    extern __shared__ __align__(16) unsigned char wb_fsk_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sg = blockIdx.x * SPB + warp;
    const bool have = sg < a.n_streams;

    wb_fsk_sc *sc = reinterpret_cast<wb_fsk_sc *>(wb_fsk_raw);
    unsigned char *regions = wb_fsk_raw + ((sizeof(wb_fsk_sc) * SPB + 127) / 128) * 128;
    float2 *xbuf = reinterpret_cast<float2 *>(regions + (size_t)warp * p.sreg);
    float2 *bbuf = xbuf + p.xlen;
    const int Ndft = p.Ndft, nh = Ndft >> 1, nstash = p.nstash, fmt = p.in_fmt;

    /* ---- per-stream state -> registers / shared memory ---- */
    wb_stream_state *st = have ? a.state + sg : nullptr;
    const unsigned char *in = a.in + (size_t)(have ? sg : 0) * a.in_stride;
    float *sdrow = a.sd + (size_t)(have ? sg : 0) * a.sd_stride + WB_CARRY_CAP;
    unsigned long long pos = 0, fill = 0, pos0 = 0, frames = 0;
    int nin = p.N;
    unsigned n_out = 0;
    float est[WB_NEQ];
    float2 stash[WB_NST];
    uint2 pre[WB_NPRE];
#pragma unroll
    for (int q = 0; q < WB_NEQ; q++) est[q] = 0.0f;
#pragma unroll
    for (int q = 0; q < WB_NST; q++) stash[q] = make_float2(0.0f, 0.0f);
    if (have) {
        pos = pos0 = st->in_pos; fill = st->in_fill; nin = st->nin; frames = st->frames;
#pragma unroll
        for (int q = 0; q < WB_NEQ; q++)
            if (lane + 32 * q < nh) est[q] = st->fft_est[lane + 32 * q];
#pragma unroll
        for (int q = 0; q < WB_NST; q++)
            if (lane + 32 * q < nstash) xbuf[lane + 32 * q] = st->samp_old[lane + 32 * q];
        if (lane == 0) {
            wb_fsk_sc &c = sc[warp];
            for (int m = 0; m < M; m++) { c.phi_c[m] = st->phi_c[m]; c.pb[m] = st->fbin[m]; c.nb[m] = 0; }
            c.norm = st->norm_rx_timing; c.ppm = st->ppm; c.rx_timing = st->rx_timing;
            c.nin = nin; c.nin_next = nin; c.active = 0; c.nanflag = 0;
            c.ebno_db = 0.0f; c.snr_est = 0.0f;
        }
    } else if (lane == 0) {
        wb_fsk_sc &c = sc[warp];
        for (int m = 0; m < M; m++) { c.phi_c[m] = make_float2(1.0f, 0.0f); c.pb[m] = 0; c.nb[m] = 0; }
        c.nin = p.N; c.active = 0; c.nin_next = p.N; c.nanflag = 0; c.norm = 0.0f; c.ppm = 0.0f;
    }
    /* prefetch the first frame */
#pragma unroll
    for (int q = 0; q < WB_NPRE; q++) {
        int n = lane + 32 * q;
        pre[q] = (have && n < p.nmax && pos + n < fill) ? wb_load_raw(fmt, in, pos + n) : make_uint2(0u, 0u);
    }

    const float omt = __fsub_rn(1.0f, p.tc);

    for (;;) {
        const bool active = have && (pos + (unsigned long long)nin <= fill) && (n_out + (unsigned)p.Nbits <= a.sd_cap);
        if (!__syncthreads_or(active)) break;
        unsigned long long pos_next = pos + nin;

        /* ================= phase A: stream warps ================= */
        if (active) {
#pragma unroll
            for (int q = 0; q < WB_NPRE; q++) {
                int n = lane + 32 * q;
                if (n < p.nmax) xbuf[nstash + n] = wb_convert(fmt, pre[q]);
            }
#pragma unroll
            for (int q = 0; q < WB_NPRE; q++) {
                int n = lane + 32 * q;
                pre[q] = (n < p.nmax && pos_next + n < fill) ? wb_load_raw(fmt, in, pos_next + n) : make_uint2(0u, 0u);
            }
            __syncwarp();
            /* window the first nin - Ndft samples, zero-pad, in the leaf order of the DIT recursion
               (reference src/fsk.c:583-603, src/kiss_fft.c:238-306) */
            const int nwin = min(nin - Ndft, Ndft);
            float2 *F = bbuf;
            for (int o = lane; o < Ndft; o += 32) {
                int idx = __ldg(&p.perm[o]);
                float2 v = make_float2(0.0f, 0.0f);
                if (idx < nwin) {
                    float h = __ldg(&p.hann[idx]);
                    float2 x = xbuf[nstash + idx];
                    v.x = __fmul_rn(h, x.x); v.y = __fmul_rn(h, x.y);
                }
                F[o] = v;
            }
            __syncwarp();
            for (int L = 0; L < p.n_levels; L++) {
                const int pp = p.lev_p[L], mm = p.lev_m[L], fs = p.lev_fstride[L];
                const int nbf = Ndft / pp;
                for (int t = lane; t < nbf; t += 32) {
                    int blk = t / mm, k = t - blk * mm;
                    int base = blk * pp * mm + k;
                    if (pp == 4) {      /* reference src/kiss_fft.c:44-90, forward */
                        float2 f0 = F[base], f1 = F[base + mm], f2 = F[base + 2 * mm], f3 = F[base + 3 * mm];
                        float2 s0 = wb_cmul2(f1, __ldg(&p.tw[k * fs]));
                        float2 s1 = wb_cmul2(f2, __ldg(&p.tw[2 * k * fs]));
                        float2 s2 = wb_cmul2(f3, __ldg(&p.tw[3 * k * fs]));
                        float2 s5 = make_float2(__fsub_rn(f0.x, s1.x), __fsub_rn(f0.y, s1.y));
                        float2 aa = make_float2(__fadd_rn(f0.x, s1.x), __fadd_rn(f0.y, s1.y));
                        float2 s3 = make_float2(__fadd_rn(s0.x, s2.x), __fadd_rn(s0.y, s2.y));
                        float2 s4 = make_float2(__fsub_rn(s0.x, s2.x), __fsub_rn(s0.y, s2.y));
                        F[base + 2 * mm] = make_float2(__fsub_rn(aa.x, s3.x), __fsub_rn(aa.y, s3.y));
                        F[base] = make_float2(__fadd_rn(aa.x, s3.x), __fadd_rn(aa.y, s3.y));
                        F[base + mm] = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
                        F[base + 3 * mm] = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
                    } else {            /* reference src/kiss_fft.c:22-42 */
                        float2 f0 = F[base], f1 = F[base + mm];
                        float2 tt = wb_cmul2(f1, __ldg(&p.tw[k * fs]));
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
                int i = lane + 32 * q;
                v[q] = 0.0f;
                if (i < nh) {
                    float2 X = F[i];
                    float pw = __fadd_rn(__fmul_rn(X.x, X.x), __fmul_rn(X.y, X.y));
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
                    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                int lo = max(bi - p.f_zero, 0), hi = min(bi + p.f_zero, Ndft);
#pragma unroll
                for (int q = 0; q < WB_NEQ; q++) {
                    int i = lane + 32 * q;
                    if (i >= lo && i < hi) v[q] = 0.0f;
                }
                freqi[m] = bi;
            }
#pragma unroll
            for (int i = 1; i < M; i++) {       /* insertion sort, M <= 4 */
#pragma unroll
                for (int j = i; j > 0; j--)
                    if (freqi[j - 1] > freqi[j]) { int t = freqi[j]; freqi[j] = freqi[j - 1]; freqi[j - 1] = t; }
            }
            /* the samples to stash for the next frame (reference src/fsk.c:851) sit where tone 0's
               integrator outputs are about to land: lift them into registers */
#pragma unroll
            for (int q = 0; q < WB_NST; q++)
                if (lane + 32 * q < nstash) stash[q] = xbuf[nin + lane + 32 * q];
            if (lane == 0) {
                wb_fsk_sc &c = sc[warp];
                const bool first = c.pb[0] == 0;         /* fsk->f_est[0] < 1, reference src/fsk.c:729 */
#pragma unroll
                for (int m = 0; m < M; m++) { c.nb[m] = freqi[m]; if (first) c.pb[m] = freqi[m]; }
                c.nin = nin; c.active = 1;
            }
        } else if (lane == 0) {
            sc[warp].active = 0;
        }
        __syncthreads();

        /* ================= phase B: warp 0, lane = (tone, stream) ================= */
        if (warp == 0) {
            const int m = lane / SPB, s = lane - m * SPB;
            wb_fsk_sc &c = sc[s];
            const int act = c.active;
            const float2 *xs_base = reinterpret_cast<const float2 *>(regions + (size_t)s * p.sreg);
            float2 *fo = (m == 0) ? const_cast<float2 *>(xs_base)
                                  : const_cast<float2 *>(xs_base) + p.xlen + (m - 1) * p.nint;
            const int fnin = c.nin, nold = p.Nmem - fnin;
            const int nin_idx = (fnin < p.N) ? 0 : (fnin == p.N ? 1 : 2);
            const int pb = c.pb[m], nbn = c.nb[m];
            float2 ph = c.phi_c[m];
            ph = wb_cmul2(__ldg(&p.back[nin_idx * nh + pb]), ph);     /* reference src/fsk.c:756-759 */
            float2 d = __ldg(&p.dphi[pb]);
            const float2 dnew = __ldg(&p.dphi[nbn]);
            const float2 *xs = xs_base + (nstash - nold);
            float2 ring[TS];
#pragma unroll
            for (int j = 0; j < TS; j++) ring[j] = make_float2(0.0f, 0.0f);
            float acc = 0.0f;
            int cnt = 0, iout = 0, n = 0;
            const int nsteps = p.nsteps, step1 = p.step - 1;
#pragma unroll 1
            for (int blk = 0; blk < p.Nsym + 2; blk++) {
#pragma unroll
                for (int j = 0; j < TS; j++, n++) {
                    if (n < nsteps) {
                        if (n == nold) {        /* old -> new samples: comp_normalize + new tone, src/fsk.c:787-788 */
                            float av = __fsqrt_rn(__fadd_rn(__fmul_rn(ph.x, ph.x), __fmul_rn(ph.y, ph.y)));
                            ph.x = __fdiv_rn(ph.x, av); ph.y = __fdiv_rn(ph.y, av);
                            d = dnew;
                        }
                        const float2 x = xs[n];
                        /* cmult(sample, cconj(phi)), reference src/fsk.c:794 */
                        ring[j].x = __fadd_rn(__fmul_rn(x.x, ph.x), __fmul_rn(x.y, ph.y));
                        ring[j].y = __fsub_rn(__fmul_rn(x.y, ph.x), __fmul_rn(x.x, ph.y));
                        ph = wb_cmul2(ph, d);
                        if (n >= TS - 1) {
                            if (cnt == 0) {
                                cnt = step1;
                                float sr = ring[0].x, si = ring[0].y;
#pragma unroll
                                for (int t = 1; t < TS; t++) { sr = __fadd_rn(sr, ring[t].x); si = __fadd_rn(si, ring[t].y); }
                                if (act) fo[iout] = make_float2(sr, si);
                                const float pw = __fadd_rn(__fmul_rn(sr, sr), __fmul_rn(si, si));
                                float e = __shfl_sync(0xffffffffu, pw, s);
#pragma unroll
                                for (int mm = 1; mm < M; mm++) e = __fadd_rn(e, __shfl_sync(0xffffffffu, pw, mm * SPB + s));
                                const float2 pf = __ldg(&p.pft[iout]);
                                acc = __fadd_rn(acc, __fmul_rn(e, m == 0 ? pf.x : pf.y));
                                iout++;
                            } else {
                                cnt--;
                            }
                        }
                    }
                }
            }
            const float tcr = __shfl_sync(0xffffffffu, acc, s);
            const float tci = __shfl_sync(0xffffffffu, acc, SPB + s);
            if (act) {
                c.phi_c[m] = ph;
                if (m == 0) {
                    const bool nan = isnan(tcr) || isnan(tci);     /* reference src/fsk.c:878-880 */
                    c.nanflag = nan;
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
                        c.nin_next = fnin;
                    }
                }
            }
        }
        __syncthreads();

        /* ================= phase C: stream warps, lanes = symbols ================= */
        if (active) {
            wb_fsk_sc &c = sc[warp];
            float *out = sdrow + n_out;
            if (!c.nanflag) {
                const int low = c.low, high = c.high;
                const float fract = c.fract, omf = __fsub_rn(1.0f, fract);
                for (int i = lane; i < p.Nsym; i += 32) {
                    const int stt = (i + 1) * p.P;
                    float tm[M];
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        const float2 *fi = (m == 0) ? xbuf : bbuf + (m - 1) * p.nint;
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
                if (lane + 32 * q < nstash) xbuf[lane + 32 * q] = stash[q];
            if (lane == 0) {
#pragma unroll
                for (int m = 0; m < M; m++) c.pb[m] = c.nb[m];              /* fsk->f_est = this frame's, :846 */
                if (a.frame_log && frames - st->frames < (unsigned long long)a.log_cap) {
                    float *l = a.frame_log + ((size_t)sg * a.log_cap + (size_t)(frames - st->frames)) * 8;
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
            if (lane + 32 * q < nstash) st->samp_old[lane + 32 * q] = xbuf[lane + 32 * q];
        unsigned long long rem = fill - pos;
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
                unsigned long long o = off + (unsigned long long)lane * 8;
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
            st->nin = nin; st->frames = frames;
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
