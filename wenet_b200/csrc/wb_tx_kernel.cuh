/*
 * wb_tx_kernel.cuh -- transmit side on the device: synthesise the signals the
 * receive path decodes directly in the engine's HBM input rows, instead of staging them from the host.
 *
 *   T1 wb_tx_frame_kernel   payload (256 B) -> CRC16 -> RA-LDPC parity -> on-air frame bits
 *        reference tx/PacketTX.py:123-137 (frame_packet), tx/ldpc_encoder.py:42-52 + src/mpdecode_core.c:72-91
 *        (encode), tx/radio_wrappers.py:385-417 (v2 scramble, MSB first) / :502-559 (v1: UART start/stop, LSB first)
 *   T2 wb_tx_mod_kernel     bits -> continuous-phase M-FSK samples (+ AWGN), written where wb_feed would put them
 *        reference src/fsk.c:1162-1204 (fsk_mod_c): tx_phase_c *= dosc_f[sym] every sample, output 2*tx_phase_c,
 *        comp_normalize after every Nsym symbols; the same float operations in the same order, so without noise
 *        the samples are bit-identical to the reference modulator's.
 *        AWGN as benchmarking/generate_lowsnr.py:70-89: complex Gaussian of variance Ts/(Eb/N0) on the unit-
 *        amplitude signal (counter-based Philox + Box-Muller: reproducible per (seed, stream, sample), not the
 *        numpy generator), then every stream scaled to max |y| = 1 (T3).
 *   T3 wb_tx_scale_kernel   peak normalisation + conversion to the engine's input format
 *
 * One thread per stream in T2: the oscillator is a dependent chain per stream and thousands of streams run side by
 * side; the frame bits are built by one warp per packet.
 */
#ifndef WB_TX_KERNEL_CUH
#define WB_TX_KERNEL_CUH

#include "wb_internal.h"

#define WB_TX_HEAD_BYTES 20                      /* 16 x 0x55 + AB CD EF 01 */
#define WB_TX_RAW_BYTES (WB_TX_HEAD_BYTES + WB_PKT_BODY_BYTES)

struct wb_tx_args {
    const uint8_t *payloads;     /* [n_streams][n_packets][256] */
    uint8_t *bits;               /* [n_streams][bits_stride] one byte per on-air bit */
    unsigned long long bits_stride;
    int n_streams, n_packets, framing;
    int lead_in, gap, frame_bits;
    const uint16_t *hrows;       /* [516][12] 0-based data columns of H1 */
    const uint8_t *scramble;     /* [1000] */
};

/* T1: one warp per packet */
__global__ void __launch_bounds__(128)
wb_tx_frame_kernel(wb_tx_args a)
{
    __shared__ uint8_t raw_s[4][WB_TX_RAW_BYTES + 1];
    __shared__ uint8_t par_s[4][WB_NPAR + 28];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long pk = (long long)blockIdx.x * 4 + w;
    if (pk >= (long long)a.n_streams * a.n_packets) return;
    const int s = (int)(pk / a.n_packets), k = (int)(pk - (long long)s * a.n_packets);
    uint8_t *raw = raw_s[w], *par = par_s[w];
    const uint8_t *pl = a.payloads + (size_t)pk * WB_PACKET_BYTES;
    for (int i = lane; i < WB_TX_HEAD_BYTES; i += 32) raw[i] = (i < 16) ? 0x55 : (i == 16 ? 0xAB : (i == 17 ? 0xCD : (i == 18 ? 0xEF : 0x01)));
    for (int i = lane; i < WB_PACKET_BYTES; i += 32) raw[WB_TX_HEAD_BYTES + i] = pl[i];
    __syncwarp();
    /* CRC16-CCITT-FALSE (init 0xFFFF, poly 0x1021), little endian after the payload: tx/PacketTX.py:131.
       256 bytes, 8 per lane, combined with the linearity of the CRC: crc(A || B) = shift(crc(A), |B|) ^ crc0(B);
       simpler and cheap enough here: lane 0 runs the bitwise CRC */
    if (lane == 0) {
        unsigned crc = 0xFFFF;
        for (int i = 0; i < WB_PACKET_BYTES; i++) {
            crc ^= (unsigned)raw[WB_TX_HEAD_BYTES + i] << 8;
#pragma unroll
            for (int b = 0; b < 8; b++) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xFFFFu : (crc << 1) & 0xFFFFu;
        }
        raw[WB_TX_HEAD_BYTES + 256] = (uint8_t)(crc & 0xFF);
        raw[WB_TX_HEAD_BYTES + 257] = (uint8_t)(crc >> 8);
    }
    __syncwarp();
    /* parity, src/mpdecode_core.c:72-91: pbits[p] = (sum of the 12 data bits of check p + pbits[p-1]) & 1, i.e. a
       running XOR: lane l takes checks 17l .. 17l+16, then an exclusive XOR scan over the lanes */
    {
        const int p0 = 17 * lane;
        unsigned loc[17];
        unsigned run = 0;
#pragma unroll
        for (int j = 0; j < 17; j++) {
            const int p = p0 + j;
            unsigned x = 0;
            if (p < WB_NPAR) {
#pragma unroll
                for (int t = 0; t < WB_ROWW; t++) {
                    const int col = a.hrows[p * WB_ROWW + t];
                    x ^= (raw[WB_TX_HEAD_BYTES + (col >> 3)] >> (7 - (col & 7))) & 1u;      /* ibits = np.unpackbits: MSB first */
                }
            }
            run ^= x;
            loc[j] = run;
        }
        unsigned pre = run;                      /* inclusive scan of the lane totals */
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre ^= t;
        }
        pre ^= run;                              /* exclusive */
#pragma unroll
        for (int j = 0; j < 17; j++) par[p0 + j] = (uint8_t)((loc[j] ^ pre) & 1u);
    }
    __syncwarp();
    /* 516 parity bits -> 65 bytes, MSB first, zero padded (np.packbits, tx/ldpc_encoder.py:52) */
    for (int i = lane; i < 65; i += 32) {
        unsigned b = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) b = (b << 1) | ((8 * i + t < WB_NPAR) ? par[8 * i + t] : 0u);
        raw[WB_TX_HEAD_BYTES + 258 + i] = (uint8_t)b;
    }
    __syncwarp();
    uint8_t *out = a.bits + (size_t)s * a.bits_stride + a.lead_in + (size_t)k * (a.frame_bits + a.gap);
    if (a.framing == WB_FRAMING_V2) {
        /* body ^= scramble table (125 bytes = the 1000 signs packed MSB first); bits MSB first */
        for (int i = lane; i < WB_TX_RAW_BYTES * 8; i += 32) {
            const int B = i >> 3, t = i & 7;
            unsigned bit = (raw[B] >> (7 - t)) & 1u;
            if (B >= WB_TX_HEAD_BYTES) bit ^= a.scramble[(8 * ((B - WB_TX_HEAD_BYTES) % 125) + t) % WB_SCRAMBLE_LEN];
            out[i] = (uint8_t)bit;
        }
    } else {
        /* UART: start 0, 8 data bits LSB first, stop 1 */
        for (int i = lane; i < WB_TX_RAW_BYTES * 10; i += 32) {
            const int B = i / 10, t = i - 10 * B;
            out[i] = (t == 0) ? 0 : (t == 9 ? 1 : (uint8_t)((raw[B] >> (t - 1)) & 1u));
        }
    }
}

/* T1 for framing NONE: the payload bytes as they are, MSB first */
__global__ void __launch_bounds__(256)
wb_tx_raw_bits_kernel(wb_tx_args a)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)WB_PACKET_BYTES * 8;
    if (g >= (long long)a.n_streams * a.n_packets * per) return;
    const long long pk = g / per;
    const int i = (int)(g - pk * per);
    const int s = (int)(pk / a.n_packets), k = (int)(pk - (long long)s * a.n_packets);
    const uint8_t B = a.payloads[(size_t)pk * WB_PACKET_BYTES + (i >> 3)];
    a.bits[(size_t)s * a.bits_stride + a.lead_in + (size_t)k * (a.frame_bits + a.gap) + i] = (uint8_t)((B >> (7 - (i & 7))) & 1u);
}

/* counter-based Philox4x32-10 (Salmon et al., SC11) */
__device__ __forceinline__ uint4 wb_philox(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

struct wb_mod_args {
    const uint8_t *bits;         /* [n_streams][bits_stride] */
    unsigned long long bits_stride;
    float *work;                 /* [n_streams][work_stride] float2 samples before scaling (cf32 engines: the input rows) */
    unsigned long long work_stride;   /* in float2 */
    const unsigned long long *work_off;   /* [n_streams] first sample of each row, or NULL */
    float *peak;                 /* [n_streams] max |y| of the noisy stream */
    int n_streams, M, Ts, Nsym;
    unsigned long long n_calls;  /* fsk_mod_c calls per stream */
    float2 dosc[WB_MAXM];        /* comp_exp_j(2 pi (f1 + m fs)/Fs), host glibc */
    float sigma;                 /* per-component noise std on the unit-amplitude signal; 0 = none */
    const float *sigma_per_stream;   /* [n_streams] or NULL */
    int unit;                    /* 1: halve the modulator output (unit amplitude) before noise */
    unsigned long long seed;
};

/* T2: one thread per stream */
__global__ void __launch_bounds__(32)
wb_tx_mod_kernel(wb_mod_args a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_streams) return;
    const uint8_t *bits = a.bits + (size_t)s * a.bits_stride;
    float2 *out = reinterpret_cast<float2 *>(a.work) + (size_t)s * a.work_stride + (a.work_off ? a.work_off[s] : 0ULL);
    float2 ph = make_float2(1.0f, 0.0f);                          /* comp_exp_j(0), src/fsk.c:237 */
    const int bps = (a.M == 2) ? 1 : 2;
    float pk = 0.0f;
    const float sigma = a.sigma_per_stream ? a.sigma_per_stream[s] : a.sigma;
    unsigned long long n = 0, bit_i = 0;
    for (unsigned long long c = 0; c < a.n_calls; c++) {
        for (int i = 0; i < a.Nsym; i++) {
            int sym = 0;
            for (int b = 0; b < bps; b++) { sym = (sym << 1) | (bits[bit_i] == 1 ? 1 : 0); bit_i++; }
            const float2 dph = a.dosc[sym];
            for (int j = 0; j < a.Ts; j++) {
                /* tx_phase_c = cmult(tx_phase_c, dph); fsk_out = fcmult(2, tx_phase_c): src/fsk.c:1192-1193 */
                float2 t;
                t.x = __fsub_rn(__fmul_rn(ph.x, dph.x), __fmul_rn(ph.y, dph.y));
                t.y = __fadd_rn(__fmul_rn(ph.x, dph.y), __fmul_rn(ph.y, dph.x));
                ph = t;
                float2 y = a.unit ? ph : make_float2(__fmul_rn(2.0f, ph.x), __fmul_rn(2.0f, ph.y));
                if (sigma > 0.0f) {
                    const uint4 r = wb_philox(make_uint4((unsigned)n, (unsigned)(n >> 32), (unsigned)s, 0x57454e45u),
                                              make_uint2((unsigned)a.seed, (unsigned)(a.seed >> 32)));
                    const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
                    const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
                    const float rad = sigma * sqrtf(-2.0f * logf(u1));
                    float sn, cs;
                    sincosf(6.28318530717958647692f * u2, &sn, &cs);
                    y.x += rad * cs; y.y += rad * sn;
                    pk = fmaxf(pk, y.x * y.x + y.y * y.y);
                }
                out[n++] = y;
            }
        }
        /* comp_normalize at the end of every fsk_mod_c call, src/fsk.c:1198 / src/comp_prim.h:133-139 */
        const float av = __fsqrt_rn(__fadd_rn(__fmul_rn(ph.x, ph.x), __fmul_rn(ph.y, ph.y)));
        ph.x = __fdiv_rn(ph.x, av); ph.y = __fdiv_rn(ph.y, av);
    }
    a.peak[s] = sqrtf(pk);
}

/* T3: y / max|y| (only with noise, generate_lowsnr.py:85-87) and conversion into the input rows */
__global__ void __launch_bounds__(256)
wb_tx_scale_kernel(const float *work, unsigned long long work_stride, const unsigned long long *work_off, const float *peak, int normalise,
                   unsigned char *in, unsigned long long in_stride, const unsigned long long *row_off, int fmt,
                   unsigned long long nsamp, int n_streams)
{
    const int s = blockIdx.y;
    if (s >= n_streams) return;
    const float2 *src = reinterpret_cast<const float2 *>(work) + (size_t)s * work_stride + (work_off ? work_off[s] : 0ULL);
    const float g = (normalise && peak[s] > 0.0f) ? peak[s] : 1.0f;
    unsigned char *row = in + (size_t)s * in_stride;
    const unsigned long long off = row_off[s];
    for (unsigned long long n = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; n < nsamp;
         n += (unsigned long long)gridDim.x * blockDim.x) {
        float2 y = src[n];
        if (normalise) { y.x = __fdiv_rn(y.x, g); y.y = __fdiv_rn(y.y, g); }
        if (fmt == WB_FMT_CF32) {
            reinterpret_cast<float2 *>(row)[off + n] = y;
        } else if (fmt == WB_FMT_CS16) {                       /* the demodulator divides by 1000 (FDMDV_SCALE) */
            short2 v;
            v.x = (short)fminf(fmaxf(rintf(y.x * 1000.0f), -32768.0f), 32767.0f);
            v.y = (short)fminf(fmaxf(rintf(y.y * 1000.0f), -32768.0f), 32767.0f);
            reinterpret_cast<short2 *>(row)[off + n] = v;
        } else if (fmt == WB_FMT_CU8) {
            uchar2 v;
            v.x = (unsigned char)fminf(fmaxf(rintf(y.x * 127.0f + 127.0f), 0.0f), 255.0f);
            v.y = (unsigned char)fminf(fmaxf(rintf(y.y * 127.0f + 127.0f), 0.0f), 255.0f);
            reinterpret_cast<uchar2 *>(row)[off + n] = v;
        } else {                                               /* real s16 */
            reinterpret_cast<short *>(row)[off + n] = (short)fminf(fmaxf(rintf(y.x * 1000.0f), -32768.0f), 32767.0f);
        }
    }
}

#endif /* WB_TX_KERNEL_CUH */
