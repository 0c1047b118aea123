/*
 * wb_ldpc_kernel.cuh -- K3 (sd_to_llr statistics) and K4 (deframe gather + LLR scaling +
 * phi-domain sum-product decode + byte pack + CRC gate) for the (2580, 2064) H2064_516 code.
 *
 * Replaces, per codeword:
 *   reference src/drs232_ldpc.c:220-225 / src/wenet_ldpc.c:207   RS232 strip + bit reversal / descramble
 *   reference src/mpdecode_core.c:569-595   sd_to_llr
 *   reference src/mpdecode_core.c:152-379   init_c_v_nodes  (here: constant edge tables)
 *   reference src/mpdecode_core.c:385-489   SumProduct
 *   reference src/drs232_ldpc.c:234-259     pack 258 bytes, CRC16 gate
 *
 * K4 mapping: one CTA of 288 threads per codeword.  All messages of one codeword live in shared
 * memory for the whole decode: one float per Tanner-graph edge, v->c and c->v messages sharing
 * the same word (the check pass reads q and overwrites it with r, the variable pass reads r and
 * overwrites it with q), sign in the float's sign bit.  Edge (check j, slot k) is word k*516 + j:
 * the check pass (thread = check) is bank-conflict free; slots are in the reference's
 * c_nodes[j].subs[] order (12 H1 columns, parity j-1, parity j) so phi_sum accumulates in the
 * reference's order, and a variable adds its incoming messages in v_nodes[i].subs[] order
 * (H_cols order; checks q, q+1 for parity column q).  516 checks = 2 rounds of 288 threads,
 * 2580 variables = 9 rounds (2592 slots): no tail round.  The variable pass gathers from scattered words, so WHICH
 * data column a thread takes decides the bank conflicts: slot s = thread + 288 * round works on column
 * vedge[s].w = wb_vperm[s], an order chosen off line for few conflicts (tools/gen_vperm.py: 3.55 -> 1.83 wavefronts
 * per gather); LLRs are staged through shared memory in natural order and the hard decisions put back in natural
 * order at the end.
 * HBM traffic per codeword: 3230 (v1) or 2584 (v2) floats in, one 280-byte record out.
 */
#ifndef WB_LDPC_KERNEL_CUH
#define WB_LDPC_KERNEL_CUH

#include "wb_internal.h"
#include "wb_math.h"
#include "wb_phi0.h"

struct wb_ldpc_args {
    /* mode A: codewords located by the deframer in the soft-decision rows */
    const float *sd;            /* [n_streams][sd_stride] */
    unsigned long long sd_stride;
    const unsigned *jobs;       /* [n_streams][job_cap] */
    const double *c4;           /* [n_streams][job_cap] 4*estEsN0 */
    const wb_cursor *cur;       /* [n_streams] */
    wb_stream_state *st;        /* packet counters */
    int job_cap;
    int framing;
    /* mode B: n codewords of ready-made LLRs */
    const float *llr_in;        /* [n][2580] or NULL */
    /* outputs */
    wb_codeword *cw;            /* mode A: [n_streams][job_cap]; mode B: NULL */
    float *llr_out;             /* mode A + KEEP_LLR: [n_streams][job_cap][2580]; else NULL */
    uint8_t *bits_out;          /* mode B: [n][323] */
    int *iters_out, *pcc_out;   /* mode B: [n] */
    int max_iter;
    /* tables */
    const ushort4 *vedge;       /* [2064] by variable-pass slot: message words of the column's three edges, .w = the column */
    const uint16_t *crc_tab;    /* [2048] CRC contribution of payload bit i */
    unsigned crc0;              /* CRC of 256 zero bytes */
    const uint8_t *scramble;    /* [1000] 1 = negate */
    const wb_phi0_flag *lut;
};

/* symbol index inside a collected packet of codeword element c */
__device__ __forceinline__ int wb_cw_symbol(int framing, int c)
{
    /* v1: byte k = symbols 10k .. 10k+9 = start, d0..d7 (LSB first), stop; element 8k+j is data bit 7-j
       (reference src/drs232_ldpc.c:220-225: unpacked[k*8+j] = symbol_buf[10k + 8 - j]) */
    return framing == WB_FRAMING_V1 ? 10 * (c >> 3) + 8 - (c & 7) : c;
}

/* ---- K3: sd_to_llr statistics, one thread per codeword --------------------
 * reference src/mpdecode_core.c:577-593.  The two running sums are sequential double additions in
 * the reference; their order is kept (bit-exact LLRs) by giving each codeword to one thread and
 * running thousands of codewords side by side.  The kernel is latency bound (one dependent DADD per element and
 * chain, a double division per element), so the elements are taken eight at a time -- one received byte: eight
 * adjacent symbols of the row -- with the eight loads and the eight divisions in flight together and only the
 * additions in element order.  Two passes over the row are inherent (the mean comes before the first x). */
template <int FRAMING>
__device__ __forceinline__ void wb_llr_stats_load8(const float *row, int k, const uint8_t *scramble, double v[8])
{
    /* elements 8k .. 8k+7 of the codeword; v1: symbols 10k+8 .. 10k+1 (wb_cw_symbol), v2: 8k .. 8k+7 descrambled */
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (FRAMING == WB_FRAMING_V1) v[j] = (double)row[10 * k + 8 - j];
        else {
            const int c = 8 * k + j;
            double t = (double)row[c];
            if (FRAMING == WB_FRAMING_V2 && scramble[c % WB_SCRAMBLE_LEN]) t = -t;
            v[j] = t;
        }
    }
}

/* a / b for many a and one b: with y = RN(1 / b) (one real division per codeword),
       q0 = RN(a y);  q1 = RN(q0 + (a - b q0) y);  q = RN(q1 + (a - b q1) y)
   is the correctly rounded quotient, i.e. bit for bit the IEEE division the reference executes: q1 is within 1/2 ulp +
   2^-105 of a / b (a y is off by at most 2^-52 relatively, the first correction removes that to second order), the
   residual a - b q1 is then exact in one fma, and a faithful quotient corrected once more with a correctly rounded
   reciprocal rounds correctly (Markstein, "IA-64 and Elementary Functions", the final step of Itanium's division).
   Checked against a / b on 3.2e10 random pairs (a any float, b plausible means, extreme exponents, significands of all
   ones and of few bits, scaled sums of floats): no difference.  5 double-precision instructions instead of the ~13 of
   the division routine, per element of the second pass.  Only for a b whose reciprocal is a normal number (else the
   plain division); a = -0 gives +0 instead of -0, which the sums that follow cannot see (they start at +0). */
__device__ __forceinline__ double wb_quot(double a, double b, double y, bool fast)
{
    if (!fast) return a / b;
    const double q0 = __dmul_rn(a, y);
    const double q1 = __fma_rn(__fma_rn(-b, q0, a), y, q0);
    return __fma_rn(__fma_rn(-b, q1, a), y, q1);
}

template <int FRAMING>
__device__ __forceinline__ double wb_llr_stats_row(const float *row, const uint8_t *scramble)
{
    constexpr int n = WB_NCODE, NB = n / 8, TAIL = n - 8 * NB;
    double sum = 0.0, sumsq = 0.0;
#pragma unroll 2
    for (int k = 0; k < NB; k++) {
        double v[8];
        wb_llr_stats_load8<FRAMING>(row, k, scramble, v);
#pragma unroll
        for (int j = 0; j < 8; j++) sum += fabs(v[j]);
    }
    {
        double v[8];
        wb_llr_stats_load8<FRAMING>(row, NB, scramble, v);       /* the row holds >= 2584 symbols: in bounds */
#pragma unroll
        for (int j = 0; j < TAIL; j++) sum += fabs(v[j]);
    }
    const double mean = sum / (double)n;
    const double rcp = 1.0 / mean;
    const bool fast = mean > 1e-290 && mean < 1e290;        /* false for 0, NaN, infinity too */
    sum = 0.0;
#pragma unroll 2
    for (int k = 0; k < NB; k++) {
        double v[8], x[8];
        wb_llr_stats_load8<FRAMING>(row, k, scramble, v);
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = wb_quot(v[j], mean, rcp, fast) - (double)((v[j] > 0.0) - (v[j] < 0.0));
#pragma unroll
        for (int j = 0; j < 8; j++) { sum += x[j]; sumsq += x[j] * x[j]; }
    }
    {
        double v[8], x[8];
        wb_llr_stats_load8<FRAMING>(row, NB, scramble, v);
#pragma unroll
        for (int j = 0; j < TAIL; j++) x[j] = wb_quot(v[j], mean, rcp, fast) - (double)((v[j] > 0.0) - (v[j] < 0.0));
#pragma unroll
        for (int j = 0; j < TAIL; j++) { sum += x[j]; sumsq += x[j] * x[j]; }
    }
    const double estvar = ((double)n * sumsq - sum * sum) / (double)(n * (n - 1));
    return 4.0 * wb_esn0_from_var(estvar);
}

__global__ void __launch_bounds__(128)
wb_llr_stats_kernel(const float *sd, unsigned long long sd_stride, const unsigned *jobs, double *c4,
                    const wb_cursor *cur, int job_cap, int n_streams, int framing, const uint8_t *scramble)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int s = g / job_cap, k = g - s * job_cap;
    if (s >= n_streams || k >= (int)cur[s].n_jobs) return;
    const float *row = sd + (size_t)s * sd_stride + jobs[(size_t)s * job_cap + k];
    c4[(size_t)s * job_cap + k] = (framing == WB_FRAMING_V1) ? wb_llr_stats_row<WB_FRAMING_V1>(row, scramble)
                                                             : wb_llr_stats_row<WB_FRAMING_V2>(row, scramble);
}

/* standalone sd_to_llr over n blocks of 2580 soft decisions (wb_sd_to_llr_batch): statistics */
__global__ void __launch_bounds__(128)
wb_llr_stats_plain_kernel(const float *sd, double *c4, long long n_blocks)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_blocks) return;
    /* the last block's tail load would read 4 floats past the array: the plain form takes it element by element */
    const float *row = sd + (size_t)g * WB_NCODE;
    const int n = WB_NCODE;
    double sum = 0.0, sumsq = 0.0, mean, estvar;
    for (int c = 0; c < n; c++) sum += fabs((double)row[c]);
    mean = sum / (double)n;
    sum = 0.0;
    for (int c = 0; c < n; c++) {
        double v = (double)row[c];
        double sign = (double)((v > 0.0) - (v < 0.0));
        double x = v / mean - sign;
        sum += x;
        sumsq += x * x;
    }
    estvar = ((double)n * sumsq - sum * sum) / (double)(n * (n - 1));
    c4[g] = 4.0 * wb_esn0_from_var(estvar);
}

/* ... and the scaling, reference src/mpdecode_core.c:594-595 */
__global__ void __launch_bounds__(256)
wb_llr_scale_kernel(const float *sd, const double *c4, float *llr, long long n_blocks)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_blocks * WB_NCODE) return;
    llr[g] = wb_llr_scale_fast(c4[g / WB_NCODE], sd[g]);
}

/* ---- K4 ------------------------------------------------------------------ */

struct wb_ldpc_smem {
    float msg[WB_LDPC_NMSG];            /* 28 896 B */
    wb_phi0_flag lut;                   /*  5 224 B */
    unsigned ballot[(WB_NCODE + 31) / 32 + 1];   /* hard decisions by slot, bit l of word w = slot 32w + l */
    unsigned nat[(WB_NCODE + 31) / 32 + 1];      /* ... and by variable, after the decode */
    unsigned crc_part[4];
    unsigned nsat[2];                   /* satisfied checks of this / the next iteration (counted with one atomic per warp) */
};

/* phi0 (reference src/phi0.c:13-218): one 4-byte load; the few buckets with a breakpoint inside carry a flag and
   take two more loads and the compare (flag form of wb_phi0.h).  The decoder is shared-memory-bandwidth bound and
   these look-ups were two thirds of its traffic as 16 bytes each. */
__device__ __forceinline__ float wb_phi0_s(const wb_phi0_flag &lut, float x)
{
    int b = ((int)__float_as_uint(x) >> 17) - (WB_PHI0_EXP0 << 6);
    b = max(0, min(b, WB_PHI0_NBUCKET));
    float v = lut.val[b];
    if ((int)__float_as_uint(v) < 0) {
        const int g = b >> 4;
        v = (x < lut.gthr[g]) ? __uint_as_float(__float_as_uint(v) & 0x7fffffffu) : lut.gvhi[g];
    }
    return v;
}

__device__ __forceinline__ float wb_signed(float mag, unsigned neg)
{
    return __uint_as_float(__float_as_uint(mag) | (neg << 31));   /* mag >= +0 */
}

__global__ void __launch_bounds__(WB_LDPC_THREADS, 5)
wb_ldpc_kernel(wb_ldpc_args a, long long n_direct)
{
    extern __shared__ __align__(16) unsigned char wb_ldpc_raw[];
    wb_ldpc_smem &sm = *reinterpret_cast<wb_ldpc_smem *>(wb_ldpc_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool direct = a.llr_in != nullptr;
    int s = 0, k = 0;
    size_t slot;
    if (direct) {
        slot = blockIdx.x;
        if ((long long)slot >= n_direct) return;
    } else {
        s = blockIdx.x / a.job_cap;
        k = blockIdx.x - s * a.job_cap;
        if (k >= (int)a.cur[s].n_jobs) return;
        slot = (size_t)s * a.job_cap + k;
    }

    /* phi0 table -> shared memory */
    {
        const unsigned *src = reinterpret_cast<const unsigned *>(a.lut);
        unsigned *dst = reinterpret_cast<unsigned *>(&sm.lut);
        for (int i = tid; i < (int)(sizeof(wb_phi0_flag) / 4); i += WB_LDPC_THREADS) dst[i] = src[i];
    }
    /* LLRs: gather + scale (mode A) or load (mode B), in natural order (coalesced), staged through the message array
       (not in use yet); then thread tid takes the LLRs of its slots tid + 288 r, r = 0..8, which it owns for the whole
       decode: registers */
    float llr[9];
    if (direct) {
        const float *src = a.llr_in + slot * WB_NCODE;
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const int c = tid + r * WB_LDPC_THREADS;
            if (c < WB_NCODE) sm.msg[c] = src[c];
        }
    } else {
        const float *row = a.sd + (size_t)s * a.sd_stride + a.jobs[slot];
        const double c4 = a.c4[slot];
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const int c = tid + r * WB_LDPC_THREADS;
            if (c < WB_NCODE) {
                float v = row[wb_cw_symbol(a.framing, c)];
                if (a.framing == WB_FRAMING_V2 && a.scramble[c % WB_SCRAMBLE_LEN]) v = -v;
                const float l = wb_llr_scale_fast(c4, v);
                sm.msg[c] = l;
                if (a.llr_out) a.llr_out[slot * WB_NCODE + c] = l;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int i = tid + r * WB_LDPC_THREADS;
        const int v = (i < WB_NDATA) ? (int)a.vedge[i].w : i;         /* parity columns stay in place */
        llr[r] = (i < WB_NCODE) ? sm.msg[v] : 0.0f;
    }
    __syncthreads();

    /* initial v->c messages, reference src/mpdecode_core.c:343-350 */
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int i = tid + r * WB_LDPC_THREADS;
        if (i < WB_NCODE) {
            const float l = llr[r];
            const float q = wb_signed(wb_phi0_s(sm.lut, fabsf(l)), (l < 0.0f) ? 1u : 0u);
            if (i < WB_NDATA) {
                const ushort4 e = a.vedge[i];
                sm.msg[e.x] = q; sm.msg[e.y] = q; sm.msg[e.z] = q;
            } else {
                const int p = i - WB_NDATA;
                sm.msg[13 * WB_NPAR + p] = q;
                if (p < WB_NPAR - 1) sm.msg[12 * WB_NPAR + p + 1] = q;
            }
        }
    }
    if (tid < 2) sm.nsat[tid] = 0;
    __syncthreads();

    int result = a.max_iter, pcc = -1;
    for (int iter = 0; iter < a.max_iter; iter++) {
        /* ---- check-node pass, reference src/mpdecode_core.c:412-436 ---- */
        /* the two rounds touch disjoint checks: no barrier between them; the satisfied checks are counted with one
           shared-memory atomic per warp and read after the variable pass (one CTA barrier less per iteration) */
        int nsat_mine = 0;
#pragma unroll 1
        for (int rnd = 0; rnd < 2; rnd++) {
            int j = tid + rnd * WB_LDPC_THREADS;
            int sat = 0;
            if (j < WB_NPAR) {
                float q[WB_LDPC_SLOTS];
#pragma unroll
                for (int t = 0; t < WB_LDPC_SLOTS; t++) q[t] = sm.msg[t * WB_NPAR + j];
                if (j == 0) q[12] = 0.0f;                    /* check 0 has no "parity j-1" edge */
                float phi_sum = fabsf(q[0]);
                unsigned sign = __float_as_uint(q[0]) >> 31;
#pragma unroll
                for (int t = 1; t < WB_LDPC_SLOTS; t++) {
                    if (t == 12 && j == 0) continue;
                    phi_sum = phi_sum + fabsf(q[t]);
                    sign ^= __float_as_uint(q[t]) >> 31;
                }
                sat = (sign == 0);
#pragma unroll
                for (int t = 0; t < WB_LDPC_SLOTS; t++) {
                    if (t == 12 && j == 0) continue;
                    float v = wb_phi0_s(sm.lut, phi_sum - fabsf(q[t]));
                    sm.msg[t * WB_NPAR + j] = wb_signed(v, sign ^ (__float_as_uint(q[t]) >> 31));
                }
            }
            nsat_mine += sat;
        }
        {
            const unsigned wsum = __reduce_add_sync(0xffffffffu, (unsigned)nsat_mine);
            if (lane == 0 && wsum) atomicAdd(&sm.nsat[iter & 1], wsum);
        }
        __syncthreads();
        if (tid == 0) sm.nsat[(iter + 1) & 1] = 0;      /* everyone read it before the barrier above; next used after the next one */
        /* ---- variable-node pass, reference src/mpdecode_core.c:439-464 ---- */
        int nz = 0;
#pragma unroll
        for (int rnd = 0; rnd < 9; rnd++) {
            const int i = tid + rnd * WB_LDPC_THREADS;
            bool bit = false;
            if (i < WB_NDATA) {
                ushort4 e = a.vedge[i];
                float r0 = sm.msg[e.x], r1 = sm.msg[e.y], r2 = sm.msg[e.z];
                float Qi = llr[rnd];
                Qi = Qi + r0; Qi = Qi + r1; Qi = Qi + r2;
                bit = Qi < 0.0f;
                float t0 = Qi - r0, t1 = Qi - r1, t2 = Qi - r2;
                sm.msg[e.x] = wb_signed(wb_phi0_s(sm.lut, fabsf(t0)), (t0 > 0.0f) ? 0u : 1u);
                sm.msg[e.y] = wb_signed(wb_phi0_s(sm.lut, fabsf(t1)), (t1 > 0.0f) ? 0u : 1u);
                sm.msg[e.z] = wb_signed(wb_phi0_s(sm.lut, fabsf(t2)), (t2 > 0.0f) ? 0u : 1u);
                nz |= bit;
            } else if (i < WB_NCODE) {
                int p = i - WB_NDATA;
                int e0 = 13 * WB_NPAR + p, e1 = 12 * WB_NPAR + p + 1;
                bool two = p < WB_NPAR - 1;
                float r0 = sm.msg[e0], r1 = two ? sm.msg[e1] : 0.0f;
                float Qi = llr[rnd];
                Qi = Qi + r0;
                if (two) Qi = Qi + r1;
                bit = Qi < 0.0f;
                float t0 = Qi - r0;
                sm.msg[e0] = wb_signed(wb_phi0_s(sm.lut, fabsf(t0)), (t0 > 0.0f) ? 0u : 1u);
                if (two) {
                    float t1 = Qi - r1;
                    sm.msg[e1] = wb_signed(wb_phi0_s(sm.lut, fabsf(t1)), (t1 > 0.0f) ? 0u : 1u);
                }
            }
            unsigned b = __ballot_sync(0xffffffffu, bit);
            int w = (rnd * WB_LDPC_THREADS >> 5) + warp;
            if (lane == 0 && w * 32 < WB_NCODE + 31) sm.ballot[w] = b;
        }
        nz = __syncthreads_or(nz);
        const int nsat = (int)sm.nsat[iter & 1];
        /* exits, reference src/mpdecode_core.c:467-483 (data[] is all zero in run_ldpc_decoder) */
        if (!nz) { result = iter + 1; break; }
        pcc = nsat;
        if (nsat == WB_NPAR) { result = iter + 1; break; }
    }
    if (a.max_iter <= 0) {
        for (int w = tid; w < (WB_NCODE + 31) / 32; w += WB_LDPC_THREADS) sm.ballot[w] = 0;
        __syncthreads();
    }

    /* ---- hard decisions back into variable order ---- */
    for (int w = tid; w < (WB_NCODE + 31) / 32 + 1; w += WB_LDPC_THREADS) sm.nat[w] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int i = tid + r * WB_LDPC_THREADS;
        if (i < WB_NCODE && ((sm.ballot[i >> 5] >> (i & 31)) & 1u)) {
            const int v = (i < WB_NDATA) ? (int)a.vedge[i].w : i;
            atomicOr(&sm.nat[v >> 5], 1u << (v & 31));
        }
    }
    __syncthreads();

    /* ---- output ---- */
    if (direct) {
        uint8_t *out = a.bits_out + slot * 323;
        for (int B = tid; B < 323; B += WB_LDPC_THREADS)
            out[B] = (uint8_t)(__brev((sm.nat[B >> 2] >> (8 * (B & 3))) & 0xffu) >> 24);
        if (tid == 0) { a.iters_out[slot] = result; a.pcc_out[slot] = pcc; }
        return;
    }
    /* CRC16-CCITT-FALSE of the 2048 payload bits as an XOR of per-bit contributions
       (the CRC is affine over GF(2)); same value as reference src/drs232_ldpc.c:91-102 */
    unsigned acc = 0;
    if (tid < 64) {
        unsigned w = sm.nat[tid];
        while (w) {
            int l = __ffs(w) - 1;
            w &= w - 1;
            acc ^= a.crc_tab[tid * 32 + l];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sm.crc_part[warp] = acc;
    }
    __syncthreads();
    wb_codeword *cw = a.cw + slot;
    for (int B = tid; B < 258; B += WB_LDPC_THREADS)
        cw->bytes[B] = (uint8_t)(__brev((sm.nat[B >> 2] >> (8 * (B & 3))) & 0xffu) >> 24);
    if (tid == 0) {
        unsigned crc = a.crc0 ^ sm.crc_part[0] ^ sm.crc_part[1];
        unsigned w64 = sm.nat[64];     /* variables 2048..2079: bytes 256, 257 are its low 16 bits */
        unsigned b256 = __brev(w64 & 0xffu) >> 24, b257 = __brev((w64 >> 8) & 0xffu) >> 24;
        unsigned tx = b256 + (b257 << 8);
        int ok = (crc == tx);
        cw->stream = s;
        cw->seq = a.cur[s].seq0 + (unsigned)k;
        cw->iters = result;
        cw->parity_ok = pcc;
        cw->crc_ok = ok;
        cw->pad[0] = cw->pad[1] = 0;
        atomicAdd(&a.st[s].packets, 1u);
        if (!ok) atomicAdd(&a.st[s].packet_errors, 1u);
    }
}

#endif /* WB_LDPC_KERNEL_CUH */
