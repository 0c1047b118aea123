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
 *   A  (stream warps)  the frame arrives as ONE TMA bulk copy per stream (cf32: cp.async.bulk + an mbarrier
 *                      counting the bytes, issued by lane 0 at the end of the previous frame; the integer
 *                      formats are fetched and converted by the warp; every frame is also prefetched into
 *                      L2 one frame ahead, so HBM latency is off the critical path and no registers are
 *                      tied up); window + 256-point FFT in the
 *                      reference's butterfly order (leaf level fused with the window, top level with
 *                      |X|^2 and the IIR; bank-conflict-free swizzled work buffer), spectrum IIR in
 *                      registers, warp-argmax peak picking.
 *   B1 (warps 0..W-1, lane = (tone, stream))  ONLY the sequential part of the mixer: oscillator
 *                      recurrence + down-mix product, written in place over the samples.  The
 *                      recurrence does not depend on the samples, so the W warps all run the same
 *                      chains and each mixes only its own segment of the frame (see the phase).
 *   B2 (stream warps, lane = block of Ts integrator outputs)  Ts-tap sums in the reference's
 *                      ring-buffer slot order, in place; e_i = |.|^2 summed over tones -> E.
 *   B3 (warp 0 = real, warp 1 = imaginary part, lane = stream)  the sequential fine-timing
 *                      accumulation t_c = sum e_i phi_ft[i]: one dependent addition per output,
 *                      operands in loop-carried registers, the (periodic) multipliers in registers
 *                      (wb_b3_chain).
 *   C  (stream warps)  every lane recomputes the frame's scalars from t_c (atan2 timing estimate,
 *                      ppm, next nin, resampling offsets); lanes = symbols: linear-interpolated
 *                      resampling and soft decisions, 48 (96) floats per frame written coalesced;
 *                      next frame's fetch.
 *
 * Four CTA barriers per frame: A|B1, B1|B2, B2|B3, B3|C.  C and the next frame's A are work of the stream's own
 * warp on the stream's own memory, so nothing separates them; the vote "any stream left?" rides on A|B1.
 *
 * The sequential phases cost a few warps' issue slots for ALL streams of the CTA (every lane carries
 * a different dependent chain); while one CTA of an SM is in B1/B3 the other one runs A/B2/C.
 *
 * Shared memory per stream: X[nst + nmax] float2 (the nst = 2Ts + Ts/2 old samples the mixer can reach
 * back to + the new ones -> tone 0 mixer products -> tone 0 integrator outputs, all in place),
 * Y[(M-1) * ylen] float2 for the other tones (the FFT work buffer in phase A) and E, the frame's
 * fine-timing terms e_i.  The L1 that is left beside 228 KB of shared memory is flushed by the sample
 * stream every frame: a global or local (spill) load is an L2 round trip and an indexed constant-bank
 * load issues only every ~10 cycles, so the frame geometry is compile-time (BLK) and tables live in
 * shared memory, registers or immediates wherever a sequential phase needs them.
 * Stream regions are an odd multiple of 8 bytes mod 128 apart so the lanes of warp 0 (one stream
 * each) hit distinct banks.  For Ts = 8 the mixer products / integrator outputs are stored with
 * their low three index bits XORed with bits 4..6 (wb_phys) so that B2's lanes, which walk blocks
 * of eight, spread over all banks.  Per CTA: the frame scalars and the FFT twiddles (1.5 KB).
 * HBM traffic: every input sample is read once (8 B as cf32), 4 B x Nbits/N written.
 */
#ifndef WB_FSK_KERNEL_CUH
#define WB_FSK_KERNEL_CUH

#include "wb_internal.h"
#include "wb_math.h"

#define WB_NST 1                     /* stash registers per lane: 2*Ts + Ts/2 <= 25 samples are carried over */
#define WB_NEQ (WB_MAX_NDFT / 2 / 32)

struct wb_fsk_sc {                 /* per-stream frame scalars in shared memory (80 bytes) */
    float2 phi_c[WB_MAXM];
    short pb[WB_MAXM];             /* estimator bins in force before this frame (fsk->f_est) */
    short nb[WB_MAXM];             /* bins estimated from this frame */
    short nin, pad0;
    int flags;                     /* bit 0 = active this frame */
    float tcr, tci;                /* fine-timing sum t_c of this frame (B3 -> C) */
    float norm, ppm, rx_timing;
    int pad1;
};
static_assert(sizeof(wb_fsk_sc) == 80, "shared-memory budget of wb_fsk_kernel");

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
    unsigned char *hard;           /* WB_FLAG_HARD_BITS: [n_streams][WB_HARD_PRE + sd_cap] one byte per bit (rx_bits of fsk_demod()), else NULL */
};

__device__ __forceinline__ float2 wb_cmul2(float2 a, float2 b)   /* reference src/comp_prim.h cmult */
{
    float2 c;
    c.x = __fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    c.y = __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
    return c;
}

/* The same products with the two multiplications of each half in one packed instruction (mul.f32x2: two IEEE
   round-to-nearest products; FMUL2 swaps the halves of an operand for free), the additions scalar: four instructions
   instead of six on the oscillator's dependent chain, every float the same.  Under load a chain step costs its issue
   slots more than its arithmetic latency, so fewer instructions is a shorter chain. */
__device__ __forceinline__ unsigned long long wb_pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void wb_mul2(unsigned long long a, unsigned long long b, float &lo, float &hi)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
/* (a.x + b.x, a.y + b.y) as one packed addition: the same two IEEE sums */
__device__ __forceinline__ float2 wb_add2p(float2 a, float2 b)
{
    unsigned long long r;
    float2 c;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(wb_pack2(a.x, a.y)), "l"(wb_pack2(b.x, b.y)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r));
    return c;
}
__device__ __forceinline__ float2 wb_sub2p(float2 a, float2 b)
{
    unsigned long long r;
    float2 c;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(wb_pack2(a.x, a.y)), "l"(wb_pack2(b.x, b.y)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r));
    return c;
}
/* a * b, reference src/comp_prim.h cmult: (a.x b.x - a.y b.y, a.x b.y + a.y b.x) */
__device__ __forceinline__ float2 wb_cmul2p(float2 a, float2 b)
{
    const unsigned long long pa = wb_pack2(a.x, a.y);
    float p0, p1, q0, q1;
    wb_mul2(pa, wb_pack2(b.x, b.y), p0, p1);          /* a.x b.x, a.y b.y */
    wb_mul2(pa, wb_pack2(b.y, b.x), q0, q1);          /* a.x b.y, a.y b.x */
    return make_float2(__fsub_rn(p0, p1), __fadd_rn(q0, q1));
}
/* x * conj(ph), reference src/fsk.c:794-797 cmult(sample, cconj(phi)): (x.x ph.x + x.y ph.y, x.y ph.x - x.x ph.y) */
__device__ __forceinline__ float2 wb_mixp(float2 x, float2 ph)
{
    const unsigned long long px = wb_pack2(x.x, x.y);
    float p0, p1, q0, q1;
    wb_mul2(px, wb_pack2(ph.x, ph.y), p0, p1);        /* x.x ph.x, x.y ph.y */
    wb_mul2(px, wb_pack2(ph.y, ph.x), q0, q1);        /* x.x ph.y, x.y ph.x */
    return make_float2(__fadd_rn(p0, p1), __fsub_rn(q1, q0));
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

/* input samples are touched exactly once: keep them out of L1 so the small constant tables stay resident */
template <bool CF32>
__device__ __forceinline__ void wb_load_raw(int fmt, const unsigned char *p, unsigned long long idx, unsigned &lo, unsigned &hi)
{
    if (CF32) {
        asm("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(reinterpret_cast<const uint2 *>(p) + idx));
    } else if (fmt == WB_FMT_CS16) {
        asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(lo) : "l"(reinterpret_cast<const unsigned *>(p) + idx));
        hi = 0u;
    } else {
        unsigned short t;
        asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(t) : "l"(reinterpret_cast<const unsigned short *>(p) + idx));
        lo = t; hi = 0u;
    }
}

/* product with a unit twiddle tw[0] = (cosf(-0.0), sinf(-0.0)) = (1, -0): the two multiplications by 1 are
   exact for every input and dropped; the two by -0 are kept so NaN/Inf samples propagate as in kiss_fft */
__device__ __forceinline__ float2 wb_ucmul(float2 a)
{
    float2 c;
    c.x = __fsub_rn(a.x, __fmul_rn(a.y, -0.0f));
    c.y = __fadd_rn(__fmul_rn(a.x, -0.0f), a.y);
    return c;
}

/* cf32 frames go global -> shared memory as ONE TMA bulk copy per stream and frame (cp.async.bulk + an mbarrier that
   counts the bytes): no registers, no conversion, one instruction where 13 cp.async per lane used to be.  Source and
   destination must be 16-byte aligned and the size a multiple of 16, while a frame starts at any sample (8 bytes): the
   copy starts one sample early when the row position is odd and the landing place in X is skewed by `dlt` in {0, 1}
   samples, chosen per frame, so that both ends are aligned (see the kernel).  The .shared::cluster destination form is
   used with the CTA's own shared-memory address (a CTA is its own cluster of one). */
__device__ __forceinline__ unsigned wb_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wb_mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wb_tma_fetch(void *smem_dst, const void *gmem_src, unsigned bytes, void *bar)
{
    /* what the lanes of this warp read from / wrote to the landing zone (generic proxy) comes before the copy's writes */
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wb_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(wb_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(wb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wb_mbar_wait(void *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(wb_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

/* Tensor memory as lane-private scratch: this kernel issues no MMA, so its 256 KB per SM sit idle, while the register
   file is the scarcest resource here (72 registers per thread, and what spills goes to L2: the L1 is flushed by the
   sample stream).  A warp may touch the 32 TMEM lanes of its quadrant (warp % 4); 32x32b accesses give every thread
   one 32-bit cell per column, so a warp's private columns work like 8 extra registers per thread with a ~12-cycle
   access.  The spectrum estimate (used only in phase A) and the carried-over samples (A -> C) wait there. */
__device__ __forceinline__ void wb_tmem_st8(unsigned taddr, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 "tcgen05.wait::st.sync.aligned;"
                 :: "r"(taddr), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7) : "memory");
}
__device__ __forceinline__ void wb_tmem_ld8(unsigned taddr, float &a0, float &a1, float &a2, float &a3, float &a4, float &a5, float &a6, float &a7)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7) : "r"(taddr) : "memory");
}

/* FFT work buffer index: XOR bits 4..5 into both bit pairs below them.  The leaf stores (lanes differ in bits
   2..5), the radix-4 level with m = 4 (lanes differ in bits 0..1 and 4..5) and the wider levels (lanes differ
   in bits 0..3) all become bank-conflict free. */
__device__ __forceinline__ int wb_fidx(int i)
{
    return i ^ (((i >> 4) & 3) * 5);
}

/* small constant tables in global memory (fine-timing oscillator, Hann window, leaf order): keep their lines in L1
   against the streaming local-memory / sample traffic */
__device__ __forceinline__ float2 wb_ldg_keep2(const float2 *p)
{
    float2 v;
    asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 wb_ldg_keep4(const float4 *p)
{
    float4 v;
    asm("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float wb_ldg_keep(const float *p)
{
    float v;
    asm("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <bool SWZ>
__device__ __forceinline__ int wb_phys(int n)
{
    /* bank swizzle of the product / integrator arrays: a permutation inside each aligned block of 8.
       Block 48 and the partial block 49 (Nsym = 48) have key 0, so nothing is mapped past the end. */
    return SWZ ? ((n & ~7) | ((n ^ (n >> 4)) & 7)) : n;
}

/* The fine-timing sum t_c += e_i * phi_ft[i] (reference src/fsk.c:870-873) for one component, one chain per lane:
   one dependent addition per output.  Hardly anything else runs on the warp's SM partition during this phase, so
   nothing hides a load; left alone, ptxas issues each load a few instructions before its use and the chain waits
   ~30 cycles per shared-memory pair (and an indexed constant-bank load issues only every ~10 cycles).  So:
   * the e pairs travel in loop-carried registers: a rolled loop whose body adds BB blocks and refills each register
     right after its last use with the pair BB blocks further on -- a load can sink no further than the loop's back
     edge, one body ahead of its use;
   * the multipliers need no loads at all in the steady state: the reference builds phi_ft by the float recurrence
     phi_ft *= e^{j 2 pi / P}, which after a short transient falls into an exactly periodic orbit of period P = one
     block (the host checks this on the table it built with the reference's arithmetic: p.pft_steady), so from block
     WB_PFT_NT on every block uses the same Ts multipliers: registers.
   q_ = e in output order (shared memory), ct = that component of phi_ft (kernel-parameter constant bank). */
template <int TS, int NBLK_>
__device__ __forceinline__ float wb_b3_chain(const float2 *q_, const float *ct)
{
    constexpr int HP = TS / 2;                  /* pairs per block */
    constexpr int BB = 3;                       /* blocks per loop body */
    constexpr int NP = BB * HP;                 /* pairs per body */
    constexpr int NT = WB_PFT_NT;               /* transient blocks */
    constexpr int NB = (NBLK_ - NT) / BB, LEFT = (NBLK_ - NT) % BB;
    static_assert(NB >= 2, "at least two bodies");
    float tacc = 0.0f;
    float cs[TS];
    float2 ev[NP];
#pragma unroll
    for (int t = 0; t < TS; t++) cs[t] = ct[NT * TS + t];
#pragma unroll
    for (int j = 0; j < NP; j++) ev[j] = q_[NT * HP + j];             /* the first body's pairs */
    /* transient blocks: multipliers straight from the table */
#pragma unroll
    for (int j = 0; j < NT * HP; j++) {
        const float2 e2 = q_[j];
        tacc = __fadd_rn(tacc, __fmul_rn(e2.x, ct[2 * j]));
        tacc = __fadd_rn(tacc, __fmul_rn(e2.y, ct[2 * j + 1]));
    }
    const float2 *qn = q_ + NT * HP + NP;       /* the next body's pairs */
    /* one body; REFILL pairs are reloaded for the next one */
#define WB_B3_BODY(REFILL)                                                                              \
    do {                                                                                                \
        _Pragma("unroll")                                                                               \
        for (int j = 0; j < NP; j++) {                                                                  \
            tacc = __fadd_rn(tacc, __fmul_rn(ev[j].x, cs[(2 * j) % TS]));                               \
            tacc = __fadd_rn(tacc, __fmul_rn(ev[j].y, cs[(2 * j + 1) % TS]));                           \
            if (j < (REFILL)) ev[j] = qn[j];                                                            \
        }                                                                                               \
        qn += NP;                                                                                       \
    } while (0)
#pragma unroll 1
    for (int k = 0; k < NB - 1; k++) WB_B3_BODY(NP);
    WB_B3_BODY(LEFT * HP);
#undef WB_B3_BODY
#pragma unroll
    for (int j = 0; j < LEFT * HP; j++) {       /* the left-over blocks */
        tacc = __fadd_rn(tacc, __fmul_rn(ev[j].x, cs[(2 * j) % TS]));
        tacc = __fadd_rn(tacc, __fmul_rn(ev[j].y, cs[(2 * j + 1) % TS]));
    }
    return tacc;
}

/* the same without the periodicity assumption: every multiplier from the constant bank */
template <int TS, int NBLK_>
__device__ __forceinline__ float wb_b3_chain_generic(const float2 *q_, const float *ct)
{
    float tacc = 0.0f;
#pragma unroll 1
    for (int b = 0; b < NBLK_; b++) {
#pragma unroll
        for (int j = 0; j < TS / 2; j++) {
            const float2 e2 = q_[b * (TS / 2) + j];
            tacc = __fadd_rn(tacc, __fmul_rn(e2.x, ct[b * TS + 2 * j]));
            tacc = __fadd_rn(tacc, __fmul_rn(e2.y, ct[b * TS + 2 * j + 1]));
        }
    }
    return tacc;
}

#ifdef WB_PHASE_CLK
/* debug build only (make dbg): cycles per phase summed over all CTAs and frames, read with wb_debug_phase_clk */
__device__ unsigned long long wb_phase_clk[12];
#define WB_CLK(I) do { if (tid == 0) { const unsigned now_ = (unsigned)clock(); atomicAdd(&wb_phase_clk[I], (unsigned long long)(now_ - clk_prev)); clk_prev = now_; } } while (0)
#else
#define WB_CLK(I) do { } while (0)
#endif

/* BLK: P == Ts (step 1), the configuration every Wenet script uses: the whole frame geometry is a compile-time
   constant (host and kernel share the wb_blk_* formulas of wb_internal.h).  !BLK: general P, geometry from p. */
template <int M, int TS, bool CF32, bool BLK, bool HARD>
__global__ void __launch_bounds__(M == 2 ? 448 : 256, 2)
wb_fsk_kernel(wb_fsk_params p, wb_fsk_args a)
{
    /* frame geometry, reference src/fsk.c:128-259 with nsyms = 48 */
    constexpr int NSYM = WB_FRAME_SYMS, N_ = TS * NSYM, NMEM = N_ + 2 * TS, NST = 2 * TS + TS / 2, NMAX = N_ + TS / 2;
    constexpr int NBITS = (M == 2) ? NSYM : 2 * NSYM;
    constexpr int NDFT = WB_MAX_NDFT, NH = NDFT / 2;
    static_assert(N_ >= 256 && N_ < 512 && NDFT == 256, "estimator FFT = highest set bit of N = 256 (src/fsk.c:169-173)");
    constexpr int XLEN = wb_geom_xlen(TS);
    const int Pc = BLK ? TS : p.P, stepc = BLK ? 1 : p.step, nintc = BLK ? (NSYM + 1) * TS : p.nint;
    const int ylen = BLK ? wb_blk_ylen(TS) : p.ylen, blen = BLK ? wb_blk_blen(M, TS) : p.blen;
    const int sreg = BLK ? wb_blk_sreg(M, TS) : p.sreg;
    constexpr int NPRE = (TS * WB_FRAME_SYMS + TS / 2 + 31) / 32;     /* lanes x NPRE >= nmax */
    constexpr bool SWZ = (TS == 8);
    constexpr int NBLK = WB_FRAME_SYMS + 1;                           /* integrator outputs come in 49 blocks of P */
    extern __shared__ __align__(16) unsigned char wb_fsk_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int spb = a.spb;
    const int sg = blockIdx.x * spb + warp;
    const bool have = sg < a.n_streams;
    constexpr int Ndft = NDFT, nh = NH, nst = NST;
    const int fmt = p.in_fmt;

    wb_fsk_sc *sc = reinterpret_cast<wb_fsk_sc *>(wb_fsk_raw);
    unsigned long long *MB = reinterpret_cast<unsigned long long *>(wb_fsk_raw + sizeof(wb_fsk_sc) * spb);   /* one mbarrier per stream */
    float2 *TW = reinterpret_cast<float2 *>(wb_fsk_raw + wb_geom_head(spb, (int)sizeof(wb_fsk_sc)) - 3 * (Ndft >> 2) * 8);
    unsigned char *regions = reinterpret_cast<unsigned char *>(TW + 3 * (Ndft >> 2));
    float2 *X = reinterpret_cast<float2 *>(regions + (size_t)warp * sreg);
    float2 *Y = X + XLEN;
    float *E = reinterpret_cast<float *>(Y + blen);

    /* ---- per-stream state -> registers / shared memory ---- */
    wb_stream_state *st = have ? a.state + sg : nullptr;
    /* row positions in samples fit 32 bits (wb_create checks); the row pointers are recomputed from the stream
       index where they are needed instead of being carried through the frame loop (register pressure) */
#define WB_ROW_IN(SG) (a.in + (size_t)(SG) * a.in_stride)
#define WB_ROW_SD(SG) (a.sd + (size_t)(SG) * a.sd_stride + WB_CARRY_CAP)
    const int sgc = have ? sg : 0;
    unsigned pos = 0, fill = 0;
    int nin = N_;
    unsigned n_out = 0;
    float est[WB_NEQ];
    float2 stash = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int q = 0; q < WB_NEQ; q++) est[q] = 0.0f;
    /* (-DWB_NO_TMEM, the `san` build of the Makefile, keeps these values in registers instead: compute-sanitizer's
       synccheck mistakes tcgen05.alloc for an uninitialised mbarrier and stops the kernel) */
    if (CF32 && lane == 0) {
        wb_mbar_init(&MB[warp], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#ifdef WB_NO_TMEM
    __syncthreads();
#endif
#ifndef WB_NO_TMEM
    /* 32 TMEM columns per CTA: 8 per warp, warps of one quadrant side by side */
    static_assert(WB_NEQ == 4, "est[4] + stash fill one 8-column TMEM row");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;\n"
                     "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                     :: "r"((unsigned)__cvta_generic_to_shared(&sc[0].pad1)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = (unsigned)sc[0].pad1;
    const unsigned taddr = tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + 8u * (unsigned)(warp >> 2);
#endif
    for (int i = tid; i < 3 * (Ndft >> 2); i += blockDim.x) TW[i] = __ldg(&p.tw[i]);
    if (have) {
        pos = (unsigned)st->in_pos; fill = (unsigned)st->in_fill; nin = st->nin;
#pragma unroll
        for (int q = 0; q < WB_NEQ; q++)
            if (lane + 32 * q < nh) est[q] = st->fft_est[lane + 32 * q];
        if (lane < nst) stash = st->samp_old[lane];
        if (lane == 0) {
            wb_fsk_sc &c = sc[warp];
            for (int m = 0; m < M; m++) { c.phi_c[m] = st->phi_c[m]; c.pb[m] = (short)st->fbin[m]; c.nb[m] = 0; }
            c.norm = st->norm_rx_timing; c.ppm = st->ppm; c.rx_timing = st->rx_timing;
            c.nin = (short)nin; c.flags = 0; c.tcr = c.tci = 0.0f;
        }
    } else if (lane == 0) {
        wb_fsk_sc &c = sc[warp];
        for (int m = 0; m < M; m++) { c.phi_c[m] = make_float2(1.0f, 0.0f); c.pb[m] = 0; c.nb[m] = 0; }
        c.nin = (short)N_; c.flags = 0; c.norm = 0.0f; c.ppm = 0.0f; c.rx_timing = 0.0f; c.tcr = c.tci = 0.0f;
    }
    /* cf32: start the copy of the NIN-sample frame at row position POS; its first sample lands at X[nst + dlt], dlt =
       (warp + nst + POS) & 1 (regions of odd warps start 8 bytes off a 16-byte boundary).  When POS is odd the copy
       starts one sample early -- that sample falls on the newest old sample's slot, which phase A rewrites after the
       landing -- and the byte count is rounded up to 16: at most X[XLEN + 1] = Y[1] is touched, which is why the FFT
       work buffer starts at Y + 2. */
#define WB_FETCH_CF32(POS, NIN)                                                                         \
    do {                                                                                                \
        if (lane == 0) {                                                                                \
            int sgv_ = sgc;                                                                             \
            asm volatile("" : "+r"(sgv_));                                                              \
            const unsigned odd_ = (POS) & 1u;                                                           \
            const float2 *src_ = reinterpret_cast<const float2 *>(WB_ROW_IN(sgv_)) + ((POS) - odd_);    \
            float2 *dst_ = X + nst + ((warp + nst + (int)(POS)) & 1) - (int)odd_;                       \
            wb_tma_fetch(dst_, src_, (((unsigned)(NIN) + odd_ + 1u) & ~1u) * 8u, &MB[warp]);            \
        }                                                                                               \
    } while (0)
#ifndef WB_NO_TMEM
    wb_tmem_st8(taddr, est[0], est[1], est[2], est[3], stash.x, stash.y, 0.0f, 0.0f);
#endif
    /* a frame is fetched if and only if it will be demodulated: the condition is the loop's `active` one pass ahead, so
       every copy is waited for before the kernel ends and every wait has its copy */
    if (CF32 && have && (pos + (unsigned)nin <= fill) && ((unsigned)NBITS <= a.sd_cap)) WB_FETCH_CF32(pos, nin);

    const float omt = __fsub_rn(1.0f, p.tc);
    constexpr bool blocked = BLK;          /* P == Ts: the configuration every Wenet script uses */

    __syncthreads();                       /* the twiddle table above is filled by all threads and read by every warp's first A */
#ifdef WB_PHASE_CLK
    unsigned clk_prev = (unsigned)clock();
#endif
    for (;;) {
        /* no CTA barrier between C and the next frame's A: both touch only the warp's own stream.  Whether any stream
           still has a frame to do is voted at the barrier that ends A (a stream that is done stays done). */
        const bool active = have && (pos + (unsigned)nin <= fill) && (n_out + (unsigned)NBITS <= a.sd_cap);
        WB_CLK(6);      /* C (thread 0's own) */
        const unsigned pos_next = pos + nin;
        const int dlt = CF32 ? ((warp + nst + (int)pos) & 1) : 0;     /* this frame's samples start at X[nst + dlt] */
        const int xo = nst - (NMEM - nin) + dlt;   /* X index of the first mixer sample */

        /* ================= A: stream warps ================= */
        if (active) {
#ifndef WB_NO_TMEM
            {
                float d2, d3;
                wb_tmem_ld8(taddr, est[0], est[1], est[2], est[3], stash.x, stash.y, d2, d3);
            }
#endif
            /* the leaf butterflies' table entries first: their latency hides behind the landing of the frame */
            constexpr int pp0 = 4, istr = Ndft / pp0;            /* 256 = 4 x 4 x 4 x 4, reference src/kiss_fft.c:311-338 */
            const int nwin = min(nin - Ndft, Ndft);
            int lbase[2];
            float lh[2][4];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int t = lane + 32 * b;
                lbase[b] = (t < istr) ? (int)__ldg(&p.perm[t * pp0]) : 0;
            }
#pragma unroll
            for (int b = 0; b < 2; b++) {
#pragma unroll
                for (int dd = 0; dd < 4; dd++) {
                    const int n = lbase[b] + dd * istr;
                    lh[b][dd] = (lane + 32 * b < istr && dd < pp0 && n < nwin) ? wb_ldg_keep(&p.hann[n]) : 0.0f;
                }
            }
            if (CF32) {
                /* this frame's samples were sent on their way (one TMA bulk copy, global -> shared) at the end of the
                   previous frame; wait for them */
                wb_mbar_wait(&MB[warp], (n_out / (unsigned)NBITS) & 1u);     /* the barrier's phase = frames so far, mod 2 */
            } else {
                /* converted formats go through registers: L2 hits (each frame is prefetched into L2 a frame ahead) */
                int sgv = sgc;
                asm volatile("" : "+r"(sgv));        /* opaque: recompute the row pointer here, do not carry it */
                const unsigned char *in = WB_ROW_IN(sgv);
                unsigned pre_lo[NPRE], pre_hi[NPRE];
#pragma unroll
                for (int q = 0; q < NPRE; q++) {
                    const int n = lane + 32 * q;
                    pre_lo[q] = pre_hi[q] = 0u;
                    if (n < NMAX && pos + n < fill) wb_load_raw<false>(fmt, in, pos + n, pre_lo[q], pre_hi[q]);
                }
#pragma unroll
                for (int q = 0; q < NPRE; q++) {
                    const int n = lane + 32 * q;
                    if (n < NMAX) X[nst + n] = wb_convert<false>(fmt, pre_lo[q], pre_hi[q]);
                }
            }
            /* the old samples the mixer reaches back to, right in front of the new ones (after the landing: a copy that
               started one sample early has written over the newest of them) */
            if (lane < nst) X[dlt + lane] = stash;
            {
                /* the frame after this one -> L2: one 128-byte line per lane, no registers held */
                int sgv = sgc;
                asm volatile("" : "+r"(sgv));
                const unsigned char *in = WB_ROW_IN(sgv);
                const unsigned long long b0 = (unsigned long long)pos_next * p.in_bps;
                const unsigned long long off = (b0 & ~127ULL) + 128ULL * lane;
                if (off < b0 + (unsigned long long)NMAX * p.in_bps && off < (unsigned long long)fill * p.in_bps)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(in + off));
            }
            __syncwarp();
            /* Estimator FFT (reference src/fsk.c:583-628, src/kiss_fft.c:238-306): decimation in time, the
               reference's butterfly order.  Leaf level fused with the window: leaf butterfly t combines the
               inputs perm[t*p] + d*Ndft/p, of which only those below nwin = nin - Ndft are non-zero. */
            float2 *F = Y + 2;                       /* Y[0..1] can hold the tail of a skewed frame (see WB_FETCH_CF32) */
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int t = lane + 32 * b;
                if (t < istr) {
                    float2 f[4];
#pragma unroll
                    for (int dd = 0; dd < 4; dd++) {
                        const int n = lbase[b] + dd * istr;
                        f[dd] = make_float2(0.0f, 0.0f);
                        if (dd < pp0 && n < nwin) {
                            const float2 x = X[nst + dlt + n];
                            f[dd].x = __fmul_rn(lh[b][dd], x.x); f[dd].y = __fmul_rn(lh[b][dd], x.y);
                        }
                    }
                    if (pp0 == 4) {     /* reference src/kiss_fft.c:44-90 with m = 1: all twiddles are tw[0] */
                        const float2 s0 = wb_ucmul(f[1]), s1 = wb_ucmul(f[2]), s2 = wb_ucmul(f[3]);
                        const float2 s5 = make_float2(__fsub_rn(f[0].x, s1.x), __fsub_rn(f[0].y, s1.y));
                        const float2 aa = make_float2(__fadd_rn(f[0].x, s1.x), __fadd_rn(f[0].y, s1.y));
                        const float2 s3 = make_float2(__fadd_rn(s0.x, s2.x), __fadd_rn(s0.y, s2.y));
                        const float2 s4 = make_float2(__fsub_rn(s0.x, s2.x), __fsub_rn(s0.y, s2.y));
                        /* (stream regions are only 8-byte aligned: float2 stores) */
                        F[wb_fidx(4 * t)] = make_float2(__fadd_rn(aa.x, s3.x), __fadd_rn(aa.y, s3.y));
                        F[wb_fidx(4 * t + 1)] = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
                        F[wb_fidx(4 * t + 2)] = make_float2(__fsub_rn(aa.x, s3.x), __fsub_rn(aa.y, s3.y));
                        F[wb_fidx(4 * t + 3)] = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
                    } else {            /* reference src/kiss_fft.c:22-42 with m = 1 */
                        const float2 tt = wb_ucmul(f[1]);
                        F[wb_fidx(2 * t)] = make_float2(__fadd_rn(f[0].x, tt.x), __fadd_rn(f[0].y, tt.y));
                        F[wb_fidx(2 * t + 1)] = make_float2(__fsub_rn(f[0].x, tt.x), __fsub_rn(f[0].y, tt.y));
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int L = 1; L < 3; L++) {                   /* middle levels: radix 4, twiddles from shared memory */
                const int sh = 2 * L, mm = 1 << sh, fs = Ndft >> (sh + 2);
                /* m <= 16 < 32: both butterflies of a lane (t = lane, lane + 32) share k and hence the twiddles */
                const int k = lane & (mm - 1);
                const float2 w1 = TW[k * fs], w2 = TW[2 * k * fs], w3 = TW[3 * k * fs];
#pragma unroll
                for (int t = lane; t < (Ndft >> 2); t += 32) {
                    const int base = ((t >> sh) << (sh + 2)) + k;
                    const int i0 = wb_fidx(base), i1 = wb_fidx(base + mm), i2 = wb_fidx(base + 2 * mm), i3 = wb_fidx(base + 3 * mm);
                    const float2 f0 = F[i0], f1 = F[i1], f2 = F[i2], f3 = F[i3];
                    /* (packed products and sums, reference src/kiss_fft.c:44-90: the same floats in fewer instructions) */
                    const float2 s0 = wb_cmul2p(f1, w1);
                    const float2 s1 = wb_cmul2p(f2, w2);
                    const float2 s2 = wb_cmul2p(f3, w3);
                    const float2 s5 = wb_sub2p(f0, s1);
                    const float2 aa = wb_add2p(f0, s1);
                    const float2 s3 = wb_add2p(s0, s2);
                    const float2 s4 = wb_sub2p(s0, s2);
                    F[i2] = wb_sub2p(aa, s3);
                    F[i0] = wb_add2p(aa, s3);
                    F[i1] = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
                    F[i3] = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
                }
                __syncwarp();
            }
            /* top level (radix 4, m = Ndft/4, fstride 1) fused with the magnitude spectrum, band limits and
               IIR (reference src/fsk.c:610-628): butterfly k yields bins k and k + m, the only ones below
               Ndft/2; lane + 32 q is exactly the layout of est[] */
            float v[WB_NEQ];
#pragma unroll
            for (int q = 0; q < WB_NEQ; q++) v[q] = 0.0f;
            {
                const int mtop = Ndft >> 2;
#define WB_TOP_BFLY(K, Q0, Q1)                                                                        \
                do {                                                                                  \
                    const int k_ = (K);                                                               \
                    const float2 f0 = F[wb_fidx(k_)], f1 = F[wb_fidx(k_ + mtop)], f2 = F[wb_fidx(k_ + 2 * mtop)], f3 = F[wb_fidx(k_ + 3 * mtop)]; \
                    const float2 s0 = wb_cmul2p(f1, TW[k_]);                                          \
                    const float2 s1 = wb_cmul2p(f2, TW[2 * k_]);                                      \
                    const float2 s2 = wb_cmul2p(f3, TW[3 * k_]);                                      \
                    const float2 s5 = wb_sub2p(f0, s1);                                               \
                    const float2 aa = wb_add2p(f0, s1);                                               \
                    const float2 s3 = wb_add2p(s0, s2);                                               \
                    const float2 s4 = wb_sub2p(s0, s2);                                               \
                    const float2 o0 = wb_add2p(aa, s3);                                               \
                    const float2 o1 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));     \
                    float pw0 = __fadd_rn(__fmul_rn(o0.x, o0.x), __fmul_rn(o0.y, o0.y));              \
                    float pw1 = __fadd_rn(__fmul_rn(o1.x, o1.x), __fmul_rn(o1.y, o1.y));              \
                    if (k_ < p.f_min || k_ >= p.f_max - 1) pw0 = 0.0f;                                \
                    if (k_ + mtop < p.f_min || k_ + mtop >= p.f_max - 1) pw1 = 0.0f;                  \
                    est[Q0] = __fadd_rn(__fmul_rn(est[Q0], omt), __fmul_rn(__fsqrt_rn(pw0), p.tc));   \
                    est[Q1] = __fadd_rn(__fmul_rn(est[Q1], omt), __fmul_rn(__fsqrt_rn(pw1), p.tc));   \
                    v[Q0] = est[Q0]; v[Q1] = est[Q1];                                                 \
                } while (0)
                if (mtop == 64) { WB_TOP_BFLY(lane, 0, 2); WB_TOP_BFLY(lane + 32, 1, 3); }
                else { WB_TOP_BFLY(lane, 0, 1); }
#undef WB_TOP_BFLY
            }
            /* M maxima with +-f_zero blanking (reference src/fsk.c:635-654), then ascending order */
            int freqi[M];
#pragma unroll
            for (int m = 0; m < M; m++) {
                float bv = 0.0f; int bi = 0;
#pragma unroll
                for (int q = 0; q < WB_NEQ; q++)
                    if (lane + 32 * q < nh && v[q] > bv) { bv = v[q]; bi = lane + 32 * q; }
                /* warp argmax, first maximum wins: the values are non-negative floats, which order like their bit
                   patterns, so two integer warp reductions do it (largest value, then smallest bin holding it) */
                {
                    const unsigned vb = __float_as_uint(bv);
                    const unsigned mx = __reduce_max_sync(0xffffffffu, vb);
                    const int cand = (vb == mx && mx != 0u) ? bi : 0x7fffffff;
                    const int bmin = __reduce_min_sync(0xffffffffu, cand);
                    bi = (mx == 0u) ? 0 : bmin;
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
            /* the samples the next frame's mixer reaches back to (reference src/fsk.c:851 keeps 4 Ts, uses at most
               2 Ts + Ts/2) sit where the mixer products are about to land: lift them into registers */
            if (lane < nst) stash = X[nin + dlt + lane];
#ifndef WB_NO_TMEM
            wb_tmem_st8(taddr, est[0], est[1], est[2], est[3], stash.x, stash.y, 0.0f, 0.0f);
#endif
            if (lane == 0) {
                wb_fsk_sc &c = sc[warp];
                const bool first = c.pb[0] == 0;         /* fsk->f_est[0] < 1, reference src/fsk.c:729 */
#pragma unroll
                for (int m = 0; m < M; m++) { c.nb[m] = (short)freqi[m]; if (first) c.pb[m] = (short)freqi[m]; }
                c.nin = (short)nin; c.flags = 1 | (dlt << 1);
            }
        } else if (lane == 0) {
            sc[warp].flags = 0;
        }
        if (!__syncthreads_or(active)) break;
        WB_CLK(0);

        /* ========== B1: warps 0..W-1, lane = (tone, stream): oscillator + down-mix ========== */
        /* The oscillator recurrence is sequential but does not depend on the samples, and a bare recurrence step
           (6 flops) is several times cheaper than a recurrence + down-mix step.  So W warps all run the SAME 28
           chains: warp j spins the bare recurrence up to its segment start b1_seg[j] and only then mixes its own
           segment [b1_seg[j], b1_seg[j+1]) of the frame.  Segment lengths shrink geometrically so all warps
           finish together; the redundant recurrences cost issue slots that are idle anyway and cut the
           latency of the phase by about W/2.  (Segments are multiples of 8: the in-place swizzled stores of one
           warp never touch samples another warp still has to read.) */
        float2 ph_end = make_float2(0.0f, 0.0f);
        bool ph_store = false;
#ifdef WB_PHASE_CLK
        const unsigned bq0 = (unsigned)clock();
        unsigned bq1 = bq0, bq2 = bq0, bq3 = bq0;
#endif
        if (warp < p.b1_w) {
            const int m = lane / spb, s = lane - m * spb;
            const bool mine = lane < M * spb && (sc[min(s, spb - 1)].flags & 1);
            /* the tone lanes of one stream read the same samples and tone 0 overwrites them in place: the lanes
               that take part meet at a __syncwarp between the loads and the stores of every batch */
            const unsigned bmask = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                wb_fsk_sc &c = sc[s];
                float2 *Xs = reinterpret_cast<float2 *>(regions + (size_t)s * sreg);
                const int fnin = c.nin, nold = NMEM - fnin;
                const int nin_idx = (fnin < N_) ? 0 : (fnin == N_ ? 1 : 2);
                const int pb = c.pb[m], nbn = c.nb[m];
                float2 ph = c.phi_c[m];
                ph = wb_cmul2(__ldg(&p.back[nin_idx * nh + pb]), ph);     /* reference src/fsk.c:756-759 */
                float2 d = __ldg(&p.dphi[pb]);
                const float2 dnew = __ldg(&p.dphi[nbn]);
                const int sdl = (c.flags >> 1) & 1;            /* that stream's landing skew this frame */
                const float2 *src = Xs + (nst - nold) + sdl;
                float2 *dst = (m == 0) ? Xs + (nst - nold) + sdl : Xs + XLEN + (m - 1) * ylen;
                const int nold_hi = 2 * TS + TS / 2;          /* >= every possible nold */
                const int seg0 = p.b1_seg[warp], seg1 = p.b1_seg[warp + 1];
#ifdef WB_PHASE_CLK
                asm volatile("" :: "f"(ph.x), "f"(d.x), "f"(dnew.x));
                bq1 = (unsigned)clock();
#endif
                /* old -> new samples: comp_normalize + this frame's tone, reference src/fsk.c:787-788 */
#define WB_B1_SWITCH()                                                                                  \
                do {                                                                                    \
                    const float av = __fsqrt_rn(__fadd_rn(__fmul_rn(ph.x, ph.x), __fmul_rn(ph.y, ph.y))); \
                    ph.x = __fdiv_rn(ph.x, av); ph.y = __fdiv_rn(ph.y, av);                             \
                    d = dnew;                                                                           \
                } while (0)
                /* cmult(sample, cconj(phi)) then phi *= dphi: reference src/fsk.c:794-798 */
#define WB_B1_STEP(J)                                                                                   \
                do {                                                                                    \
                    xv[J] = wb_mixp(xv[J], ph);                                                         \
                    ph = wb_cmul2p(ph, d);                                                              \
                } while (0)
                /* (0) bare recurrence up to the segment start */
                int n = 0;
                {
                    /* the switch sits at step 2 Ts - Ts/2, 2 Ts or 2 Ts + Ts/2 (nin = N + Ts/2, N, N - Ts/2): straight runs
                       between those three points instead of a compare and a branch on every one of the first steps */
                    const int na = min(seg0, nold_hi + 1);
                    if (na == nold_hi + 1) {
                        constexpr int S0 = 2 * TS - TS / 2, HS = TS / 2;
#pragma unroll
                        for (int j = 0; j < S0; j++) ph = wb_cmul2p(ph, d);
#pragma unroll
                        for (int g = 0; g < 3; g++) {
                            if (S0 + g * HS == nold) WB_B1_SWITCH();
                            if (g < 2) {
#pragma unroll
                                for (int j = 0; j < HS; j++) ph = wb_cmul2p(ph, d);
                            } else {
                                ph = wb_cmul2p(ph, d);         /* step nold_hi itself */
                            }
                        }
                        n = na;
                    } else {
#pragma unroll 1
                        for (; n < na; n++) {
                            if (n == nold) WB_B1_SWITCH();
                            ph = wb_cmul2p(ph, d);
                        }
                    }
#ifdef WB_PHASE_CLK
                    asm volatile("" :: "f"(ph.x));
                    bq2 = (unsigned)clock();
#endif
#pragma unroll 1
                    for (; n + 32 <= seg0; n += 32) {        /* long bodies: the loop's back edge costs the chain a bubble */
#pragma unroll
                        for (int j = 0; j < 32; j++) ph = wb_cmul2p(ph, d);
                    }
#pragma unroll 1
                    for (; n + 8 <= seg0; n += 8) {
#pragma unroll
                        for (int j = 0; j < 8; j++) ph = wb_cmul2p(ph, d);
                    }
#pragma unroll 1
                    for (; n < seg0; n++) ph = wb_cmul2p(ph, d);
                }
#ifdef WB_PHASE_CLK
                asm volatile("" :: "f"(ph.x));
                bq3 = (unsigned)clock();
                if (lane == 0 && warp == p.b1_w - 1) {
                    atomicAdd(&wb_phase_clk[4], (unsigned long long)(bq1 - bq0));
                    atomicAdd(&wb_phase_clk[5], (unsigned long long)(bq2 - bq1));
                    atomicAdd(&wb_phase_clk[7], (unsigned long long)(bq3 - bq2));
                }
#endif
                /* batches of eight steps: all eight samples are loaded before the (in-place, swizzled) stores */
                int n0 = seg0;
                /* (a) the batches that can contain the switch */
#pragma unroll 1
                for (; n0 <= nold_hi && n0 + 8 <= seg1; n0 += 8) {
                    const int kb = SWZ ? ((n0 >> 4) & 7) : 0;
                    float2 xv[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) xv[j] = src[n0 + j];
                    __syncwarp(bmask);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (n0 + j == nold) WB_B1_SWITCH();
                        WB_B1_STEP(j);
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) dst[(n0 ^ kb) ^ j] = xv[j];          /* n0 is a multiple of 8: n0 + (j ^ kb) */
                }
                /* (b) full batches, branch-free */
#pragma unroll 1
                for (; n0 + 8 <= seg1; n0 += 8) {
                    const int kb = SWZ ? ((n0 >> 4) & 7) : 0;
                    float2 xv[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) xv[j] = src[n0 + j];
                    __syncwarp(bmask);
#pragma unroll
                    for (int j = 0; j < 8; j++) WB_B1_STEP(j);
#pragma unroll
                    for (int j = 0; j < 8; j++) dst[(n0 ^ kb) ^ j] = xv[j];          /* n0 is a multiple of 8: n0 + (j ^ kb) */
                }
                /* (c) a partial batch only ends the last segment (its swizzle key is 0) */
                {
                    const int kb = SWZ ? ((n0 >> 4) & 7) : 0;
#pragma unroll 1
                    for (int j = 0; n0 + j < seg1; j++) {
                        float2 xv[1];
                        xv[0] = src[n0 + j];
                        __syncwarp(bmask);
                        if (n0 + j == nold) WB_B1_SWITCH();
                        WB_B1_STEP(0);
                        dst[n0 + (j ^ kb)] = xv[0];
                    }
                }
#undef WB_B1_STEP
#undef WB_B1_SWITCH
                /* the last segment ends the frame; phi_c is stored after the barrier, when every warp has read it */
                ph_end = ph;
                ph_store = warp == p.b1_w - 1;
            }
#ifdef WB_PHASE_CLK
            asm volatile("" :: "f"(ph_end.x));
            if (lane == 0) atomicAdd(&wb_phase_clk[8 + warp], (unsigned long long)((unsigned)clock() - bq0));
#endif
        }
        __syncthreads();
        WB_CLK(1);
        if (ph_store) sc[lane % spb].phi_c[lane / spb] = ph_end;

        /* ================= B2: stream warps: Ts-tap integrator sums, |.|^2 over tones ================= */
        /* f_int[m][i] = sum of the Ts ring-buffer slots after mixer step i*step + Ts - 1, added in slot order
           (reference src/fsk.c:835-838): slot j holds the product of the step n in [i*step, i*step + Ts) with
           n mod Ts == j.  In place: output i overwrites product i. */
        if (active && blocked) {
            /* step = 1: lane u owns outputs Ts*u .. Ts*u + Ts - 1.  It needs products Ts*u .. Ts*u + 2Ts - 2; output
               Ts*u + t adds, in slot order, the first t products of block u+1 (a running prefix shared by all t)
               and then products t .. Ts-1 of block u.  e_i = sum over tones of |f_int|^2 (reference src/fsk.c:864-867)
               goes to B3 through E. */
            static_assert(NBLK <= 64, "two rounds of 32 lanes cover the blocks");
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int u = 32 * rr + lane;
                const bool valid = u < NBLK;
                float e[TS];
#pragma unroll
                for (int t = 0; t < TS; t++) e[t] = 0.0f;
#pragma unroll
                for (int m = 0; m < M; m++) {
                    float2 *P = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                    float2 vv[2 * TS - 1];
                    if (valid) {
                        /* swizzled index 8 u + (o ^ k) = (8 u ^ k) ^ o for Ts = 8: one XOR with an immediate per access */
                        const int k0 = SWZ ? ((u >> 1) & 7) : 0, k1 = SWZ ? (((u + 1) >> 1) & 7) : 0;
                        const int b0 = (TS * u) ^ k0, b1 = (TS * (u + 1)) ^ k1;
#pragma unroll
                        for (int o = 0; o < TS; o++) vv[o] = SWZ ? P[b0 ^ o] : P[TS * u + o];
#pragma unroll
                        for (int o = 0; o < TS - 1; o++) vv[TS + o] = SWZ ? P[b1 ^ o] : P[TS * (u + 1) + o];
                    }
                    __syncwarp();
                    if (valid) {
                        const int k0 = SWZ ? ((u >> 1) & 7) : 0;
                        /* (real and imaginary sums side by side in packed additions: the same floats, half the instructions) */
                        float2 q = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int t = 0; t < TS; t++) {
                            float2 sacc;
                            int o0;
                            if (t == 0) { sacc = vv[0]; o0 = 1; }
                            else {
                                if (t == 1) q = vv[TS];
                                else q = wb_add2p(q, vv[TS + t - 1]);
                                sacc = q; o0 = t;
                            }
#pragma unroll
                            for (int o = 0; o < TS; o++)
                                if (o >= o0) sacc = wb_add2p(sacc, vv[o]);
                            if (SWZ) P[((TS * u) ^ k0) ^ t] = sacc; else P[TS * u + t] = sacc;
                            const float sr = sacc.x, si = sacc.y;
                            const float pw = __fadd_rn(__fmul_rn(sr, sr), __fmul_rn(si, si));
                            e[t] = (m == 0) ? pw : __fadd_rn(e[t], pw);      /* reference src/fsk.c:864-867 */
                        }
                    }
                    __syncwarp();
                }
                if (valid) {
                    /* E in output order (8-byte stores, 4-way bank conflicts: 16 extra wavefronts per frame) so that
                       B3 walks it with compile-time offsets */
                    float2 *eo = reinterpret_cast<float2 *>(E) + u * (TS / 2);
#pragma unroll
                    for (int q = 0; q < TS / 2; q++) eo[q] = make_float2(e[2 * q], e[2 * q + 1]);
                }
            }
        } else if (active) {
            /* general P: lanes = outputs, ring slots by modular arithmetic, E linear */
            const int step = stepc;
            for (int i0 = 0; i0 < nintc; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < nintc;
                float2 f[M];
                float e = 0.0f;
                if (valid) {
                    const int n0 = i * step, r = n0 % TS;
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        const float2 *P = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                        float sr = 0.0f, si = 0.0f;
#pragma unroll
                        for (int j = 0; j < TS; j++) {
                            int off = j - r;
                            if (off < 0) off += TS;
                            const float2 w = P[wb_phys<SWZ>(n0 + off)];
                            if (j == 0) { sr = w.x; si = w.y; }
                            else { sr = __fadd_rn(sr, w.x); si = __fadd_rn(si, w.y); }
                        }
                        f[m] = make_float2(sr, si);
                        const float pw = __fadd_rn(__fmul_rn(sr, sr), __fmul_rn(si, si));
                        e = (m == 0) ? pw : __fadd_rn(e, pw);
                    }
                }
                __syncwarp();
                if (valid) {
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        float2 *P = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                        P[wb_phys<SWZ>(i)] = f[m];
                    }
                    E[i] = e;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        WB_CLK(2);

        /* ============ B3: warp 0 = re, warp 1 = im, lane = stream: fine-timing accumulation ============ */
        /* t_c = sum_i e_i * phi_ft[i] is strictly sequential (reference src/fsk.c:858-873): one dependent addition
           per output and component.  The two components run in two warps; phi_ft is a compile-time index into the
           kernel-parameter constant bank, so a step is one multiply (off the chain) and one add (on it), and the
           e_i arrive by 8-byte loads the compiler hoists ahead of the chain. */
        if (blocked) {
            if (warp < 2) {
                const int s = min(lane, spb - 1);
                if (lane < spb && (sc[s].flags & 1)) {
                    const float2 *q_ = reinterpret_cast<const float2 *>(
                        reinterpret_cast<const float2 *>(regions + (size_t)s * sreg) + XLEN + blen);
                    /* warp 0: real parts, warp 1: imaginary parts of phi_ft */
                    const float *ct = reinterpret_cast<const float *>(p.pftc4[warp]);
                    const float tacc = p.pft_steady ? wb_b3_chain<TS, NBLK>(q_, ct) : wb_b3_chain_generic<TS, NBLK>(q_, ct);
                    if (warp == 0) sc[s].tcr = tacc; else sc[s].tci = tacc;
                }
            }
        } else if (warp == 0) {
            /* general P: lane = (re/im, stream) */
            const int b3c = (lane < spb) ? 0 : 1, b3s = min(lane - b3c * spb, spb - 1);
            if ((lane < 2 * spb) && (sc[b3s].flags & 1)) {
                const float *Es = reinterpret_cast<const float *>(
                    reinterpret_cast<const float2 *>(regions + (size_t)b3s * sreg) + XLEN + blen);
                float tacc = 0.0f;
                for (int i = 0; i < nintc; i++) {
                    const float w = reinterpret_cast<const float *>(p.pftc4[b3c])[i];
                    tacc = __fadd_rn(tacc, __fmul_rn(Es[i], w));                          /* reference src/fsk.c:870 */
                }
                if (b3c) sc[b3s].tci = tacc; else sc[b3s].tcr = tacc;
            }
        }
        __syncthreads();
        WB_CLK(3);

        /* ================= C: stream warps, lanes = symbols ================= */
        if (active) {
            wb_fsk_sc &c = sc[warp];
            /* timing estimate -> resampling offsets, ppm, next nin (reference src/fsk.c:876-907): every lane of the
               stream warp computes the same scalars, lane 0 keeps the state */
            const float tcr = c.tcr, tci = c.tci, norm_old = c.norm, ppm_old = c.ppm;
            const bool nan = isnan(tcr) || isnan(tci);     /* reference src/fsk.c:878-880 */
            int nin_next = nin, low = 0, high = 0;
            float fract = 0.0f, norm = norm_old, ppm = ppm_old, rx_timing = 0.0f;
            if (!nan) {
                /* reference: atan2f(..) / (2 * M_PI), a double division rounded to float.  Multiplying by the double
                   nearest to 1/(2 pi) gives the same float for EVERY float in [-3.2, 3.2] (checked exhaustively,
                   tests/test_hostmath.py::test_division_shortcuts) and takes the division routine off the chain
                   every stream warp waits on at the start of this phase */
                norm = (float)((double)wb_atan2f(tci, tcr) * WB_INV_2PI);
                rx_timing = __fmul_rn(norm, (float)Pc);
                const float dn = __fsub_rn(norm, norm_old);
                if ((double)fabsf(dn) < .2) {
                    /* 1e6 * dn / Nsym with Nsym = 48: the same float from a multiplication for every |dn| < 0.2 (same test) */
                    static_assert(NSYM == 48, "WB_INV_48");
                    const float appm = (float)((1e6 * (double)dn) * WB_INV_48);
                    ppm = (float)(.9 * (double)ppm_old + .1 * (double)appm);
                }
                if ((double)norm > 0.25) nin_next = N_ + TS / 2;
                else if ((double)norm < -0.25) nin_next = N_ - TS / 2;
                else nin_next = N_;
                low = (int)floorf(rx_timing);
                fract = __fsub_rn(rx_timing, (float)low);
                high = (int)ceilf(rx_timing);
            }
            __syncwarp();
            if (lane == 0 && !nan) { c.norm = norm; c.ppm = ppm; c.rx_timing = rx_timing; }
            int sgw = sgc;
            asm volatile("" : "+r"(sgw));
            float *out = WB_ROW_SD(sgw) + n_out;
            if (!nan) {
                const float omf = __fsub_rn(1.0f, fract);
                for (int i = lane; i < NSYM; i += 32) {
                    const int stt = (i + 1) * Pc;
                    const int il = wb_phys<SWZ>(stt + low), ih = wb_phys<SWZ>(stt + high);
                    float tm[M];
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        const float2 *fi = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                        const float2 lo = fi[il], hi = fi[ih];
                        const float tr = __fadd_rn(__fmul_rn(omf, lo.x), __fmul_rn(fract, hi.x));
                        const float ti = __fadd_rn(__fmul_rn(omf, lo.y), __fmul_rn(fract, hi.y));
                        tm[m] = __fsqrt_rn(__fadd_rn(__fmul_rn(tr, tr), __fmul_rn(ti, ti)));
                    }
                    if (M == 2) {
                        __stcs(out + i, __fsub_rn(tm[0], tm[1]));            /* reference src/fsk.c:966 */
                    } else {                                                  /* reference src/fsk.c:969-979 */
                        float b1 = -tm[0], b0 = -tm[0];
                        b1 = __fadd_rn(b1, tm[1]);  b0 = __fadd_rn(b0, -tm[1]);
                        b1 = __fadd_rn(b1, -tm[2]); b0 = __fadd_rn(b0, tm[2]);
                        b1 = __fadd_rn(b1, tm[3]);  b0 = __fadd_rn(b0, tm[3]);
                        __stcs(out + 2 * i + 1, b1); __stcs(out + 2 * i, b0);
                    }
                }
            } else {
                /* NaN guard: the reference returns before writing, so the caller's buffer still holds
                   the previous frame's values (zeros before the first frame) */
                for (int i = lane; i < NBITS; i += 32)
                    out[i] = (n_out >= (unsigned)NBITS) ? out[i - NBITS] : st->sd_last[i];
            }
            if (HARD) {   /* a separate instantiation: the soft-only kernel keeps its register allocation */
                /* rx_bits of fsk_demod(), reference src/fsk.c:936-959: the tone with the largest |t|^2 (first one
                   wins a tie), NOT the sign of the soft decision (for 4-FSK the two disagree under noise) */
                unsigned char *hb = a.hard + (size_t)sgw * (WB_HARD_PRE + (size_t)a.sd_cap) + WB_HARD_PRE + n_out;
                if (!nan) {
                    const float omf = __fsub_rn(1.0f, fract);
                    for (int i = lane; i < NSYM; i += 32) {
                        const int stt = (i + 1) * Pc;
                        const int il = wb_phys<SWZ>(stt + low), ih = wb_phys<SWZ>(stt + high);
                        float mx = 0.0f;
                        int sym = 0;
#pragma unroll
                        for (int m = 0; m < M; m++) {
                            const float2 *fi = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                            const float2 lo = fi[il], hi = fi[ih];
                            const float tr = __fadd_rn(__fmul_rn(omf, lo.x), __fmul_rn(fract, hi.x));
                            const float ti = __fadd_rn(__fmul_rn(omf, lo.y), __fmul_rn(fract, hi.y));
                            const float t2 = __fadd_rn(__fmul_rn(tr, tr), __fmul_rn(ti, ti));
                            if (m == 0) mx = t2;
                            else if (t2 > mx) { mx = t2; sym = m; }
                        }
                        if (M == 2) hb[i] = (unsigned char)(sym == 1);
                        else { hb[2 * i + 1] = (unsigned char)(sym & 1); hb[2 * i] = (unsigned char)((sym & 2) >> 1); }
                    }
                } else {
                    /* NaN guard: the caller's bit buffer keeps the previous frame's bits (for the first frame of a
                       chunk they sit right before it in the row, where the previous launch left them) */
                    for (int i = lane; i < NBITS; i += 32) hb[i] = hb[i - NBITS];
                }
            }
            if (p.stats && !nan) {
                /* Eb/N0 terms, reference src/fsk.c:985-1010: meanebno / stdebno accumulate over the symbols in order
                   (every lane repeats the 48-step sum from shuffled-in values), the log10 is left to the host */
                const float omf = __fsub_rn(1.0f, fract);
                float mxv[2] = {0.0f, 0.0f};
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int i = lane + 32 * rr;
                    if (i < NSYM) {
                        const int stt = (i + 1) * Pc;
                        const int il = wb_phys<SWZ>(stt + low), ih = wb_phys<SWZ>(stt + high);
                        float mx = 0.0f;
#pragma unroll
                        for (int m = 0; m < M; m++) {
                            const float2 *fi = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                            const float2 lo = fi[il], hi = fi[ih];
                            const float tr = __fadd_rn(__fmul_rn(omf, lo.x), __fmul_rn(fract, hi.x));
                            const float ti = __fadd_rn(__fmul_rn(omf, lo.y), __fmul_rn(fract, hi.y));
                            const float t2 = __fadd_rn(__fmul_rn(tr, tr), __fmul_rn(ti, ti));
                            mx = (m == 0 || t2 > mx) ? t2 : mx;
                        }
                        mxv[rr] = mx;
                    }
                }
                float meane = 0.0f, stde = 0.0f;
                for (int i = 0; i < NSYM; i++) {
                    const float v = __shfl_sync(0xffffffffu, (i < 32) ? mxv[0] : mxv[1], i & 31);
                    stde = __fadd_rn(stde, v);
                    meane = __fadd_rn(meane, __fsqrt_rn(v));
                }
                meane = __fdiv_rn(meane, (float)NSYM);
                stde = __fsub_rn(__fdiv_rn(stde, (float)NSYM), __fmul_rn(meane, meane));
                stde = (stde > 0.0f) ? (float)sqrt((double)stde) : 0.0f;
                if (lane == 0) {
                    st->eb_arg[st->eb_count & 31] = (float)((1e-6 + (double)meane) / (1e-6 + (double)stde));
                    st->eb_count = st->eb_count + 1;
                    st->eye_high = high;
                }
                /* eye-diagram tap: the first WB_EYE_KEEP integrator outputs of every tone */
#pragma unroll
                for (int m = 0; m < M; m++) {
                    const float2 *fi = (m == 0) ? X + xo : Y + (m - 1) * ylen;
                    for (int i = lane; i < WB_EYE_KEEP && i < nintc; i += 32) st->eye_fint[m][i] = fi[wb_phys<SWZ>(i)];
                }
            }
            __syncwarp();
            /* the integrator outputs in X are consumed: send the next frame's samples on their way (if there is one) */
            if (CF32 && (pos_next + (unsigned)nin_next <= fill) && (n_out + 2u * (unsigned)NBITS <= a.sd_cap)) {
                const unsigned pn = pos_next;
                WB_FETCH_CF32(pn, nin_next);
            }
            if (lane == 0) {
#pragma unroll
                for (int m = 0; m < M; m++) c.pb[m] = c.nb[m];              /* fsk->f_est = this frame's, :846 */
                if (a.frame_log && n_out / NBITS < (unsigned)a.log_cap) {       /* n_out / NBITS = frames so far */
                    float *l = a.frame_log + ((size_t)sg * a.log_cap + n_out / NBITS) * 8;
                    l[0] = (float)nin;
                    for (int m = 0; m < 4; m++) l[1 + m] = m < M ? (float)c.nb[m] : 0.0f;
                    l[5] = norm; l[6] = ppm; l[7] = nan ? c.rx_timing : rx_timing;
                }
            }
            n_out += NBITS;
            pos = pos_next;
            nin = nin_next;
            __syncwarp();
        }
    }

    /* ---- write the state back ---- */
#ifndef WB_NO_TMEM
    {
        float d2, d3;
        wb_tmem_ld8(taddr, est[0], est[1], est[2], est[3], stash.x, stash.y, d2, d3);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tmem_base) : "memory");
#endif
    if (have) {
        const wb_fsk_sc &c = sc[warp];
#pragma unroll
        for (int q = 0; q < WB_NEQ; q++)
            if (lane + 32 * q < nh) st->fft_est[lane + 32 * q] = est[q];
        if (lane < nst) st->samp_old[lane] = stash;
        if (n_out >= (unsigned)NBITS) {
            const float *lastf = WB_ROW_SD(sg) + n_out - NBITS;
            for (int i = lane; i < NBITS; i += 32) st->sd_last[i] = lastf[i];
            if (HARD) {
                unsigned char *hrow = a.hard + (size_t)sg * (WB_HARD_PRE + (size_t)a.sd_cap) + WB_HARD_PRE;
                for (int i = lane; i < NBITS; i += 32) hrow[i - NBITS] = hrow[n_out - NBITS + i];
            }
        }
        const unsigned char *in = WB_ROW_IN(sg);
        const unsigned long long rem = fill - pos;
        const unsigned long long dstpos = (unsigned long long)a.headroom - rem;   /* remainder ends at the headroom mark */
        if (a.compact && rem > 0 && pos > dstpos) {
            /* less than one frame is left over: park it right before the headroom mark so that every
               stream's next wb_feed lands at the same row offset (one strided copy for all streams).
               dst < src, ascending copy through registers. */
            const int bps = p.in_bps;
            unsigned char *row = const_cast<unsigned char *>(in);
            const unsigned long long nbytes = rem * bps, sb = (unsigned long long)pos * bps, db = dstpos * bps;
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
            const unsigned long long pos_start = st->in_pos;
            for (int m = 0; m < M; m++) { st->phi_c[m] = c.phi_c[m]; st->fbin[m] = c.pb[m]; }
            st->norm_rx_timing = c.norm; st->ppm = c.ppm; st->rx_timing = c.rx_timing;
            st->nin = nin; st->frames += n_out / NBITS;
            if (a.compact) { st->in_pos = dstpos; st->in_fill = a.headroom; }
            else { st->in_pos = pos; st->in_fill = fill; }
            wb_cursor &cu = a.cursor[sg];
            cu.in_fill = rem;
            cu.consumed = (unsigned long long)pos - pos_start;
            cu.n_sd = n_out;
            cu.nin = nin;
        }
    }
}

#endif /* WB_FSK_KERNEL_CUH */
