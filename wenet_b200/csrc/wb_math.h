/*
 * wb_math.h -- scalar numeric building blocks shared by the CUDA kernels.
 *
 * Everything here is written so that it compiles both as CUDA device code and
 * as plain C/C++ on the host (wb_hostmath.c wraps it for the CPU unit tests,
 * which compare each routine against glibc / x87 / the compiled reference).
 * All arithmetic is IEEE-754 round-to-nearest with NO fused multiply-add:
 * the kernels are built with -fmad=false, the host wrapper with
 * -ffp-contract=off, because the reference is x86-64 gcc -O3 without FMA.
 */
#ifndef WB_MATH_H
#define WB_MATH_H

#include <stdint.h>

#ifdef __CUDACC__
#define WB_HD __host__ __device__ __forceinline__
#else
#define WB_HD static inline
#include <math.h>
#include <string.h>
#endif

/* ---- bit casts --------------------------------------------------------- */
WB_HD uint32_t wb_f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
WB_HD float wb_u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
WB_HD uint64_t wb_d2u(double d)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
WB_HD double wb_u2d(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
WB_HD float wb_sqrtf(float x)
{
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
WB_HD float wb_divf(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
WB_HD int wb_clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
WB_HD uint64_t wb_mulhi64(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

/* ---- complex helpers (reference src/comp_prim.h:55-63 operand order) --- */
typedef struct { float r, i; } wb_cpx;

WB_HD wb_cpx wb_cmul(wb_cpx a, wb_cpx b)
{
    wb_cpx c;
    c.r = a.r * b.r - a.i * b.i;
    c.i = a.r * b.i + a.i * b.r;
    return c;
}

/* ---- atanf / atan2f ----------------------------------------------------
 * glibc 2.39 (the libm of this image and of the reference build) implements
 * atan2f on x86-64 with the generic fdlibm-derived float code
 * (sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c): a fixed sequence of float
 * operations, restated here so the device reproduces norm_rx_timing
 * (reference src/fsk.c:883) bit for bit.  tests/test_hostmath.py checks it
 * against the running libm on ~1e7 arguments.
 */
WB_HD float wb_atanf(float x)
{
    const float atanhi0 = 4.6364760399e-01f, atanhi1 = 7.8539812565e-01f,
                atanhi2 = 9.8279368877e-01f, atanhi3 = 1.5707962513e+00f;
    const float atanlo0 = 5.0121582440e-09f, atanlo1 = 3.7748947079e-08f,
                atanlo2 = 3.4473217170e-08f, atanlo3 = 7.5497894159e-08f;
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
                aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
                aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
                aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    float w, s1, s2, z, hi, lo;
    int32_t hx = (int32_t)wb_f2u(x);
    int32_t ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {                /* |x| >= 2^25 */
        if (ix > 0x7f800000) return x + x; /* NaN */
        if (hx > 0) return atanhi3 + atanlo3;
        return -atanhi3 - atanlo3;
    }
    if (ix < 0x3ee00000) {                 /* |x| < 0.4375 */
        if (ix < 0x31000000) return x;     /* |x| < 2^-29 */
        id = -1; hi = 0.0f; lo = 0.0f;
    } else {
        x = wb_u2f((uint32_t)ix);          /* fabsf */
        if (ix < 0x3f980000) {             /* |x| < 1.1875 */
            if (ix < 0x3f300000) {         /* 7/16 <= |x| < 11/16 */
                id = 0; hi = atanhi0; lo = atanlo0;
                x = wb_divf(2.0f * x - 1.0f, 2.0f + x);
            } else {                       /* 11/16 <= |x| < 19/16 */
                id = 1; hi = atanhi1; lo = atanlo1;
                x = wb_divf(x - 1.0f, x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) {         /* |x| < 2.4375 */
                id = 2; hi = atanhi2; lo = atanlo2;
                x = wb_divf(x - 1.5f, 1.0f + 1.5f * x);
            } else {                       /* 2.4375 <= |x| < 2^25 */
                id = 3; hi = atanhi3; lo = atanlo3;
                x = wb_divf(-1.0f, x);
            }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    z = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -z : z;
}

WB_HD float wb_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t hx = (int32_t)wb_f2u(x), hy = (int32_t)wb_f2u(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    int32_t k, m;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;  /* NaN */
    if (hx == 0x3f800000) return wb_atanf(y);               /* x == 1.0 */
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                /* 2*sign(x) + sign(y) */
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return 3.0f * pi_o_4 + tiny;
            default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;                  /* |y/x| > 2^60 */
    else if (hx < 0 && k < -60) z = 0.0f;                   /* |y|/x < -2^60 */
    else z = wb_atanf(wb_u2f(wb_f2u(wb_divf(y, x)) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return wb_u2f(wb_f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

/* ---- x87 extended-precision emulation ----------------------------------
 * reference src/mpdecode_core.c:593-595 writes `2.0L`, `4.0L`: on x86-64 gcc
 * evaluates  1.0/(2.0L*estvar + 1E-3)  and  4.0L*estEsN0*sd[i]  in 80-bit
 * long double (64-bit significand), then rounds to double / float.  These
 * helpers reproduce exactly that for the value ranges that occur (positive
 * finite normal operands; anything else falls back to plain double maths).
 */
typedef struct { uint64_t m; int e; } wb_x87;   /* value = m * 2^(e-63), m has bit 63 set */

WB_HD wb_x87 wb_x87_from_double_pos(double d)   /* d > 0, normal */
{
    uint64_t u = wb_d2u(d);
    wb_x87 r;
    r.m = ((u & 0x000fffffffffffffULL) | 0x0010000000000000ULL) << 11;
    r.e = (int)((u >> 52) & 0x7ff) - 1023;
    return r;
}

/* round-to-nearest-even of the 128-bit significand (hi:lo, hi has bit 63 set) to 64 bits */
WB_HD wb_x87 wb_x87_round(uint64_t hi, uint64_t lo, int e)
{
    wb_x87 r;
    uint64_t half = 0x8000000000000000ULL;
    if (lo > half || (lo == half && (hi & 1))) {
        hi += 1;
        if (hi == 0) { hi = half; e += 1; }
    }
    r.m = hi; r.e = e;
    return r;
}

WB_HD wb_x87 wb_x87_add_pos(wb_x87 a, wb_x87 b)  /* a, b > 0 */
{
    uint64_t hi, lo, bh, bl;
    int sh;
    if (a.e < b.e) { wb_x87 t = a; a = b; b = t; }
    sh = a.e - b.e;
    if (sh == 0) { bh = b.m; bl = 0; }
    else if (sh < 64) { bh = b.m >> sh; bl = b.m << (64 - sh); }
    else if (sh < 128) {
        bh = 0; bl = (sh == 64) ? b.m : (b.m >> (sh - 64));
        if (sh > 64 && (b.m << (128 - sh))) bl |= 1;   /* sticky */
    } else { bh = 0; bl = 1; }
    hi = a.m + bh; lo = bl;
    if (hi < a.m) {                                   /* carry out: shift right by one */
        uint64_t sticky = lo & 1;
        lo = (lo >> 1) | (hi << 63) | sticky;
        hi = (hi >> 1) | 0x8000000000000000ULL;
        return wb_x87_round(hi, lo, a.e + 1);
    }
    return wb_x87_round(hi, lo, a.e);
}

/* 1.0 / a, correctly rounded to 64 bits */
WB_HD wb_x87 wb_x87_recip(wb_x87 a)
{
    /* long division of 2^127 by a.m: quotient in (2^63, 2^64] */
    uint64_t rem, q = 0;
    int e = -a.e, i;
    wb_x87 r;
    if (a.m == 0x8000000000000000ULL) { r.m = a.m; r.e = -a.e; return r; }  /* exact power of two */
    /* 2^127 / m with m in (2^63, 2^64): first quotient bit (2^64 place) is 0; generate 65 bits:
       64 quotient bits + 1 guard, keeping the remainder for the sticky */
    rem = 0x8000000000000000ULL;      /* 2^127 >> 64 */
    /* rem < m always holds here (m > 2^63) */
    for (i = 0; i < 64; i++) {
        uint64_t top = rem >> 63;
        rem <<= 1;
        q <<= 1;
        if (top || rem >= a.m) { rem -= a.m; q |= 1; }
    }
    /* now q = floor(2^127 / m) in (2^63, 2^64) and rem = 2^127 mod m;
       1/a = (2^127/m) * 2^((-a.e-1) - 63) */
    {
        /* guard bit + sticky from the remainder */
        uint64_t top = rem >> 63, g;
        uint64_t rem2 = rem << 1;
        if (top || rem2 >= a.m) { g = 1; rem2 -= a.m; } else g = 0;
        e = -a.e - 1;
        if (g && (rem2 != 0 || (q & 1))) {
            q += 1;
            if (q == 0) { q = 0x8000000000000000ULL; e += 1; }
        }
    }
    r.m = q; r.e = e;
    return r;
}

/* round a positive x87 value to double (RN-even); assumes the result is a normal double */
WB_HD double wb_x87_to_double_pos(wb_x87 a)
{
    uint64_t m = a.m >> 11, rest = a.m & 0x7ff;
    int e = a.e;
    if (rest > 0x400 || (rest == 0x400 && (m & 1))) {
        m += 1;
        if (m == 0x0020000000000000ULL) { m >>= 1; e += 1; }
    }
    return wb_u2d(((uint64_t)(e + 1023) << 52) | (m & 0x000fffffffffffffULL));
}

/* estEsN0 = (double)(1.0 / (2.0L*estvar + 1E-3)), reference src/mpdecode_core.c:593 */
WB_HD double wb_esn0_from_var(double estvar)
{
    double two_v = 2.0 * estvar;                     /* exact */
    if (two_v == 0.0) return wb_x87_to_double_pos(wb_x87_recip(wb_x87_from_double_pos(1E-3)));
    if (!(two_v > 1e-300 && two_v < 1e300)) return 1.0 / (two_v + 1E-3);   /* negative / NaN / huge */
    return wb_x87_to_double_pos(wb_x87_recip(wb_x87_add_pos(wb_x87_from_double_pos(two_v),
                                                               wb_x87_from_double_pos(1E-3))));
}

/* llr = (float)(4.0L * estEsN0 * sd) with sd a float-valued double,
   reference src/mpdecode_core.c:594-595: 64-bit-significand product, then RN to float */
WB_HD float wb_llr_scale(double four_esn0, float sd)
{
    uint64_t uc = wb_d2u(four_esn0);
    uint32_t us = wb_f2u(sd);
    int ec = (int)((uc >> 52) & 0x7ff), es = (int)((us >> 23) & 0xff);
    uint64_t mc, ms, hi, lo;
    int e, lz;
    uint32_t sign = us & 0x80000000u, mant, rest_nonzero;
    uint64_t m64, low40;
    if (ec == 0 || ec == 0x7ff || es == 0 || es == 0xff || (uc >> 63))
        return (float)(four_esn0 * (double)sd);      /* zero, subnormal, inf, NaN: plain double */
    mc = (uc & 0x000fffffffffffffULL) | 0x0010000000000000ULL;   /* 53 bits */
    ms = (uint64_t)((us & 0x007fffffu) | 0x00800000u);           /* 24 bits */
    /* exact 77-bit product, left-aligned into hi:lo */
    mc <<= 11;                                        /* bit 63 set */
    ms <<= 40;                                        /* bit 63 set */
    hi = wb_mulhi64(mc, ms);
    lo = mc * ms;
    e = (ec - 1023) + (es - 127);                     /* value = (hi:lo) * 2^(e - 126) */
    lz = (hi >> 63) ? 0 : 1;
    if (lz) { hi = (hi << 1) | (lo >> 63); lo <<= 1; } else e += 1;
    /* first rounding: to 64 bits (x87 fmul) */
    {
        wb_x87 r = wb_x87_round(hi, lo, e);
        m64 = r.m; e = r.e;
    }
    /* second rounding: 64 -> 24 bits (fstps) */
    mant = (uint32_t)(m64 >> 40);
    low40 = m64 & 0xffffffffffULL;
    rest_nonzero = (low40 > 0x8000000000ULL) || (low40 == 0x8000000000ULL && (mant & 1));
    if (rest_nonzero) {
        mant += 1;
        if (mant == 0x01000000u) { mant >>= 1; e += 1; }
    }
    /* value = mant * 2^(e-23), mant in [2^23, 2^24) */
    {
        int ef = e + 127;
        if (ef <= 0 || ef >= 255) return (float)(four_esn0 * (double)sd);
        return wb_u2f(sign | ((uint32_t)ef << 23) | (mant & 0x007fffffu));
    }
}

/* The same value from ONE double multiplication wherever that is provably enough.  The reference rounds the exact
   77-bit product first to 64 bits (x87 fmul) and then to float; p = RN53(exact) rounded to float gives the same float
   unless a float rounding midpoint lies within one double ulp of the exact product -- then and only then the low 29 bits
   of p's significand are 2^28 - 1, 2^28 or 2^28 + 1 -- and those cases (3 in 2^29) take the exact 64-bit emulation.
   Zeros, subnormals, infinities, NaNs and results outside the normal float range are the plain double product in
   wb_llr_scale too.  (The decoder's prologue spent 8 % of the kernel's instructions in the emulation.) */
WB_HD float wb_llr_scale_fast(double four_esn0, float sd)
{
    const double p = four_esn0 * (double)sd;
    const uint32_t low = (uint32_t)wb_d2u(p) & 0x1fffffffu;
    if (low - 0x0fffffffu <= 2u) return wb_llr_scale(four_esn0, sd);
    return (float)p;
}

/* reciprocals that replace two double divisions of the frame-scalar chain (exhaustively checked over every float
   the divisions can see: wbh_check_div_2pi / wbh_check_div_48 in wb_hostmath.c) */
#define WB_INV_2PI 0.15915494309189535      /* the double nearest to 1 / 6.283185307179586 */
#define WB_INV_48  (1.0 / 48.0)

#endif /* WB_MATH_H */
