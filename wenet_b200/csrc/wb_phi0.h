/*
 * wb_phi0.h -- the phi0 nonlinearity of the sum-product decoder as ONE table
 * lookup + ONE compare (reference src/phi0.c:13-218 is a Q16 compare tree).
 *
 * The reference maps the float argument to q = (int32_t)(x*65536) (x86
 * cvttss2si: INT32_MIN on NaN/overflow) and returns a step function of q with
 * 103 steps (wb_phi0_brk / wb_phi0_val in wb_tables.h, measured from the
 * compiled reference).  Here the argument's own float bits select a bucket:
 *
 *     b = clamp((bits >> 17) - (113 << 6), 0, 1152)   exponent + 6 mantissa bits
 *
 * i.e. 64 buckets per octave over [2^-14, 16).  Every bucket holds at most one
 * breakpoint (checked when the table is built), so
 *
 *     phi0(x) = (x < thr[b]) ? vlo[b] : vhi[b]
 *
 * is exact for every float:  x < 2^-14 (and all negatives / -NaN) clamp to
 * bucket 0 (10.0); bucket 1152 is x >= 16: 0.0 below 32768, 10.0 from 32768 up
 * (the cvttss2si overflow quirk) and for NaN (the compare is false).
 * tests/test_hostmath.py sweeps every breakpoint neighbourhood and 1e7 random
 * floats against the compiled reference.
 */
#ifndef WB_PHI0_H
#define WB_PHI0_H

#include "wb_math.h"
#include "wb_tables.h"

#define WB_PHI0_EXP0 113                 /* biased exponent of 2^-14 */
#define WB_PHI0_NBUCKET (18 * 64)        /* [2^-14, 2^4) */
#define WB_PHI0_NENTRY (WB_PHI0_NBUCKET + 1)

typedef struct { float thr, vlo, vhi, pad; } wb_phi0_entry;
typedef struct { wb_phi0_entry e[WB_PHI0_NENTRY]; } wb_phi0_lut;

/* value of the reference step function at Q16 integer q >= 0 */
static inline float wb_phi0_step(int64_t q)
{
    int s = 0, k;
    for (k = 0; k < WB_PHI0_NSTEPS; k++) if ((int64_t)wb_phi0_brk[k] <= q) s = k;
    return wb_phi0_val[s];
}

/* returns 0 on success, -1 if some bucket would need two thresholds */
static inline int wb_phi0_build(wb_phi0_lut *lut)
{
    int b, k;
    for (b = 0; b < WB_PHI0_NBUCKET; b++) {
        int E = -14 + b / 64, j = b % 64;
        /* bucket [xs, xe) with xs = (64+j) * 2^(E-6); in Q16 scaled by a further 2^20 (always integral) */
        int64_t qs36 = ((int64_t)(64 + j)) << (E + 30);
        int64_t qe36 = ((int64_t)(64 + j + 1)) << (E + 30);
        int64_t q_first, q_last;
        int nb = 0;
        wb_phi0_entry en;
        q_first = qs36 >> 20;                         /* trunc(xs * 65536) */
        q_last = (qe36 - 1) >> 20;                    /* largest q reached inside the bucket */
        en.thr = 0.0f; en.pad = 0.0f;
        en.vlo = en.vhi = wb_phi0_step(q_first);
        for (k = 0; k < WB_PHI0_NSTEPS; k++) {
            int64_t B = wb_phi0_brk[k];
            if (B > q_first && B <= q_last) {
                nb++;
                en.thr = (float)((double)B / 65536.0);   /* exact */
                en.vhi = wb_phi0_val[k];
            }
        }
        if (nb > 1) return -1;
        lut->e[b] = en;
    }
    lut->e[WB_PHI0_NBUCKET].thr = 32768.0f;
    lut->e[WB_PHI0_NBUCKET].vlo = 0.0f;
    lut->e[WB_PHI0_NBUCKET].vhi = 10.0f;
    lut->e[WB_PHI0_NBUCKET].pad = 0.0f;
    return 0;
}

WB_HD int wb_phi0_bucket(float x)
{
    int b = ((int32_t)wb_f2u(x) >> 17) - (WB_PHI0_EXP0 << 6);
    b = b < 0 ? 0 : b;
    return b > WB_PHI0_NBUCKET ? WB_PHI0_NBUCKET : b;
}

WB_HD float wb_phi0_eval(const wb_phi0_lut *lut, float x)
{
    wb_phi0_entry en = lut->e[wb_phi0_bucket(x)];
    return (x < en.thr) ? en.vlo : en.vhi;
}


/* ---- pair form: one level of independent loads ------------------------------
 * pt[b] = {threshold inside bucket b (or +inf), value at the start of bucket b}; a bucket holds at most one
 * breakpoint, after which the value is the one the next bucket starts with:
 *     phi0(x) = (x < pt[b].thr) ? pt[b].val : pt[b+1].val
 * Entry NENTRY (one past the clamp bucket) = {+inf, 10}: x >= 32768 and NaN (cvttss2si overflow quirk). */
typedef struct { float thr, val; } wb_phi0_pair;
#define WB_PHI0_NPAIR (WB_PHI0_NENTRY + 1)
typedef struct { wb_phi0_pair pt[WB_PHI0_NPAIR]; } wb_phi0_pairs;     /* 1154 x 8 B */

static inline int wb_phi0_build_pairs(wb_phi0_pairs *t)
{
    wb_phi0_lut full;
    int b;
    if (wb_phi0_build(&full) != 0) return -1;
    for (b = 0; b < WB_PHI0_NENTRY; b++) {
        t->pt[b].val = full.e[b].vlo;
        t->pt[b].thr = (full.e[b].vlo != full.e[b].vhi || b == WB_PHI0_NBUCKET) ? full.e[b].thr : wb_u2f(0x7f800000u);
    }
    t->pt[WB_PHI0_NENTRY].thr = wb_u2f(0x7f800000u);
    t->pt[WB_PHI0_NENTRY].val = 10.0f;
    /* consistency: the value after a breakpoint must be what the next bucket starts with */
    for (b = 0; b < WB_PHI0_NBUCKET; b++)
        if (full.e[b].vlo != full.e[b].vhi && full.e[b].vhi != full.e[b + 1].vlo) return -2;
    return 0;
}

WB_HD float wb_phi0_eval_pairs(const wb_phi0_pairs *t, float x)
{
    const int b = wb_phi0_bucket(x);
    const wb_phi0_pair p0 = t->pt[b], p1 = t->pt[b + 1];
    return (x < p0.thr) ? p0.val : p1.val;
}


/* ---- flag form: one 4-byte load per call -------------------------------------
 * The decoder is bound by shared-memory bandwidth and the table look-ups are two thirds of its traffic, so the
 * common case must be one small load.  val[b] = value at the start of bucket b; only a bucket with a breakpoint
 * strictly inside needs more, and those are few and far apart: from 1 up every step of the reference is a multiple
 * of 1/16 = a bucket start; below 1 the steps sit one Q16 unit above 2^-j (first bucket of an octave) and at
 * 2^-j / sqrt 2 (bucket 26 of an octave), so no two of them share a group of 16 buckets (checked when the table is
 * built).  Such a bucket carries its value with the sign bit set and its group g = b >> 4 holds the threshold and the
 * value above it:
 *     v = val[b];  if (v has the sign bit) v = (x < gthr[g]) ? |v| : gvhi[g];
 * The clamp bucket (x >= 16) is flagged with threshold 32768: 0.0 below, 10.0 from there up and for NaN (the
 * cvttss2si overflow quirk, the compare is false). */
#define WB_PHI0_NGROUP (WB_PHI0_NBUCKET / 16 + 1)       /* 72 groups of 16 buckets + the clamp bucket */
typedef struct {
    float val[WB_PHI0_NENTRY + 3];                       /* 1153 used */
    float gthr[WB_PHI0_NGROUP + 3], gvhi[WB_PHI0_NGROUP + 3];
} wb_phi0_flag;

static inline int wb_phi0_build_flag(wb_phi0_flag *t)
{
    wb_phi0_lut full;
    int b, g, seen[WB_PHI0_NGROUP];
    if (wb_phi0_build(&full) != 0) return -1;
    for (g = 0; g < WB_PHI0_NGROUP + 3; g++) { t->gthr[g] = 0.0f; t->gvhi[g] = 0.0f; }
    for (g = 0; g < WB_PHI0_NGROUP; g++) seen[g] = 0;
    for (b = 0; b < (int)(sizeof(t->val) / sizeof(float)); b++) t->val[b] = 0.0f;
    for (b = 0; b < WB_PHI0_NENTRY; b++) {
        const wb_phi0_entry en = full.e[b];
        const int inside = (b == WB_PHI0_NBUCKET) || (en.vlo != en.vhi);     /* a compare is needed in this bucket */
        if (wb_f2u(en.vlo) >> 31) return -2;                                  /* the sign bit is ours */
        t->val[b] = en.vlo;
        if (inside) {
            g = b >> 4;
            if (seen[g]) return -3;                                           /* two such buckets in one group */
            seen[g] = 1;
            t->val[b] = wb_u2f(wb_f2u(en.vlo) | 0x80000000u);
            t->gthr[g] = en.thr;
            t->gvhi[g] = en.vhi;
        }
    }
    return 0;
}

WB_HD float wb_phi0_eval_flag(const wb_phi0_flag *t, float x)
{
    const int b = wb_phi0_bucket(x);
    float v = t->val[b];
    if (wb_f2u(v) >> 31) {
        const int g = b >> 4;
        v = (x < t->gthr[g]) ? wb_u2f(wb_f2u(v) & 0x7fffffffu) : t->gvhi[g];
    }
    return v;
}

#endif /* WB_PHI0_H */
