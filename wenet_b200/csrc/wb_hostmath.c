/*
 * wb_hostmath.c -- host build of the scalar helpers in wb_math.h / wb_phi0.h so
 * the CPU-only unit tests (tests/test_hostmath.py) can compare them with glibc,
 * real x87 long double arithmetic and the compiled reference.  Not part of the
 * product library; built by wenet_b200/csrc/Makefile as libwb_hostmath.so with
 * -ffp-contract=off.
 */
#include "wb_math.h"
#include "wb_phi0.h"

void wbh_atan2f(const float *y, const float *x, float *out, long n)
{
    long i;
    for (i = 0; i < n; i++) out[i] = wb_atan2f(y[i], x[i]);
}

void wbh_esn0_from_var(const double *v, double *out, long n)
{
    long i;
    for (i = 0; i < n; i++) out[i] = wb_esn0_from_var(v[i]);
}

void wbh_llr_scale(const double *c, const float *sd, float *out, long n)
{
    long i;
    for (i = 0; i < n; i++) out[i] = wb_llr_scale(c[i], sd[i]);
}

void wbh_llr_scale_fast(const double *c, const float *sd, float *out, long n)
{
    long i;
    for (i = 0; i < n; i++) out[i] = wb_llr_scale_fast(c[i], sd[i]);
}

/* the same expressions in genuine x87 long double, as gcc compiles the reference */
void wbh_esn0_from_var_x87(const double *v, double *out, long n)
{
    long i;
    for (i = 0; i < n; i++) out[i] = 1.0 / (2.0L * v[i] + 1E-3);
}

void wbh_llr_scale_x87(const double *c, const float *sd, float *out, long n)
{
    long i;
    for (i = 0; i < n; i++) { double s = sd[i]; out[i] = (float)((long double)c[i] * s); }
}

void wbh_phi0(const float *x, float *out, long n)
{
    long i;
    wb_phi0_lut lut;
    wb_phi0_build(&lut);
    for (i = 0; i < n; i++) out[i] = wb_phi0_eval(&lut, x[i]);
}

long wbh_phi0_pairs(const float *x, float *out, long n)
{
    long i;
    static wb_phi0_pairs t;
    int rc = wb_phi0_build_pairs(&t);
    if (rc) return rc;
    for (i = 0; i < n; i++) out[i] = wb_phi0_eval_pairs(&t, x[i]);
    return 0;
}

long wbh_phi0_flag(const float *x, float *out, long n)
{
    long i;
    static wb_phi0_flag t;
    int rc = wb_phi0_build_flag(&t);
    if (rc) return rc;
    for (i = 0; i < n; i++) out[i] = wb_phi0_eval_flag(&t, x[i]);
    return 0;
}

/* fine-timing oscillator table as wb_engine.cu: upload_tables builds it (reference src/fsk.c:858-873: phi_ft starts
   at 1 and is multiplied by comp_exp_j(2 pi Rs/(P Rs)) after every use, float arithmetic, never normalised) */
#include <math.h>
void wbh_pft_table(int P, int Rs, int n, float *re, float *im)
{
    float a = (float)(2 * M_PI * ((float)Rs / (float)(P * Rs)));
    float dr = cosf(a), di = sinf(a), pr = 1.0f, pi_ = 0.0f;
    for (int i = 0; i < n; i++) {
        re[i] = pr; im[i] = pi_;
        float nr = pr * dr - pi_ * di, ni = pr * di + pi_ * dr;     /* cmult, src/comp_prim.h (no FMA: -ffp-contract=off) */
        pr = nr; pi_ = ni;
    }
}

/* the two double divisions of the per-frame scalars (reference src/fsk.c:883,888) against the multiplications the
   kernel uses: number of floats for which they differ, over EVERY float the division can see */
#include <string.h>
long wbh_check_div_2pi(void)
{
    const double c = 6.283185307179586;        /* 2 * M_PI */
    float top = 3.2f;                          /* |atan2f| <= pi */
    uint32_t utop, u;
    long bad = 0;
    memcpy(&utop, &top, 4);
    for (u = 0; u <= utop; u++) {
        float a, q1, q2;
        memcpy(&a, &u, 4);
        q1 = (float)((double)a / c);
        q2 = (float)((double)a * WB_INV_2PI);
        if (memcmp(&q1, &q2, 4)) bad++;
        q1 = (float)((double)-a / c);
        q2 = (float)((double)-a * WB_INV_2PI);
        if (memcmp(&q1, &q2, 4)) bad++;
    }
    return bad;
}
long wbh_check_div_48(void)
{
    float top = 0.2f;                          /* the ppm update is gated by |dn| < .2 */
    uint32_t utop, u;
    long bad = 0;
    memcpy(&utop, &top, 4);
    for (u = 0; u <= utop; u++) {
        float a, q1, q2;
        double t;
        memcpy(&a, &u, 4);
        t = 1e6 * (double)a;
        q1 = (float)(t / (double)(float)48);
        q2 = (float)(t * WB_INV_48);
        if (memcmp(&q1, &q2, 4)) bad++;
        q1 = (float)(-t / (double)(float)48);
        q2 = (float)(-t * WB_INV_48);
        if (memcmp(&q1, &q2, 4)) bad++;
    }
    return bad;
}

