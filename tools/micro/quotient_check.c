// gcc -O2 -ffp-contract=off -o quotient_check quotient_check.c -lm && ./quotient_check 2000000000 SEED
// Round 2: 16 seeds x 2e9 trials, 0 differences (wb_quot of wenet_b200/csrc/wb_ldpc_kernel.cuh).
// exhaustive-ish check: Markstein-corrected quotient == IEEE double division, for a = (double)float, b = positive double
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint64_t rng(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }
static inline double mk(double a, double b, double y)
{
    double q0 = a * y;
    double r0 = fma(-b, q0, a);
    double q1 = fma(r0, y, q0);
    double r1 = fma(-b, q1, a);
    return fma(r1, y, q1);
}
int main(int argc, char **argv)
{
    uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull, s = argc > 2 ? strtoull(argv[2], 0, 10) : 88172645463325252ull;
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t r = rng(&s), r2 = rng(&s);
        uint32_t fb = (uint32_t)r;                       // any float pattern
        float f; memcpy(&f, &fb, 4);
        if (!(f == f) || isinf(f)) continue;
        double a = (double)f, b;
        int mode = (r >> 32) & 7;
        if (mode < 4) {                                   // plausible means: a random double in [2^-20, 2^20)
            uint64_t bb = ((uint64_t)(1023 - 20 + (r2 >> 58) % 40) << 52) | (r2 & 0xfffffffffffffull);
            memcpy(&b, &bb, 8);
        } else if (mode < 6) {                            // wide exponent range
            uint64_t bb = ((uint64_t)(1023 - 200 + (r2 >> 54) % 400) << 52) | (r2 & 0xfffffffffffffull);
            memcpy(&b, &bb, 8);
        } else if (mode == 6) {                           // significands near all-ones / all-zeros / few bits
            uint64_t m = (r2 & 1) ? (0xfffffffffffffull ^ (r2 >> 40 & 0xfff)) : (r2 >> 40 & 0xfff) << ((r2 >> 8) % 40);
            uint64_t bb = ((uint64_t)(1023 - 10 + (r2 >> 58) % 20) << 52) | (m & 0xfffffffffffffull);
            memcpy(&b, &bb, 8);
        } else {                                          // b that is itself a sum of floats scaled (like a mean): k * float / n
            uint32_t gb = (uint32_t)r2 & 0x7fffffffu; float g; memcpy(&g, &gb, 4);
            if (!(g == g) || isinf(g) || g == 0) continue;
            b = (double)g * (double)(1 + (r2 >> 40) % 2580) / 2580.0;
            if (!(b > 1e-300 && b < 1e300)) continue;
        }
        double y = 1.0 / b;
        double q = a / b, m = mk(a, b, y);
        if (memcmp(&q, &m, 8) != 0 && !(q == 0 && m == 0)) {
            if (bad < 10) printf("MISMATCH a=%a b=%a q=%a mk=%a\n", a, b, q, m);
            bad++;
        }
    }
    printf("n=%llu bad=%llu\n", (unsigned long long)n, (unsigned long long)bad);
    return bad != 0;
}
