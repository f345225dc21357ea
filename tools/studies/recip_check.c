/* Exhaustive check of the reciprocal recurrence of the order-exact Welford kernels (csrc/preprocess.cu, WelfordStep<double>):
 * from y = RN(1/c), the reciprocals of c+1 .. c+8 (the kernel jumps by up to 4) each by THREE fused Newton steps e = fma(-d, y, 1), y = fma(y, e, y)
 * started at y. Must equal RN(1/d) (the IEEE divide) for every d the kernel can meet: c >= 16384 (the device
 * falls back to the exact reciprocal below), d < 2^32.
 * build: gcc -O2 -mfma -ffp-contract=off -fopenmp -o /tmp/recip_check tools/studies/recip_check.c -lm */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    /* default: every count below 2^32 (~25 s on 16 threads); argv[1] = log2 of the upper bound */
    const int64_t lo = 16384, hi = (int64_t)1 << (argc > 1 ? atoi(argv[1]) : 32);
    int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t c = lo; c < hi; ++c) {
        const double y0 = 1.0 / (double)c;
        for (int k = 1; k <= 8; ++k) {
            const double d = (double)(c + k);
            double y = y0;
            for (int it = 0; it < 3; ++it) {
                const double e = fma(-d, y, 1.0);
                y = fma(y, e, y);
            }
            if (y != 1.0 / d) ++bad;
        }
    }
    printf("c in [%lld, %lld), jumps 1..8: %lld mismatches\n", (long long)lo, (long long)hi, (long long)bad);
    return bad != 0;
}
