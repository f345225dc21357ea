// dependent-issue latency of the fp64 instructions the order-exact Welford chain is made of (one warp, one SM)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/fp64_latency tools/studies/fp64_latency.cu
#include <cstdio>
__global__ void k(double a, double b, long long *out, double *sink, int n) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = __fma_rn(x, b, a);
    long long t1 = clock64();
    double y = a;
    for (int i = 0; i < n; ++i) y = __dadd_rn(y, b);
    long long t2 = clock64();
    double z = a;
    for (int i = 0; i < n; ++i) z = __dmul_rn(z, b);
    long long t3 = clock64();
    // the Welford chain itself: delta = v - mu; q0 = delta*y; r = fma(-c, q0, delta); q = fma(r, y, q0); mu += q
    double mu = a, c = 20000.0, yy = 1.0 / 20000.0;
    for (int i = 0; i < n; ++i) {
        const double delta = __dsub_rn(b, mu);
        const double q0 = __dmul_rn(delta, yy);
        const double r = __fma_rn(-c, q0, delta);
        mu = __dadd_rn(mu, __fma_rn(r, yy, q0));
    }
    long long t4 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3;
        sink[0] = x + y + z + mu;
    }
}
int main() {
    long long *d, h[4]; double *s;
    cudaMalloc(&d, 32); cudaMalloc(&s, 8);
    const int n = 100000;
    for (int threads : {32, 64, 128}) {
        k<<<1, threads>>>(1.0, 0.999999, d, s, n);
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("%3d threads: DFMA %.1f, DADD %.1f, DMUL %.1f cycles per dependent op; Welford mean chain %.1f cycles per element\n", threads,
               (double)h[0] / n, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n);
    }
    return 0;
}
