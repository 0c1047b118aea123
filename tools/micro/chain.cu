// micro-benchmark: latency of the oscillator recurrence ph *= d as a dependent chain, one warp per SM partition
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o chain chain.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 cmul_s(float2 a, float2 b)
{
    float2 c;
    c.x = __fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    c.y = __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
    return c;
}
__global__ void k_scalar(float2 *out, long long *cyc, float2 d, int n)
{
    float2 ph = make_float2(1.0f, 0.0f);
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; i++) ph = cmul_s(ph, d);
    long long t1 = clock64();
    out[threadIdx.x] = ph;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_packed(float2 *out, long long *cyc, float2 d, int n)
{
    float2 ph = make_float2(1.0f, 0.0f);
    const float2 dxx = make_float2(d.x, d.x), dyn = make_float2(-d.y, d.y);
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; i++) {
        const float2 m1 = __fmul2_rn(ph, dxx);                       // (ph.x d.x, ph.y d.x)
        const float2 m2 = __fmul2_rn(make_float2(ph.y, ph.x), dyn);  // (-ph.y d.y, ph.x d.y)
        ph = __fadd2_rn(m1, m2);
    }
    long long t1 = clock64();
    out[threadIdx.x] = ph;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fadd(float *out, long long *cyc, float d, int n)
{
    float a = 1.0f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) a = __fadd_rn(a, d);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fmul_fadd(float *out, long long *cyc, float d, int n)
{
    float a = 1.0f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) a = __fadd_rn(__fmul_rn(a, d), d);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    float2 *o2; float *o1; long long *c, h;
    cudaMalloc(&o2, 1024 * 8); cudaMalloc(&o1, 1024 * 4); cudaMalloc(&c, 8);
    const int n = 1 << 16;
    float2 d = make_float2(0.99f, 0.14106736f), r1, r2;
    for (int rep = 0; rep < 2; rep++) {
        k_scalar<<<1, 32>>>(o2, c, d, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r1, o2, 8, cudaMemcpyDeviceToHost);
        printf("scalar cmul chain: %.2f cycles/step\n", (double)h / n);
        k_packed<<<1, 32>>>(o2, c, d, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r2, o2, 8, cudaMemcpyDeviceToHost);
        printf("packed cmul chain: %.2f cycles/step   same bits: %d\n", (double)h / n, r1.x == r2.x && r1.y == r2.y);
        k_fadd<<<1, 32>>>(o1, c, 1e-3f, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("fadd chain: %.2f cycles/op\n", (double)h / n);
        k_fmul_fadd<<<1, 32>>>(o1, c, 0.999f, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("fmul->fadd chain: %.2f cycles/pair\n", (double)h / n);
    }
    return 0;
}
