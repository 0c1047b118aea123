// micro-benchmark 2: what paces the oscillator recurrence ph *= d inside wb_fsk_kernel, where every lane has its own d in
// REGISTERS (chain.cu passes d as a kernel parameter, i.e. a constant-bank operand): scalar / packed, 32 / 16 active lanes,
// one / two independent chains per lane.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o chain2 chain2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 cmul_s(float2 a, float2 b)
{
    float2 c;
    c.x = __fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    c.y = __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
    return c;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 cmul_p(float2 a, float2 b)
{
    unsigned long long pa = pack2(a.x, a.y), A, B;
    float p0, p1, q0, q1;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(A) : "l"(pa), "l"(pack2(b.x, b.y)));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(B) : "l"(pa), "l"(pack2(b.y, b.x)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(A));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(B));
    return make_float2(__fsub_rn(p0, p1), __fadd_rn(q0, q1));
}
template <int MODE, int NCH>
__global__ void k(float2 *out, long long *cyc, const float2 *dtab, int n, int nact)
{
    if ((int)threadIdx.x >= nact) return;
    float2 d[NCH], ph[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) { d[c] = dtab[(threadIdx.x + 7 * c) & 31]; ph[c] = make_float2(1.0f, 0.0f); }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i += 32) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
#pragma unroll
            for (int c = 0; c < NCH; c++) ph[c] = MODE ? cmul_p(ph[c], d[c]) : cmul_s(ph[c], d[c]);
        }
    }
    long long t1 = clock64();
    float2 acc = ph[0];
#pragma unroll
    for (int c = 1; c < NCH; c++) { acc.x += ph[c].x; acc.y += ph[c].y; }
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// the same chain in warps 0..nchain-1 of a bigger CTA whose other warps wait at the CTA barrier (as in phase B1)
template <int MODE>
__global__ void kb(float2 *out, long long *cyc, const float2 *dtab, int n, int nchain)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 d = dtab[lane], ph = make_float2(1.0f, 0.0f);
    long long t0 = 0, t1 = 0;
    if (warp < nchain) {
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < n; i += 32) {
#pragma unroll
            for (int j = 0; j < 32; j++) ph = MODE ? cmul_p(ph, d) : cmul_s(ph, d);
        }
        t1 = clock64();
    }
    __syncthreads();
    out[threadIdx.x] = ph;
    if (lane == 0 && warp == nchain - 1) cyc[0] = t1 - t0;
}
int main()
{
    float2 *o2, *dt, hd[32]; long long *c, h;
    cudaMalloc(&o2, 1024 * 8); cudaMalloc(&dt, 32 * 8); cudaMalloc(&c, 8);
    for (int i = 0; i < 32; i++) { float a = 0.05f + 0.01f * i; hd[i] = make_float2(cosf(a), sinf(a)); }
    cudaMemcpy(dt, hd, sizeof(hd), cudaMemcpyHostToDevice);
    const int n = 1 << 16;
#define RUN(MODE, NCH, NACT, NAME) k<MODE, NCH><<<1, 32>>>(o2, c, dt, n, NACT); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-44s %6.2f cycles/step\n", NAME, (double)h / n);
    for (int rep = 0; rep < 2; rep++) {
        RUN(0, 1, 32, "scalar, d in registers, 32 lanes");
        RUN(1, 1, 32, "packed, d in registers, 32 lanes");
        RUN(0, 1, 16, "scalar, 16 lanes");
        RUN(1, 1, 16, "packed, 16 lanes");
        RUN(0, 1, 1, "scalar, 1 lane");
#define RUNB(MODE, NT, NCHAIN, NAME) kb<MODE><<<1, NT>>>(o2, c, dt, n, NCHAIN); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-44s %6.2f cycles/step\n", NAME, (double)h / n);
        RUNB(0, 128, 4, "scalar, 4 chain warps (one per partition)");
        RUNB(1, 128, 4, "packed, 4 chain warps");
        RUNB(0, 448, 4, "scalar, 4 chain warps + 10 at the barrier");
        RUNB(1, 448, 4, "packed, 4 chain warps + 10 at the barrier");
        RUNB(0, 448, 8, "scalar, 8 chain warps + 6 at the barrier");
        RUNB(1, 448, 8, "packed, 8 chain warps + 6 at the barrier");
        RUN(0, 2, 32, "scalar, two chains per lane (per step of both)");
        RUN(1, 2, 32, "packed, two chains per lane (per step of both)");
    }
    return 0;
}
