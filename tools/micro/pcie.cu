// micro-benchmark: pinned host -> device copy rates, one flat copy vs the engine's strided 2D copy (one row per stream)
// nvcc -O2 -o pcie pcie.cu
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t rows = 2048, row_bytes = 1 << 20, dpitch = row_bytes + 512 * 2 + 256;   // engine rows: headroom + chunk + slack
    unsigned char *h, *d;
    cudaHostAlloc(&h, rows * row_bytes, cudaHostAllocDefault);
    cudaMalloc(&d, rows * dpitch);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e9f;
        for (int it = 0; it < 6; it++) {
            cudaEventRecord(e0, s);
            if (mode == 0) cudaMemcpyAsync(d, h, rows * row_bytes, cudaMemcpyHostToDevice, s);
            else if (mode == 1) cudaMemcpy2DAsync(d + 1024, dpitch, h, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s);
            else for (size_t r = 0; r < rows; r += 256) cudaMemcpy2DAsync(d + 1024 + r * dpitch, dpitch, h + r * row_bytes, row_bytes, row_bytes, 256, cudaMemcpyHostToDevice, s);
            cudaEventRecord(e1, s); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s: %.2f GB/s\n", mode == 0 ? "flat 2 GiB" : mode == 1 ? "2D 2048 rows x 1 MiB (engine layout)" : "2D in 8 pieces", rows * row_bytes / best / 1e6);
    }
    return 0;
}
