// Micro-benchmark: peak FP64 throughput of the DFMA pipe vs DMMA (mma.sync.m8n8k4.f64) on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_peak.cu -o gpurun_out/fp64_peak && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
    double a[16], b = 1.0000001, c = 0.9999999;
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_kernel(double* out, int iters) {
    double d[16][2], a = 1.0000001 + threadIdx.x * 1e-12, b = 0.9999999;
    for (int i = 0; i < 16; ++i) { d[i][0] = i; d[i][1] = -i; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < 16; ++i) s += d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 512 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {256, 512}) for (int bps : {1, 2, 4}) {
        int grid = 148 * bps;
        for (int which = 0; which < 2; ++which) {
            if (which == 0) dfma_kernel<<<grid, threads>>>(out, 100); else dmma_kernel<<<grid, threads>>>(out, 100);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            if (which == 0) dfma_kernel<<<grid, threads>>>(out, iters); else dmma_kernel<<<grid, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = which == 0 ? 2.0 * 16 * iters * (double)grid * threads : 2.0 * 256 * 16 * iters * (double)grid * (threads / 32);
            printf("%s threads %d blocks/SM %d: %.2f TFLOP/s\n", which == 0 ? "DFMA" : "DMMA", threads, bps, flops / ms / 1e9);
        }
    }
    return 0;
}
