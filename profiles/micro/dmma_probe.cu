// FP64 tensor-core probe: mma.sync.aligned.m8n8k4.row.col.f64 throughput on sm_100a (is DMMA usable for the chain rule?)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dmma_kernel(long iters, double *out)
{
    double a = threadIdx.x * 1e-3 + 0.5, b = 0.999 + threadIdx.x * 1e-6;
    double c0[2] = {0.1, 0.2}, c1[2] = {0.3, 0.4}, c2[2] = {0.5, 0.6}, c3[2] = {0.7, 0.8};
    for (long i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[(long)blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

int main()
{
    const int blocks = 148 * 8;
    const long iters = 100000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dmma_kernel<<<blocks, 256>>>(1000, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        dmma_kernel<<<blocks, 256>>>(iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // per warp per iteration: 4 mma x (8*8*4 FMA = 512 flop)
    const double flops = (double)blocks * 8 * iters * 4 * 512.0;
    printf("DMMA m8n8k4: %.3f ms, %.2f TFLOP/s (err=%s)\n", best, flops / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
