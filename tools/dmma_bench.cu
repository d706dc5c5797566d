// FP64 tensor-core (DMMA m8n8k4) and DFMA issue rates on this GPU -- the two ceilings of qp_lane.cu's phases.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_bench tools/dmma_bench.cu && /tmp/dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_kernel(double *out, int iters)
{
    double c[ILP][2];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double *out, int iters)
{
    double c[ILP];
    for (int i = 0; i < ILP; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32 > 1024 ? 1024 : warps * 32, blocks = p.multiProcessorCount * (warps * 32 / threads);
        float ms;
        dmma_kernel<8><<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0);
        dmma_kernel<8><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * threads / 32;
        printf("DMMA m8n8k4  %2d warps/SM: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        dfma_kernel<8><<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0);
        dfma_kernel<8><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl2 = 2.0 * 8.0 * iters * (double)blocks * threads;
        printf("DFMA         %2d warps/SM: %.2f TFLOP/s\n", warps, fl2 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
