// peak.cu -- FP64 pipe peak microbenchmark (DFMA-saturating), used only by bench.py to obtain the measured
// FP64 roofline denominator that MEASURED_PEAKS.json does not carry.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s; // never true: keeps the chain alive
}
} // namespace

extern "C" int chb_measure_fp64_tflops(chb_ctx *c, double *tflops)
{
    CHB_CHECK(c, c && tflops, CHB_EINVAL, "NULL argument");
    CHB_CUDA(c, cudaSetDevice(c->device));
    double *d = nullptr;
    CHB_CUDA(c, cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 14, blocks = c->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        dfma_kernel<<<blocks, 256, 0, c->stream>>>(d, iters, 1.0);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * iters * 256.0 * blocks;
        if (rep > 0) best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    CHB_CUDA(c, cudaGetLastError());
    *tflops = best;
    return CHB_OK;
}
