// peak.cu -- FP64 pipe peak (DFMA-saturating) and L2 gather bandwidth microbenchmarks, used only by bench.py to obtain
// the measured roofline denominators that MEASURED_PEAKS.json does not carry.
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
// A warp gathers RIF pseudo-random 1104-byte "rows" at a time (consecutive lanes read consecutive 16-byte pieces of a row:
// the access pattern of the QP kernels' neighbour-row gather), all loads of the batch issued before any is consumed.  The
// buffer (22 MB) is far larger than an SM's L1 and far smaller than L2, and rows are drawn at random: what a gather with no
// reuse inside the SM can get from the L2 fabric.  NC = 1: ld.global.nc (the QP kernels' __ldg), 0: ld.global.cg.
template <int RIF, int NC>
__global__ void __launch_bounds__(256) l2_read_kernel(const double2 *__restrict__ buf, int64_t nrows, int row_vec, int iters, double *out)
{
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double2 acc = make_double2(0.0, 0.0);
    uint64_t r = (uint64_t)gw * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        const double2 *row[RIF];
#pragma unroll
        for (int u = 0; u < RIF; ++u) {
            r = r * 6364136223846793005ull + 1442695040888963407ull;
            row[u] = buf + (int64_t)((r >> 20) % (uint64_t)nrows) * row_vec;
        }
        double2 v[RIF][3]; // row_vec <= 96: at most three 16-byte pieces per lane and row
#pragma unroll
        for (int u = 0; u < RIF; ++u)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int e = lane + 32 * q;
                v[u][q] = e < row_vec ? (NC ? __ldg(row[u] + e) : __ldcg(row[u] + e)) : make_double2(0.0, 0.0);
            }
#pragma unroll
        for (int u = 0; u < RIF; ++u)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                acc.x += v[u][q].x;
                acc.y += v[u][q].y;
            }
    }
    if (acc.x + acc.y == 123.456) out[0] = acc.x;
}

template <int RIF, int NC>
double run_l2_variant(chb_ctx *c, const double2 *buf, int64_t nrows, int row_vec, double *d, cudaEvent_t e0, cudaEvent_t e1)
{
    const int iters = 512 / RIF, blocks = c->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, c->stream);
        l2_read_kernel<RIF, NC><<<blocks, 256, 0, c->stream>>>(buf, nrows, row_vec, iters, d);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 16.0 * row_vec * (double)RIF * (double)iters * (256.0 / 32.0) * blocks;
        if (rep > 0) best = fmax(best, bytes / (ms * 1e-3) / 1e9);
    }
    return best;
}
} // namespace

extern "C" int chb_measure_l2_gbs(chb_ctx *c, double *gbs)
{
    CHB_CHECK(c, c && gbs, CHB_EINVAL, "NULL argument");
    CHB_CUDA(c, cudaSetDevice(c->device));
    const int row_vec = 69;          // 69 x 16 B = 1104 B: one 138-double feature row (d = 137, 16-byte pitch)
    const int64_t nrows = 20000;     // 22 MB: the 20k-contig feature matrix, far inside the 126 MB L2
    double2 *buf = nullptr;
    double *d = nullptr;
    CHB_CUDA(c, cudaMalloc(&buf, sizeof(double2) * (size_t)nrows * row_vec));
    CHB_CUDA(c, cudaMalloc(&d, 8));
    CHB_CUDA(c, cudaMemsetAsync(buf, 0, sizeof(double2) * (size_t)nrows * row_vec, c->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // the best of a few shapes of the same gather (batch depth, load flavour): the ceiling is what the fabric can deliver
    double best = 0.0;
    best = fmax(best, run_l2_variant<2, 1>(c, buf, nrows, row_vec, d, e0, e1));
    best = fmax(best, run_l2_variant<4, 1>(c, buf, nrows, row_vec, d, e0, e1));
    best = fmax(best, run_l2_variant<8, 1>(c, buf, nrows, row_vec, d, e0, e1));
    best = fmax(best, run_l2_variant<4, 0>(c, buf, nrows, row_vec, d, e0, e1));
    best = fmax(best, run_l2_variant<8, 0>(c, buf, nrows, row_vec, d, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(d);
    CHB_CUDA(c, cudaGetLastError());
    *gbs = best;
    return CHB_OK;
}

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
