// approx.cu -- FP32 candidate distances (squared) for the kNN filter (sm_100a).
//
// The reference ranks neighbours on scipy's FP64 cdist values (distance_matrix.py:26,41,60-62).  Forming every
// one of those exactly costs 3 non-fusable FP64 operations per (query, point, feature) (distance.cu) and was the
// largest kernel of the stage.  Neighbour SETS only depend on the k smallest distances per bin, so the scan
// (knn.cu) filters on a cheap FP32 value with a rigorous error bound and recomputes the exact scipy recipe only
// for the few candidates that can reach a top-k list.  This file produces the FP32 values:
//
//     A[r][i] = fl32( nrm[r] + nrm[i] - 2 * sum_t xf[r][t] * xf[i][t] )          (Gram form, FFMA accumulation)
//
// with xf = fl32(x - mu) (mu = column means: distances are translation invariant and centring shrinks the bound)
// and nrm = |xf|^2 evaluated in FP64 and rounded UP to FP32.  Error bound used by the filter
// (standard gamma_d analysis, inputs rounded once, d FMAs, three final roundings):
//     | A[r][i] - |x_r - x_i|^2 |  <=  (d + 16) * 2^-23 * (nrm[r] + nrm_max)
// Valid while no intermediate overflows or flushes to zero: features must satisfy 1e-15 < |x|_max < 1e15, which
// chb_build_distance_matrix checks (normalised k-mer / coverage features are in [0, 1]).
//
// Roofline: FP32 FMA pipe, 1 FFMA per (query, point, feature); output U*n*4 bytes to HBM.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int PADM = BM + 4;

// column sums of X (for the centring translation), in a fixed summation order so that every rank of a multi-GPU run
// derives bit-identical centred features: block b sums rows [256 b, 256 b + 256) into part[b][:], then one block adds
// the partial sums in block order.
__global__ void __launch_bounds__(256) colsum_part_kernel(const double *__restrict__ X, int32_t ldx, int32_t d, int64_t n,
                                                            double *__restrict__ part)
{
    const int64_t r0 = (int64_t)blockIdx.x * 256;
    const int64_t r1 = r0 + 256 < n ? r0 + 256 : n;
    for (int t = threadIdx.x; t < d; t += 256) {
        double s = 0.0;
        for (int64_t r = r0; r < r1; ++r) s += X[r * ldx + t];
        part[(int64_t)blockIdx.x * d + t] = s;
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const double *__restrict__ part, int64_t nparts, int32_t d,
                                                             double *__restrict__ colsum)
{
    for (int t = threadIdx.x; t < d; t += 256) {
        double s = 0.0;
        for (int64_t b = 0; b < nparts; ++b) s += part[b * d + t];
        colsum[t] = s;
    }
}

// Xf = fl32(x - mu), nrm = |x - mu|^2 rounded up.  Distances are translation invariant, and centring the all-positive
// k-mer profiles shrinks |x|^2 -- and with it the filter's error bound E ~ eps (nrm[r] + nrm_max) -- several times.
__global__ void __launch_bounds__(256) prep_f32_kernel(const double *__restrict__ X, int32_t ldx, int32_t d, int64_t n,
                                                         const double *__restrict__ colsum, double inv_n,
                                                         float *__restrict__ Xf, int32_t ldf, float *__restrict__ nrm,
                                                         unsigned int *__restrict__ nrm_max_bits)
{
    // one warp per point
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    double s = 0.0;
    // the first 160 features with all five loads of a lane in flight (one outstanding 8-byte load per lane kept this pass at
    // 1.5 TB/s of HBM traffic); same per-lane summation order as the plain loop that finishes wider rows
    constexpr int PF_U = 5;
    double xv[PF_U];
#pragma unroll
    for (int u = 0; u < PF_U; ++u) {
        const int t = lane + 32 * u;
        xv[u] = t < d ? X[i * ldx + t] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < PF_U; ++u) {
        const int t = lane + 32 * u;
        if (t < ldf) {
            const double v = t < d ? xv[u] - colsum[t] * inv_n : 0.0;
            const float vf = (float)v;
            Xf[i * ldf + t] = vf;
            s = fma((double)vf, (double)vf, s); // the norm of the value the Gram kernels actually contract
        }
    }
    for (int t = lane + 32 * PF_U; t < ldf; t += 32) {
        const double v = t < d ? X[i * ldx + t] - colsum[t] * inv_n : 0.0;
        const float vf = (float)v;
        Xf[i * ldf + t] = vf;
        s = fma((double)vf, (double)vf, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(CHB_FULL, s, o);
    if (lane == 0) {
        const float f = __double2float_ru(s * (1.0 + 1e-12));
        nrm[i] = f;
        // non-negative floats order like their bit patterns; a maximum only grows, so a plain (L2) read screens out nearly every
        // point before the atomic -- n atomics on ONE word serialised
        if (__float_as_uint(f) > __ldcg(nrm_max_bits)) atomicMax(nrm_max_bits, __float_as_uint(f));
    }
}

__global__ void __launch_bounds__(256) approx_rows_kernel(const float *__restrict__ Xf, int32_t ldf, const float *__restrict__ nrm,
                                                            const int32_t *__restrict__ rows, int64_t nrows, int64_t n,
                                                            float *__restrict__ out, int64_t ldo)
{
    __shared__ __align__(16) float sA[BK][PADM];
    __shared__ __align__(16) float sB[BK][PADM];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.y * BM, col0 = (int64_t)blockIdx.x * BN;
    const int lt = tid & 15, lr = tid >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int t0 = 0; t0 < ldf; t0 += BK) {
#pragma unroll
        for (int pass = 0; pass < BM / 16; ++pass) {
            const int r = pass * 16 + lr;
            const int64_t gr = row0 + r, gc = col0 + r;
            const int t = t0 + lt;
            float va = 0.f, vb = 0.f;
            if (t < ldf) {
                if (gr < nrows) va = Xf[(int64_t)rows[gr] * ldf + t];
                if (gc < n) vb = Xf[gc * ldf + t];
            }
            sA[lt][r] = va;
            sB[lt][r] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&sA[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&sA[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&sB[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&sB[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gr = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gr >= nrows) continue;
        const float nr = nrm[rows[gr]];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t gc = col0 + h * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t c = gc + j;
                const float nc = c < n ? nrm[c] : 0.f;
                v[j] = fmaf(-2.f, acc[i][h * 4 + j], nr + nc);
            }
            float *o = out + gr * ldo + gc;
            if (gc + 3 < n && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gc + j < n) o[j] = v[j];
            }
        }
    }
}

} // namespace

int chb_launch_prep_f32(chb_ctx *ctx)
{
    const int64_t n = ctx->n;
    CHB_CUDA(ctx, cudaMemsetAsync(&ctx->counters[5], 0, sizeof(int32_t), ctx->stream));
    const int64_t nparts = (n + 255) / 256;
    if (ctx->cap_colpart < nparts * ctx->d) {
        if (ctx->colpart) cudaFree(ctx->colpart);
        ctx->colpart = nullptr;
        ctx->cap_colpart = 0;
        CHB_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->colpart), sizeof(double) * (size_t)(nparts * ctx->d)));
        ctx->cap_colpart = nparts * ctx->d;
    }
    colsum_part_kernel<<<(unsigned)nparts, 256, 0, ctx->stream>>>(ctx->X, ctx->ldx, ctx->d, n, ctx->colpart);
    colsum_final_kernel<<<1, 256, 0, ctx->stream>>>(ctx->colpart, nparts, ctx->d, ctx->colsum);
    prep_f32_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->X, ctx->ldx, ctx->d, n, ctx->colsum, 1.0 / (double)n, ctx->Xf, ctx->ldf, ctx->nrm,
        reinterpret_cast<unsigned int *>(&ctx->counters[5]));
    CHB_CUDA(ctx, cudaGetLastError());
    ctx->tm.launches_other += 3;
    return CHB_OK;
}

int chb_launch_approx_rows(chb_ctx *ctx, const int32_t *rows_dev, int64_t nrows, float *out_dev, int64_t ldo)
{
    if (nrows <= 0) return CHB_OK;
    dim3 grid((unsigned)((ctx->n + BN - 1) / BN), (unsigned)((nrows + BM - 1) / BM));
    CHB_CHECK(ctx, grid.y <= 65535u, CHB_EINVAL, "approx rows: too many rows per launch (%lld)", (long long)nrows);
    {
        chb_stage_timer t(ctx, CHB_ST_DISTANCE);
        approx_rows_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->Xf, ctx->ldf, ctx->nrm, rows_dev, nrows, ctx->n, out_dev, ldo);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}
