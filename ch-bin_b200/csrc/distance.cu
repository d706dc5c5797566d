// distance.cu -- exact FP64 Euclidean distance rows (sm_100a).
//
// Replaces scipy.spatial.distance.cdist(arr, arr, 'euclidean') as called by
// /root/reference/ch_bin/core/clustering/distance_matrix.py:26,41.  The reference's neighbour ranking is done on
// these very doubles, so the kernel reproduces scipy's arithmetic bit for bit:
//     s = 0;  for t in 0..d-1:  diff = a[t] - b[t];  s = s + diff*diff   (separate multiply and add, no FMA)
//     D = sqrt(s)                                                         (IEEE round-to-nearest)
// (tests/test_oracle_golden.py pins that recipe against scipy itself; tests/test_gpu_parity.py::test_distance_rows_bit_exact
// pins this kernel against the oracle.)  The sum MUST run in ascending t per output, so the contraction is register-tiled over
// outputs (4x4 per thread) and streamed over t -- it cannot be re-associated into tensor-core MMAs.
//
// Roofline: FP64 pipe.  3 dependent-free FP64 instructions per (row, column, t): DADD(sub), DMUL, DADD.
#include "common.cuh"

namespace {

constexpr int TM = 64;   // query rows per CTA
constexpr int TN = 64;   // columns (points) per CTA
constexpr int TK = 16;   // feature chunk staged in shared memory
constexpr int LDS_ = TM + 2; // padded, keeps 16-byte alignment of 4-double groups

__global__ void __launch_bounds__(256) distance_rows_kernel(const double *__restrict__ X, int32_t ldx, int32_t d,
                                                              const int32_t *__restrict__ rows, int64_t nrows, int64_t n,
                                                              double *__restrict__ out)
{
    __shared__ __align__(16) double sA[TK][LDS_];
    __shared__ __align__(16) double sB[TK][LDS_];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.y * TM;
    const int64_t col0 = (int64_t)blockIdx.x * TN;

    // loader mapping: 16 consecutive threads read 16 consecutive features of one row (128 B), 16 rows per pass
    const int lt = tid & 15, lr = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int t0 = 0; t0 < d; t0 += TK) {
#pragma unroll
        for (int pass = 0; pass < TM / 16; ++pass) {
            int r = pass * 16 + lr;
            int64_t gr = row0 + r;
            double va = 0.0, vb = 0.0;
            int t = t0 + lt;
            if (t < d) {
                if (gr < nrows) va = X[(int64_t)rows[gr] * ldx + t];
                int64_t gc = col0 + r;
                if (gc < n) vb = X[gc * ldx + t];
            }
            sA[lt][r] = va;
            sB[lt][r] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < TK; ++tt) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&sA[tt][ty * 4]);
            const double2 a23 = *reinterpret_cast<const double2 *>(&sA[tt][ty * 4 + 2]);
            const double2 b01 = *reinterpret_cast<const double2 *>(&sB[tt][tx * 4]);
            const double2 b23 = *reinterpret_cast<const double2 *>(&sB[tt][tx * 4 + 2]);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y};
            const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double df = __dsub_rn(a[i], b[j]);
                    acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(df, df)); // never an FMA
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t gr = row0 + ty * 4 + i;
        if (gr >= nrows) continue;
        int64_t gc = col0 + tx * 4;
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __dsqrt_rn(acc[i][j]);
        double *o = out + gr * n + gc;
        if (gc + 3 < n && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            reinterpret_cast<double2 *>(o)[0] = make_double2(v[0], v[1]);
            reinterpret_cast<double2 *>(o)[1] = make_double2(v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gc + j < n) o[j] = v[j];
        }
    }
}

} // namespace

int chb_launch_distance_rows(chb_ctx *ctx, const int32_t *rows_dev, int64_t nrows, double *out_dev)
{
    if (nrows <= 0) return CHB_OK;
    dim3 grid((unsigned)((ctx->n + TN - 1) / TN), (unsigned)((nrows + TM - 1) / TM));
    CHB_CHECK(ctx, grid.y <= 65535u, CHB_EINVAL, "distance rows: too many rows per launch (%lld)", (long long)nrows);
    {
        chb_stage_timer t(ctx, CHB_ST_DISTANCE);
        distance_rows_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->X, ctx->ldx, ctx->d, rows_dev, nrows, ctx->n, out_dev);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}
