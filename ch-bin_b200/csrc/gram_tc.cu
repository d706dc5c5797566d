// gram_tc.cu -- candidate squared distances on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as approx.cu (FP32 value A with a proven bound |A - d^2| <= E, consumed by knn.cu), but the Gram
// contraction x_r . x_i runs as tcgen05.mma kind::tf32 with FP32 accumulators in tensor memory, operands brought in
// by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a 3-stage mbarrier pipeline (two CTAs per SM).
//
// Precision: TF32 keeps 10 mantissa bits, which alone would make the filter slack ~3 % of an in-bin distance.
// Each FP32 feature is therefore split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared, lo = x - hi,
// itself truncated to TF32) and the kernel contracts  [hi | hi | lo] . [hi | lo | hi]  along a 3x longer K:
//     hi.hi + hi.lo + lo.hi  =  x.y - lo.lo - (truncation of lo)      -> relative error <= 4 * 2^-20 of |x||y|
// plus the FP32 accumulation error of the tensor core over 3d terms.  Bound used by the filter:
//     | A - |x_r - x_i|^2 |  <=  (3d + 64) * 2^-23 * (nrm[r] + nrm_max)         (checked in tests/test_gpu_parity.py;
//     worst case of round-toward-zero FP32 accumulation over 3d terms; observed maximum ~0.1 of the bound)
//
// Tile: 128 (queries) x 128 (points) per CTA, K step 32 floats (one 128-byte swizzle atom), UMMA 128x128x8.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 2-5 epilogue
// (tcgen05.ld -> nrm[r] + nrm[i] - 2 acc -> 16-byte global stores).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32; // BK floats = 128 bytes
constexpr int STAGES = 3;
constexpr int UMMA_K = 8; // tf32: 32 bytes per MMA k-step
constexpr uint32_t TILE_BYTES = BM * BK * 4; // 16 KB per operand and stage
constexpr int GEMM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 128;

struct SharedStorage {
    alignas(1024) uint8_t a[STAGES][TILE_BYTES];
    alignas(1024) uint8_t b[STAGES][TILE_BYTES];
    alignas(8) uint64_t full_bar[STAGES];
    alignas(8) uint64_t empty_bar[STAGES];
    alignas(8) uint64_t tmem_full_bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    // K-major, SWIZZLE_128B: 8-row groups of 1024 bytes; descriptor version 1 (sm_100)
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);       // start address
    desc |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    desc |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset between 8-row groups
    desc |= (uint64_t)1 << 46;                             // version
    desc |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return desc;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(GEMM_THREADS, 2)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int num_kb,
               const float *__restrict__ nrm, const int32_t *__restrict__ rows, int64_t nrows, int64_t n,
               float *__restrict__ out, int64_t ldo)
{
    extern __shared__ uint8_t smem_raw[];
    SharedStorage &S = *reinterpret_cast<SharedStorage *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&S.full_bar[s], 1);
            mbar_init(&S.empty_bar[s], 1);
        }
        mbar_init(&S.tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 2) { // TMEM allocation is warp-wide
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&S.empty_bar[s], ph ^ 1);
                mbar_expect_tx(&S.full_bar[s], 2 * TILE_BYTES);
                tma_load_2d(S.a[s], &map_a, &S.full_bar[s], kb * BK, m0);
                tma_load_2d(S.b[s], &map_b, &S.full_bar[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&S.full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = make_smem_desc(smem_u32(S.a[s]));
                const uint64_t bdesc = make_smem_desc(smem_u32(S.b[s]));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4); // 32 bytes per k-step inside the swizzle atom
                    umma_tf32(tmem_base, adesc + koff, bdesc + koff, idesc, (kb | k) != 0);
                }
                umma_commit(&S.empty_bar[s]); // frees the stage when these MMAs have read it
            }
            umma_commit(&S.tmem_full_bar);
        }
    } else {
        // epilogue warps 2..5: TMEM lane quadrant = warp % 4
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int64_t gr = (int64_t)m0 + row;
        mbar_wait(&S.tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float nr = (gr < nrows) ? nrm[rows[gr]] : 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (gr < nrows) {
                const int64_t gc0 = (int64_t)n0 + c0;
                float *o = out + gr * ldo + gc0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float r4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t c = gc0 + j + u;
                        const float nc = c < n ? __ldg(nrm + c) : 0.f;
                        r4[u] = fmaf(-2.f, __uint_as_float(v[j + u]), nr + nc);
                    }
                    if (gc0 + j + 3 < n) {
                        *reinterpret_cast<float4 *>(o + j) = make_float4(r4[0], r4[1], r4[2], r4[3]); // ldo % 4 == 0
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (gc0 + j + u < n) o[j + u] = r4[u];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// centred FP32 features (Xf) -> TF32 hi/lo split, laid out for the 3-term contraction.  which = 0: [hi | hi | lo] (queries, rows
// gathered through `rows`), which = 1: [hi | lo | hi] (points).  Kp floats per row, zero padded.
__global__ void split_tf32_kernel(const float *__restrict__ Xf, int32_t ldf, int32_t d, const int32_t *__restrict__ rows,
                                  int64_t nrows, int32_t Kp, int which, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * Kp) return;
    const int64_t r = i / Kp;
    const int32_t t = (int32_t)(i - r * Kp);
    float v = 0.f;
    if (t < 3 * d) {
        const int part = t / d, f = t - part * d;
        const int64_t src = rows ? (int64_t)rows[r] : r;
        const float xf = Xf[src * ldf + f]; // fl32(x - mu), written by prep_f32_kernel
        const float hi = __uint_as_float(__float_as_uint(xf) & 0xffffe000u);
        const float lo = __uint_as_float(__float_as_uint(xf - hi) & 0xffffe000u);
        const bool want_lo = which == 0 ? (part == 2) : (part == 1);
        v = want_lo ? lo : hi;
    }
    out[i] = v;
}

typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn_t get_encode_fn()
{
    static encode_fn_t fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_fn_t>(p);
    }
    return fn;
}

int make_map(chb_ctx *ctx, CUtensorMap *map, float *base, int64_t nrows, int32_t Kp)
{
    encode_fn_t enc = get_encode_fn();
    CHB_CHECK(ctx, enc != nullptr, CHB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)nrows};
    const cuuint64_t gstr[1] = {(cuuint64_t)Kp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CHB_CHECK(ctx, r == CUDA_SUCCESS, CHB_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CHB_OK;
}

} // namespace

// Builds the split operands for the owned query rows (gathered through rows_dev) and for all points.
int chb_gram_tc_prepare(chb_ctx *ctx, const int32_t *rows_dev, int64_t nrows, float *a_split, float *b_split, int32_t Kp,
                        bool do_b)
{
    const int64_t na = nrows * Kp, nb = ctx->n * Kp;
    if (na > 0) {
        split_tf32_kernel<<<(unsigned)((na + 255) / 256), 256, 0, ctx->stream>>>(ctx->Xf, ctx->ldf, ctx->d, rows_dev, nrows, Kp, 0,
                                                                                 a_split);
        ++ctx->tm.launches_other;
    }
    if (do_b && nb > 0) {
        split_tf32_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(ctx->Xf, ctx->ldf, ctx->d, nullptr, ctx->n, Kp, 1,
                                                                                 b_split);
        ++ctx->tm.launches_other;
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}

int chb_launch_gram_tc(chb_ctx *ctx, const float *a_split, const float *b_split, int32_t Kp, const int32_t *rows_dev,
                       int64_t nrows, float *out_dev, int64_t ldo)
{
    if (nrows <= 0) return CHB_OK;
    CUtensorMap map_a, map_b;
    int rc = make_map(ctx, &map_a, const_cast<float *>(a_split), nrows, Kp);
    if (rc != CHB_OK) return rc;
    rc = make_map(ctx, &map_b, const_cast<float *>(b_split), ctx->n, Kp);
    if (rc != CHB_OK) return rc;
    const size_t smem = sizeof(SharedStorage) + 1024;
    // a per-device attribute: set per call, a context may live on any device of this process
    CHB_CUDA(ctx, cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((ctx->n + BN - 1) / BN), (unsigned)((nrows + BM - 1) / BM));
    CHB_CHECK(ctx, grid.y <= 65535u, CHB_EINVAL, "gram_tc: too many rows per launch (%lld)", (long long)nrows);
    {
        chb_stage_timer t(ctx, CHB_ST_DISTANCE);
        gram_tc_kernel<<<grid, GEMM_THREADS, smem, ctx->stream>>>(map_a, map_b, Kp / BK, ctx->nrm, rows_dev, nrows, ctx->n, out_dev,
                                                                  ldo);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}
