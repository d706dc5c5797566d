// fused.cu -- distance mode 2: tensor-core Gram contraction FUSED with the per-bin neighbour selection (sm_100a).
//
// Replaces, for one speculate/repair round, create_in_mem_distance_matrix + find_nearest_from_cluster
// (/root/reference/ch_bin/core/clustering/distance_matrix.py:33-62) for every (query, bin) pair, without ever
// writing a distance matrix:
//
//  1. "column entries": every point becomes a column of the bin(s) in which some query can see it this round.
//     The label the query at permutation position p sees for point i is  pos[i] < p ? tent[i] : old[i]
//     (algorithm.py:46-60), so point i is a column of bin tent[i] visible when p > pos[i], and -- if different --
//     of bin old[i] visible when p < pos[i]; when both labels agree it is one column visible when p != pos[i]
//     (which also removes the query itself, algorithm.py:50).  Columns are counting-sorted by bin, each bin padded
//     to a multiple of 128, and the TF32 hi/lo operand rows are gathered in that order.
//  2. gram_select_kernel: persistent CTA per 128-query row block, looping over the 128-column tiles (= one bin
//     each).  TMA -> 3-stage smem ring -> tcgen05.mma kind::tf32 (3-term split, see gram_tc.cu) -> FP32
//     accumulators double-buffered in TMEM.  While the tensor core works on tile t+1, the four epilogue warps read
//     tile t from TMEM: thread r owns query row r, forms A = nrm[r] + nrm[c] - 2 acc for its 128 columns, masks
//     (two threads per row, one per 64-column half) by visibility, and keeps the KR smallest (A, point) pairs of the
//     current bin in REGISTERS
//     (branch-free swap insertion; candidates are found with a 32-column bitmask and the warp loops only
//     max-over-lanes popcount times).  At a bin boundary the list is flushed to global memory.
//  3. rerank_kernel: per (query, bin): with |A - d^2| <= E, the k smallest-A candidates are EXACTLY the reference's
//     neighbours whenever A_(k+1) > A_(k) + 2E; otherwise only the ambiguous candidates (A within 2E of the
//     boundary) get scipy's exact recipe (sequential sum, no FMA) and are ranked by exact (distance, index).
//     A pair goes to the QP work list only if its neighbour SET changed; if the kept list could have missed a
//     candidate (more than KR keys inside the slack: duplicate contigs) the query falls back to knn.cu.
//
// Roofline: tensor pipe / L2->SM operand traffic (both operands stream: 2*128*Kp*4 bytes per 128x128 tile).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int UMMA_K = 8;
constexpr uint32_t TILE_BYTES = BM * BK * 4;
constexpr int FUSED_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int EPI_THREADS = 256;
constexpr uint32_t TMEM_COLS = 256; // two 128-column accumulator buffers

struct ColMeta { // per-tile column metadata staged in shared memory
    float nrm[BN];
    int pt[BN];
    int a[BN]; // visible to the query at position p iff p > a || p < b
    int b[BN];
};

struct SharedStorage {
    alignas(1024) uint8_t a[STAGES][TILE_BYTES];
    alignas(1024) uint8_t b[STAGES][TILE_BYTES];
    alignas(16) ColMeta meta[2];
    alignas(16) float stage_vals[2][32][BM]; // per column half: one 32-column batch of A values, [column][row]
    alignas(8) uint64_t full_bar[STAGES];
    alignas(8) uint64_t empty_bar[STAGES];
    alignas(8) uint64_t tmem_full_bar[2];
    alignas(8) uint64_t tmem_empty_bar[2];
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    desc |= (uint64_t)1 << 16;
    desc |= (uint64_t)(1024u >> 4) << 32;
    desc |= (uint64_t)1 << 46;
    desc |= (uint64_t)2 << 61;
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// column entries
// ---------------------------------------------------------------------------------------------------------
__global__ void entries_count_kernel(const int32_t *__restrict__ tent, const int32_t *__restrict__ old, int64_t n, int32_t C,
                                     int32_t *__restrict__ bin_cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = tent[i], o = old[i];
    if (t >= 0 && t < C) atomicAdd(&bin_cnt[t], 1);
    if (o >= 0 && o < C && o != t) atomicAdd(&bin_cnt[o], 1);
}

// one block: segment offsets (each bin padded to a multiple of BN), tile -> bin map, cursors reset
__global__ void entries_scan_kernel(const int32_t *__restrict__ bin_cnt, int32_t C, int32_t *__restrict__ seg_off,
                                    int32_t *__restrict__ cursor, int32_t *__restrict__ tile_bin, int32_t *__restrict__ ntiles_out)
{
    __shared__ int total;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int c = 0; c < C; ++c) {
            seg_off[c] = off;
            off += (bin_cnt[c] + BN - 1) / BN * BN;
        }
        seg_off[C] = off;
        total = off;
        *ntiles_out = off / BN;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        cursor[c] = 0;
        for (int t = seg_off[c] / BN; t < seg_off[c + 1] / BN; ++t) tile_bin[t] = c;
    }
    (void)total;
}

__global__ void entries_scatter_kernel(const int32_t *__restrict__ tent, const int32_t *__restrict__ old,
                                       const int32_t *__restrict__ pos, int64_t n, int32_t C, const int32_t *__restrict__ seg_off,
                                       int32_t *__restrict__ cursor, int32_t *__restrict__ col_pt, int32_t *__restrict__ col_a,
                                       int32_t *__restrict__ col_b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = tent[i], o = old[i], ps = pos[i];
    const bool tv = t >= 0 && t < C, ov = o >= 0 && o < C;
    if (tv) {
        const int e = seg_off[t] + atomicAdd(&cursor[t], 1);
        col_pt[e] = (int)i;
        col_a[e] = ps;                      // visible when p > pos[i] ...
        col_b[e] = (ov && o == t) ? ps : INT32_MIN; // ... and also when p < pos[i] if the old label agrees
    }
    if (ov && o != t) {
        const int e = seg_off[o] + atomicAdd(&cursor[o], 1);
        col_pt[e] = (int)i;
        col_a[e] = INT32_MAX;
        col_b[e] = ps;                      // visible only when p < pos[i]
    }
}

__global__ void entries_fill_kernel(int64_t ncol, int32_t *__restrict__ col_pt, int32_t *__restrict__ col_a,
                                    int32_t *__restrict__ col_b)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ncol) return;
    col_pt[e] = -1;
    col_a[e] = INT32_MAX; // never visible
    col_b[e] = INT32_MIN;
}

// gathers the TF32-split operand rows and norms into column order; padding columns are zero
__global__ void entries_gather_kernel(const int32_t *__restrict__ col_pt, const int32_t *__restrict__ ntiles, const float *__restrict__ bsplit,
                                      const float *__restrict__ nrm, int32_t Kp, float *__restrict__ bperm, float *__restrict__ col_nrm)
{
    const int64_t ncol = (int64_t)(*ntiles) * BN;
    const int64_t kq = Kp / 4; // float4 per row
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncol * kq) return;
    const int64_t e = i / kq;
    const int q = (int)(i - e * kq);
    const int ptx = col_pt[e];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ptx >= 0) v = __ldg(reinterpret_cast<const float4 *>(bsplit + (int64_t)ptx * Kp) + q);
    reinterpret_cast<float4 *>(bperm + e * Kp)[q] = v;
    if (q == 0) col_nrm[e] = ptx >= 0 ? nrm[ptx] : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// fused Gram + selection
// ---------------------------------------------------------------------------------------------------------
template <int KR>
struct TopList {
    float key[KR];
    int idx[KR];
    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int s = 0; s < KR; ++s) { key[s] = INFINITY; idx[s] = INT32_MAX; }
    }
    // keeps the KR smallest (key, idx) pairs in ascending order; branch-free
    __device__ __forceinline__ void insert(float ka, int ki)
    {
#pragma unroll
        for (int s = 0; s < KR; ++s) {
            const bool lt = (ka < key[s]) || (ka == key[s] && ki < idx[s]);
            const float tk = lt ? key[s] : ka;
            const int ti = lt ? idx[s] : ki;
            key[s] = lt ? ka : key[s];
            idx[s] = lt ? ki : idx[s];
            ka = tk;
            ki = ti;
        }
    }
};

template <int KR>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
gram_select_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int num_kb,
                   const int32_t *__restrict__ ntiles_p, const int32_t *__restrict__ tile_bin, const int32_t *__restrict__ col_pt,
                   const int32_t *__restrict__ col_a, const int32_t *__restrict__ col_b, const float *__restrict__ col_nrm,
                   const float *__restrict__ nrm, const int32_t *__restrict__ row_point, const int32_t *__restrict__ pos,
                   int64_t nrows, int32_t C, float *__restrict__ cand_key, int32_t *__restrict__ cand_idx)
{
    extern __shared__ uint8_t smem_raw[];
    SharedStorage &S = *reinterpret_cast<SharedStorage *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = *ntiles_p;
    const int nrb = (int)((nrows + BM - 1) / BM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&S.full_bar[s], 1);
            mbar_init(&S.empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&S.tmem_full_bar[b], 1);
            mbar_init(&S.tmem_empty_bar[b], 8); // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int64_t it = 0;
            for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
                for (int t = 0; t < ntiles; ++t) {
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int s = (int)(it % STAGES);
                        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                        mbar_wait(&S.empty_bar[s], ph ^ 1);
                        mbar_expect_tx(&S.full_bar[s], 2 * TILE_BYTES);
                        tma_load_2d(S.a[s], &map_a, &S.full_bar[s], kb * BK, rb * BM);
                        tma_load_2d(S.b[s], &map_b, &S.full_bar[s], kb * BK, t * BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int64_t it = 0, tt = 0;
            for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const int buf = (int)(tt & 1);
                    const uint32_t tph = (uint32_t)((tt >> 1) & 1);
                    mbar_wait(&S.tmem_empty_bar[buf], tph ^ 1); // epilogue has drained this accumulator buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int s = (int)(it % STAGES);
                        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                        mbar_wait(&S.full_bar[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t adesc = make_smem_desc(smem_u32(S.a[s]));
                        const uint64_t bdesc = make_smem_desc(smem_u32(S.b[s]));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
                            umma_tf32(tmem_d, adesc + koff, bdesc + koff, idesc, (kb | k) != 0);
                        }
                        umma_commit(&S.empty_bar[s]);
                    }
                    umma_commit(&S.tmem_full_bar[buf]);
                }
            }
        }
    } else {
        // ---------------- epilogue warps 2..9: thread <-> (query row, column half)
        const int q = warp & 3;                 // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;       // 0: columns [0,64) of a tile, 1: columns [64,128)
        const int row = q * 32 + lane;          // TMEM lane == row within the block
        const int et = threadIdx.x - 64;        // 0..255 index among the epilogue threads
        TopList<KR> L;
        int64_t tt = 0;
        for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            const int64_t gr = (int64_t)rb * BM + row;
            const bool rvalid = gr < nrows;
            int p = INT32_MIN + 1;
            float nr = 0.f;
            if (rvalid) {
                const int pt = row_point[gr];
                p = pos[pt];
                nr = nrm[pt];
            }
            L.reset();
            int cur_bin = ntiles > 0 ? tile_bin[0] : -1;
            for (int t = 0; t < ntiles; ++t, ++tt) {
                const int buf = (int)(tt & 1);
                const uint32_t tph = (uint32_t)((tt >> 1) & 1);
                const int tb = tile_bin[t];
                if (tb != cur_bin) { // bin boundary: flush the finished (half-)list
                    if (rvalid) {
                        float *ok = cand_key + ((gr * C + cur_bin) * 2 + half) * KR;
                        int32_t *oi = cand_idx + ((gr * C + cur_bin) * 2 + half) * KR;
#pragma unroll
                        for (int s = 0; s < KR; ++s) { ok[s] = L.key[s]; oi[s] = L.idx[s]; }
                    }
                    L.reset();
                    cur_bin = tb;
                }
                if (et < BN) { // stage this tile's column metadata
                    const int64_t e = (int64_t)t * BN + et;
                    ColMeta &Mw = S.meta[buf];
                    Mw.nrm[et] = col_nrm[e];
                    Mw.pt[et] = col_pt[e];
                    Mw.a[et] = col_a[e];
                    Mw.b[et] = col_b[e];
                }
                epi_bar_sync();
                const ColMeta &M = S.meta[buf];
                mbar_wait(&S.tmem_full_bar[buf], tph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int cb = 0; cb < 2; ++cb) {
                    const int c0 = half * 64 + cb * 32;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (cb == 1) {
                        // this warp has read its part of the accumulator tile: hand the buffer back to the MMA warp
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.tmem_empty_bar[buf]);
                    }
                    const float thr = L.key[KR - 1];
                    unsigned mask = 0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 n4 = *reinterpret_cast<const float4 *>(&M.nrm[c0 + j]);
                        const int4 a4 = *reinterpret_cast<const int4 *>(&M.a[c0 + j]);
                        const int4 b4 = *reinterpret_cast<const int4 *>(&M.b[c0 + j]);
                        const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
                        const int aa[4] = {a4.x, a4.y, a4.z, a4.w};
                        const int bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float av = fmaf(-2.f, __uint_as_float(v[j + u]), nr + nn[u]);
                            const bool cand = ((p > aa[u]) || (p < bb[u])) && (av <= thr);
                            if (cand) {
                                S.stage_vals[half][j + u][row] = av;
                                mask |= 1u << (j + u);
                            }
                        }
                    }
                    // insert candidates: the warp iterates max-over-lanes popcount times
                    while (__any_sync(CHB_FULL, mask != 0)) {
                        if (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const float av = S.stage_vals[half][j][row];
                            if (av <= L.key[KR - 1]) L.insert(av, M.pt[c0 + j]);
                        }
                    }
                }
                epi_bar_sync(); // everyone is done with meta[buf] before it is overwritten two tiles later
            }
            if (rvalid && cur_bin >= 0) {
                float *ok = cand_key + ((gr * C + cur_bin) * 2 + half) * KR;
                int32_t *oi = cand_idx + ((gr * C + cur_bin) * 2 + half) * KR;
#pragma unroll
                for (int s = 0; s < KR; ++s) { ok[s] = L.key[s]; oi[s] = L.idx[s]; }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------------------
// re-rank: candidates -> exact neighbour sets, dirty detection, QP work list
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double exact_distance_g(const double *__restrict__ xq_s, const double *__restrict__ xi, int d)
{
    double acc = 0.0;
    const int d8 = d & ~7;
    double v[8];
    if (d8 > 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = xi[u];
    }
    for (int t = 0; t < d8; t += 8) {
        double sq[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double df = __dsub_rn(xq_s[t + u], v[u]);
            sq[u] = __dmul_rn(df, df);
        }
        if (t + 8 < d8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = xi[t + 8 + u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, sq[u]);
    }
    for (int t = d8; t < d; ++t) {
        const double df = __dsub_rn(xq_s[t], xi[t]);
        acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    return __dsqrt_rn(acc);
}

// One CTA (128 threads) per owned query; warp w handles bins w, w+4, ...; lane s < 2*KR holds one kept candidate
// (two half-lists of KR per pair, one per column half of the fused kernel's tiles).
// Bins that never received a flush (no columns at all) are recognised through bin_cnt.
__global__ void __launch_bounds__(128) rerank_kernel(const float *__restrict__ cand_key, const int32_t *__restrict__ cand_idx, int KR,
                                                     const int32_t *__restrict__ bin_cnt, const double *__restrict__ X, int32_t ldx,
                                                     int32_t d, const int32_t *__restrict__ row_point, const float *__restrict__ nrm,
                                                     const unsigned int *__restrict__ nrm_max_bits, double eps_rel, int32_t C,
                                                     int32_t k, int32_t *__restrict__ knn_idx, int32_t *__restrict__ knn_cnt,
                                                     int2 *__restrict__ work, int32_t *__restrict__ work_count,
                                                     int32_t *__restrict__ fb_rows, int32_t *__restrict__ fb_count)
{
    extern __shared__ __align__(16) double xq_s[];
    const int64_t r = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int jq = row_point[r];
    for (int t = threadIdx.x; t < d; t += 128) xq_s[t] = X[(int64_t)jq * ldx + t];
    __shared__ int s_overflow;
    if (threadIdx.x == 0) s_overflow = 0;
    __syncthreads();
    const double nmax = (double)__uint_as_float(*nrm_max_bits);
    const float slack2 = __double2float_ru(2.0 * (eps_rel * ((double)nrm[jq] + nmax) + 1e-30));
    const int K2 = 2 * KR;

    for (int c = warp; c < C; c += 4) {
        const int64_t pair = r * C + c;
        float ka = INFINITY;
        int ki = INT32_MAX;
        if (bin_cnt[c] > 0 && lane < K2) {
            ka = cand_key[pair * K2 + lane];
            ki = cand_idx[pair * K2 + lane];
        }
        const bool valid = ka < INFINITY; // +inf = empty slot
        const unsigned vm = __ballot_sync(CHB_FULL, valid);
        const int nc = __popc(vm);
        const int m = nc < k ? nc : k;  // |bin| <= k: all members (distance_matrix.py:58-59)
        // rank of every candidate by (key, index)
        int rnk = 0;
        for (int t = 0; t < K2; ++t) {
            const float ok = __shfl_sync(CHB_FULL, ka, t);
            const int oi = __shfl_sync(CHB_FULL, ki, t);
            if (((vm >> t) & 1u) && (ok < ka || (ok == ka && oi < ki))) ++rnk;
        }
        bool sel = false;
        if (nc <= k) {
            sel = valid;
        } else {
            const int src_k = __ffs(__ballot_sync(CHB_FULL, valid && rnk == k - 1)) - 1;
            const int src_k1 = __ffs(__ballot_sync(CHB_FULL, valid && rnk == k)) - 1;
            const float a_k = __shfl_sync(CHB_FULL, ka, src_k);   // k-th smallest key
            const float a_k1 = __shfl_sync(CHB_FULL, ka, src_k1); // (k+1)-th
            const float hi = __fadd_ru(a_k, slack2);
            // a half-list that is full and entirely inside the slack may have dropped a closer point
            const bool h0 = lane < KR, h1 = lane >= KR && lane < K2;
            const unsigned in0 = __ballot_sync(CHB_FULL, h0 && valid && ka <= hi), in1 = __ballot_sync(CHB_FULL, h1 && valid && ka <= hi);
            if (__popc(in0) == KR || __popc(in1) == KR) {
                if (lane == 0) s_overflow = 1;
            }
            const bool could = valid && ka <= hi;                      // may belong to the exact top-k
            const bool sure = valid && (__fadd_ru(ka, slack2) < a_k1); // certainly belongs to it
            const unsigned sm = __ballot_sync(CHB_FULL, sure), cm = __ballot_sync(CHB_FULL, could);
            const int nsure = __popc(sm);
            if (__popc(cm) == k) {
                sel = could; // unambiguous: the k smallest keys ARE the neighbours
            } else {
                // ambiguous candidates: exact scipy-recipe distance, rank by (distance, index)
                const bool amb = could && !sure;
                double de = 0.0;
                if (amb) de = exact_distance_g(xq_s, X + (int64_t)ki * ldx, d);
                const unsigned am = __ballot_sync(CHB_FULL, amb);
                int rk = 0;
                for (int t = 0; t < K2; ++t) {
                    const double od = __shfl_sync(CHB_FULL, de, t);
                    const int oi = __shfl_sync(CHB_FULL, ki, t);
                    if (((am >> t) & 1u) && (od < de || (od == de && oi < ki))) ++rk;
                }
                sel = sure || (amb && rk < k - nsure);
            }
        }
        // neighbour set, stored in ascending index order (canonical for set comparison)
        int srt = 0;
        const unsigned selm = __ballot_sync(CHB_FULL, sel);
        for (int t = 0; t < K2; ++t) {
            const int oi = __shfl_sync(CHB_FULL, ki, t);
            if (((selm >> t) & 1u) && oi < ki) ++srt;
        }
        const int mo = knn_cnt[pair];
        bool same = (mo == m);
        if (same) {
            const bool diff = sel && (knn_idx[pair * k + srt] != ki);
            same = !__any_sync(CHB_FULL, diff);
        }
        if (!same) {
            if (lane < k) knn_idx[pair * k + lane] = -1;
            __syncwarp();
            if (sel) knn_idx[pair * k + srt] = ki;
            if (lane == 0) {
                knn_cnt[pair] = m;
                const int w = atomicAdd(work_count, 1);
                work[w] = make_int2((int)r, c);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_overflow) fb_rows[atomicAdd(fb_count, 1)] = (int)r;
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

template <int KR>
int launch_fused(chb_ctx *c, const CUtensorMap &ma, const CUtensorMap &mb, int64_t nrows)
{
    const size_t smem = sizeof(SharedStorage) + 1024;
    static bool configured = false;
    if (!configured) {
        CHB_CUDA(c, cudaFuncSetAttribute(gram_select_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int nrb = (int)((nrows + BM - 1) / BM);
    const int grid = std::min(nrb, c->sm_count);
    {
        chb_stage_timer t(c, CHB_ST_KNN);
        gram_select_kernel<KR><<<grid, FUSED_THREADS, smem, c->stream>>>(
            ma, mb, c->Kp / BK, c->f_ntiles, c->f_tile_bin, c->f_col_pt, c->f_col_a, c->f_col_b, c->f_col_nrm, c->nrm,
            c->qpoint + c->u0, c->pos, nrows, c->C, c->f_cand_key, c->f_cand_idx);
    }
    CHB_CUDA(c, cudaGetLastError());
    return CHB_OK;
}

template <typename T>
int reserve(chb_ctx *ctx, T **p, int64_t *cap, int64_t count)
{
    if (*p && *cap >= count) return CHB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), sizeof(T) * (size_t)std::max<int64_t>(count, 1));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return chb_fail(ctx, CHB_ENOMEM, "cudaMalloc of %lld bytes failed: %s", (long long)(sizeof(T) * (size_t)count),
                        cudaGetErrorString(e));
    }
    *cap = count;
    return CHB_OK;
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

} // namespace

bool chb_fused_supported(const chb_ctx *c) { return c->k + 3 <= 16; }

void chb_fused_free(chb_ctx *c)
{
    cudaFree(c->f_bin_cnt); cudaFree(c->f_seg_off); cudaFree(c->f_cursor); cudaFree(c->f_tile_bin); cudaFree(c->f_ntiles);
    cudaFree(c->f_col_pt); cudaFree(c->f_col_a); cudaFree(c->f_col_b); cudaFree(c->f_col_nrm); cudaFree(c->f_bperm);
    cudaFree(c->f_cand_key); cudaFree(c->f_cand_idx); cudaFree(c->f_fb_rows);
    c->f_bin_cnt = c->f_seg_off = c->f_cursor = c->f_tile_bin = c->f_ntiles = c->f_col_pt = c->f_col_a = c->f_col_b = nullptr;
    c->f_col_nrm = c->f_bperm = c->f_cand_key = nullptr;
    c->f_cand_idx = c->f_fb_rows = nullptr;
}

// Runs steps 1-3 of the header comment for ALL owned query slots against the current (pos, tent, old) labels.
// Appends changed (slot_local, bin) pairs to ctx->work (count in counters[0]); queries needing the exact fallback are
// listed in f_fb_rows (count in counters[6]).
int chb_round_fused(chb_ctx *c)
{
    const int64_t nown = c->u1 - c->u0;
    if (nown <= 0) return CHB_OK;
    const int64_t n = c->n;
    const int32_t C = c->C, k = c->k;
    const int KR = (k + 3 <= 8) ? 8 : 16;
    c->Kp = (3 * c->d + 31) & ~31;
    const int64_t ncol_max = ((2 * n + (int64_t)BN * C + BN - 1) / BN) * BN;

    CHB_CHECK(c, chb_fused_supported(c), CHB_EINVAL, "fused mode supports num_neighbors <= 13");
    if (c->f_cap_bins < C + 1) {
        int64_t z = 0;
        z = 0; if (reserve(c, &c->f_bin_cnt, &z, C + 1)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_seg_off, &z, C + 2)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_cursor, &z, C + 1)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_ntiles, &z, 4)) return CHB_ENOMEM;
        c->f_cap_bins = C + 1;
    }
    if (c->f_cap_cols < ncol_max) {
        int64_t z = 0;
        z = 0; if (reserve(c, &c->f_col_pt, &z, ncol_max)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_col_a, &z, ncol_max)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_col_b, &z, ncol_max)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_col_nrm, &z, ncol_max)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_tile_bin, &z, ncol_max / BN + 1)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_bperm, &z, ncol_max * c->Kp)) return CHB_ENOMEM;
        c->f_cap_cols = ncol_max;
    }
    if (c->f_cap_cand < nown * C * KR * 2) {
        int64_t z = 0;
        z = 0; if (reserve(c, &c->f_cand_key, &z, nown * C * KR * 2)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_cand_idx, &z, nown * C * KR * 2)) return CHB_ENOMEM;
        z = 0; if (reserve(c, &c->f_fb_rows, &z, nown)) return CHB_ENOMEM;
        c->f_cap_cand = nown * C * KR * 2;
    }
    // split operands: points (once per feature set) and owned query rows (once per label set)
    if (!c->bsplit_ready || c->cap_Bsplit < n * c->Kp) {
        int64_t z = c->cap_Bsplit;
        if (reserve(c, &c->Bsplit, &z, n * c->Kp)) return CHB_ENOMEM;
        c->cap_Bsplit = z;
    }
    if (!c->f_asplit_ready || c->cap_Asplit < nown * c->Kp) {
        int64_t z = c->cap_Asplit;
        if (reserve(c, &c->Asplit, &z, nown * c->Kp)) return CHB_ENOMEM;
        c->cap_Asplit = z;
    }
    if (!c->bsplit_ready || !c->f_asplit_ready) {
        int rc = chb_gram_tc_prepare(c, c->qpoint + c->u0, c->f_asplit_ready ? 0 : nown, c->Asplit, c->Bsplit, c->Kp, !c->bsplit_ready);
        if (rc != CHB_OK) return rc;
        c->bsplit_ready = true;
        c->f_asplit_ready = true;
    }

    // ---- 1. column entries
    CHB_CUDA(c, cudaMemsetAsync(c->f_bin_cnt, 0, sizeof(int32_t) * (size_t)(C + 1), c->stream));
    entries_count_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(c->tent_pt, c->old_label, n, C, c->f_bin_cnt);
    entries_scan_kernel<<<1, 256, 0, c->stream>>>(c->f_bin_cnt, C, c->f_seg_off, c->f_cursor, c->f_tile_bin, c->f_ntiles);
    entries_fill_kernel<<<nblk(ncol_max, 256), 256, 0, c->stream>>>(ncol_max, c->f_col_pt, c->f_col_a, c->f_col_b);
    entries_scatter_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(c->tent_pt, c->old_label, c->pos, n, C, c->f_seg_off, c->f_cursor,
                                                                 c->f_col_pt, c->f_col_a, c->f_col_b);
    entries_gather_kernel<<<nblk(ncol_max * (c->Kp / 4), 256), 256, 0, c->stream>>>(c->f_col_pt, c->f_ntiles, c->Bsplit, c->nrm, c->Kp,
                                                                                    c->f_bperm, c->f_col_nrm);
    CHB_CUDA(c, cudaGetLastError());
    c->tm.launches_other += 5;

    // ---- 2. fused Gram + selection
    CUtensorMap ma, mb;
    int rc = make_map(c, &ma, c->Asplit, nown, c->Kp);
    if (rc != CHB_OK) return rc;
    rc = make_map(c, &mb, c->f_bperm, ncol_max, c->Kp);
    if (rc != CHB_OK) return rc;
    rc = (KR == 8) ? launch_fused<8>(c, ma, mb, nown) : launch_fused<16>(c, ma, mb, nown);
    if (rc != CHB_OK) return rc;
    c->tm.rows_scanned += nown;

    // ---- 3. re-rank
    CHB_CUDA(c, cudaMemsetAsync(&c->counters[6], 0, sizeof(int32_t), c->stream));
    {
        chb_stage_timer t(c, CHB_ST_KNN);
        const double eps_rel = (double)(3 * c->d + 64) * 1.1920928955078125e-07;
        rerank_kernel<<<(unsigned)nown, 128, sizeof(double) * (size_t)((c->d + 1) & ~1), c->stream>>>(
            c->f_cand_key, c->f_cand_idx, KR, c->f_bin_cnt, c->X, c->ldx, c->d, c->qpoint + c->u0, c->nrm,
            reinterpret_cast<const unsigned int *>(&c->counters[5]), eps_rel, C, k, c->knn_idx, c->knn_cnt, c->work, c->counters,
            c->f_fb_rows, &c->counters[6]);
    }
    CHB_CUDA(c, cudaGetLastError());
    return CHB_OK;
}
