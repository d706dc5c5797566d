// qp_lane.cu -- hull-distance QPs with 11..24 neighbours: FP64 tensor-core Gram + one lane per pair (sm_100a).
//
// Same contract as qp.cu / qp_small.cu / qp_mid.cu (hull_distance.py:7-35 + solve_qp.py:18-51 + quadprog):
//   minimise a'Ga on the simplex, G = W W', W = V - 1x' (the k neighbour rows minus the query), distance = sqrt(a'Ga).
// The warp-per-pair kernel of qp.cu spends ~6000 warp instructions on a 12-vertex problem (13 of 32 lanes busy, a
// shuffle reduction or a shared-memory broadcast between any two steps).  Here the work is split in two kernels:
//  * gram_dmma_kernel (one warp per pair, no shared memory, full occupancy: the L2 gather latency hides behind other
//    warps) forms the Gram matrix on the FP64 tensor cores: DMMA m8n8k4 with the SAME register fragment as A and as B
//    (lane (g, t) holds W[8b + g][col + t] for row block b -- the A layout (row g, k t) and the B layout (k t, column g)
//    coincide for a Gram matrix), operands straight from L2 with 16-byte loads (two k-steps per load), the lower triangle
//    of the 8 x 8 tiles per k-step.  G goes to a scratch laid out [batch of 32 pairs][entry][pair in batch];
//  * qp_lane_solve_kernel gives every lane ITS OWN pair of a batch: block principal pivoting on the exchanged tableau of
//    M = G + s 11' kept in shared memory as a symmetric lower triangle, one private copy per lane ([entry][lane]: a lane
//    only ever touches its own bank pair, so lane-dependent entry indices -- the pivot row -- are conflict-free and cost
//    nothing extra to address).  A sweep reads r_i = sum_{c in S} T_ic off the tableau: for i
//    in S that is y_i (weights ~ r_i / sum r), for i outside S it is (M y)_i, and the multiplier test g_i < f is r_i < 1.
//    All vertices with a negative weight leave and all violated ones enter, one principal pivot (rank-1 update of the
//    triangle) each;
//  * the face a lane ends on is only accepted after the KKT conditions have been checked against the UNTOUCHED Gram
//    matrix (read again from the scratch): weights >= 0, multipliers of every excluded vertex (banned ones included)
//    >= -tol, stationarity on the face.  The objective a'Ga of a feasible a is second-order accurate in the rounding the
//    pivots accumulated, so no re-solve is needed;
//  * as in qp_mid.cu: distance = sqrt(a'Ga) when a'Ga > 1e-5 max G_ii; otherwise -- and whenever a lane did not end on a
//    verified face (cycling, vanishing pivots from duplicate contigs, the sweep cap) -- the pair goes to qp.cu.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int LD = 32;        // pitch of one tableau entry across the 32 lanes (doubles): lane l only ever touches bank pair l mod 16
constexpr int SWEEP_CAP = 10;
constexpr int GRAM_WARPS = 8;

__host__ __device__ __forceinline__ constexpr int lidx(int i, int j) { return i * (i + 1) / 2 + j; } // i >= j

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------ Gram matrices
// One warp per work item of the chunk [item0, item0 + chunk): G (lower triangle, KMAX (KMAX + 1) / 2 entries) of the pair
// to gout[(local >> 5) * NE * 32 + entry * 32 + (local & 31)], local = item - item0.
template <int KMAX>
__global__ void __launch_bounds__(GRAM_WARPS * 32) gram_dmma_kernel(chb_qp_args a, int64_t item0, int64_t chunk, double *__restrict__ gout)
{
    constexpr int NE = KMAX * (KMAX + 1) / 2;
    constexpr int NB = (KMAX + 7) / 8;
    constexpr int NT = NB * (NB + 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int64_t n_work = a.work_count ? (int64_t)*a.work_count : a.n_work;
    const int64_t end = (item0 + chunk < n_work) ? item0 + chunk : n_work;
    const int ldx = a.ldx, k = a.k, C = a.C;
    const int nfull = ldx / 8; // k-step pairs (8 columns) whose columns are inside the row for every lane

    for (int64_t item = item0 + (int64_t)blockIdx.x * GRAM_WARPS + warp; item < end; item += (int64_t)gridDim.x * GRAM_WARPS) {
        const int2 wk = a.work[item];
        const int64_t pair = (int64_t)wk.x * C + wk.y;
        const int m = a.knn_cnt[pair];
        if (m <= 0) continue; // the solve kernel does not read G of an empty bin
        const int qpt = a.row_point[wk.x];
        const double *xq = a.X + (int64_t)qpt * ldx + 2 * tq;
        const double *rp[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int r = 8 * b + gq;
            rp[b] = a.X + (int64_t)((r < m && r < KMAX) ? a.knn_idx[pair * k + r] : qpt) * ldx + 2 * tq; // r >= m: the query row, W row = 0
        }
        // lane (g, t) covers columns c, c + 1 with c = 8 step + 2t (one 16-byte load per row block): two k-steps of the MMA, whose
        // k slot t then means column c (first, accumulator set 0) / c + 1 (second, set 1) for A and B alike
        double acc[2][NT][2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[u][i][0] = acc[u][i][1] = 0.0;
#pragma unroll 3
        for (int st = 0; st < nfull; ++st) {
            const double2 xv = __ldg(reinterpret_cast<const double2 *>(xq + 8 * st));
            double2 w[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(rp[b] + 8 * st));
                w[b].x = v.x - xv.x;
                w[b].y = v.y - xv.y;
            }
#pragma unroll
            for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    dmma884(acc[0][lidx(bi, bj)], w[bi].x, w[bj].x);
                    dmma884(acc[1][lidx(bi, bj)], w[bi].y, w[bj].y);
                }
        }
        if (nfull * 8 < ldx) { // the ragged last step: lanes whose columns lie beyond the row contribute w = 0
            const bool inb = nfull * 8 + 2 * tq < ldx; // ldx is even: a lane's two columns are inside or outside together
            double2 xv = make_double2(0.0, 0.0);
            if (inb) xv = __ldg(reinterpret_cast<const double2 *>(xq + 8 * nfull));
            double2 w[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                w[b] = make_double2(0.0, 0.0);
                if (inb) {
                    const double2 v = __ldg(reinterpret_cast<const double2 *>(rp[b] + 8 * nfull));
                    w[b].x = v.x - xv.x;
                    w[b].y = v.y - xv.y;
                }
            }
#pragma unroll
            for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    dmma884(acc[0][lidx(bi, bj)], w[bi].x, w[bj].x);
                    dmma884(acc[1][lidx(bi, bj)], w[bi].y, w[bj].y);
                }
        }
        // C fragment: lane (g, t) holds rows 8 bi + g, columns 8 bj + 2t, + 1
        const int64_t local = item - item0;
        double *go = gout + (local >> 5) * (int64_t)(NE * 32) + (local & 31);
#pragma unroll
        for (int bi = 0; bi < NB; ++bi)
#pragma unroll
            for (int bj = 0; bj <= bi; ++bj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = 8 * bi + gq, j = 8 * bj + 2 * tq + e;
                    if (i < KMAX && j <= i) go[lidx(i, j) * 32] = acc[0][lidx(bi, bj)][e] + acc[1][lidx(bi, bj)][e];
                }
    }
}

// ------------------------------------------------------------------------------------------------ one lane per pair
// One principal pivot of the lane's tableau on vertex kv (in or out of the face S).  With T the symmetric store of the
// exchanged tableau (T_SS = (M_SS)^-1, T_NS = M_NS (M_SS)^-1, T_NN = Schur complement):
//   T_ij <- T_ij - pi_ij T_ik T_jk / p,   pi_ij = -1 if i and j are both on the side kv is moving TO, else +1
//   T_jk <- -/+ T_jk / p (minus for j in S),   T_kk <- 1 / p
// Returns false (tableau untouched) if the pivot p is not above min_piv.
template <int KMAX>
__device__ __forceinline__ bool lane_pivot(double *__restrict__ T, int kv, unsigned &S, double min_piv)
{
    const int bk = kv * (kv + 1) / 2;
    const double p = T[(bk + kv) * LD];
    if (!(p > min_piv)) return false;
    const double rinv = 1.0 / p;
    const bool entering = !((S >> kv) & 1u);
    const unsigned B = entering ? S : ~S;
    // straight-line code over all KMAX (KMAX + 1) / 2 entries (rows beyond the pair's m are bystanders that nobody reads): no
    // per-row branch, so the loads of later rows are in flight behind the arithmetic of earlier ones
    double c[KMAX], h[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        const int e = (i >= kv) ? (i * (i + 1) / 2 + kv) : (bk + i); // entry (max, min) of the triangle
        c[i] = T[e * LD]; // c[kv] = p spoils row / column kv in the update below: they are rewritten afterwards
        h[i] = ((B >> i) & 1u) ? c[i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        const double fi = c[i] * rinv;
        const double hi = 2.0 * h[i] * rinv;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double v = T[lidx(i, j) * LD];
            v = fma(-fi, c[j], v);
            v = fma(hi, h[j], v);
            T[lidx(i, j) * LD] = v;
        }
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        const int e = (j >= kv) ? (j * (j + 1) / 2 + kv) : (bk + j);
        const double v = c[j] * rinv;
        if (j != kv) T[e * LD] = ((S >> j) & 1u) ? -v : v;
    }
    T[(bk + kv) * LD] = rinv;
    S ^= 1u << kv;
    return true;
}

template <int KMAX, int LANE_WARPS>
__global__ void __launch_bounds__(LANE_WARPS * 32) qp_lane_solve_kernel(chb_qp_args a, int64_t item0, int64_t chunk, const double *__restrict__ gin,
                                                                         int2 *__restrict__ fallback, int32_t *__restrict__ fallback_count)
{
    constexpr int NE = KMAX * (KMAX + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *T = reinterpret_cast<double *>(smem_raw) + (size_t)warp * NE * LD + lane;
    const int64_t n_work = a.work_count ? (int64_t)*a.work_count : a.n_work;
    const int64_t end = (item0 + chunk < n_work) ? item0 + chunk : n_work;
    const int k = a.k, C = a.C;
    const int64_t stride = (int64_t)gridDim.x * LANE_WARPS * 32;

    for (int64_t base = item0 + ((int64_t)blockIdx.x * LANE_WARPS + warp) * 32; base < end; base += stride) {
        const int64_t my_item = base + lane;
        const bool valid = my_item < end;
        int2 wk = make_int2(0, 0);
        int64_t pair = 0;
        int m = 0;
        if (valid) {
            wk = a.work[my_item];
            pair = (int64_t)wk.x * C + wk.y;
            m = a.knn_cnt[pair];
        }
        const double *scr = gin + ((base - item0) >> 5) * (int64_t)(NE * 32) + lane; // G of this lane's pair: scr[e * 32]
        double best = 0.0, scale = 0.0;
        int status = CHB_QP_OK;
        bool solved = false;
        double beta[KMAX];
#pragma unroll
        for (int i = 0; i < KMAX; ++i) beta[i] = 0.0;
        int start = 0;
        {
            double dmin = DBL_MAX;
#pragma unroll
            for (int i = 0; i < KMAX; ++i) {
                if (i < m) {
                    const double gii = scr[lidx(i, i) * 32];
                    scale = fmax(scale, gii);
                    if (gii < dmin) { dmin = gii; start = i; }
                }
            }
        }
        const bool active = m > 0 && scale > 0.0;
        if (m > 0 && !(scale > 0.0)) { // every neighbour coincides with the query (or NaN input: best stays 0 as in qp_mid.cu)
            beta[0] = 1.0;
            solved = true;
        }
        // the tableau starts at M = G + scale 11' (G stays untouched in the scratch for the final check)
        if (__any_sync(CHB_FULL, active)) {
#pragma unroll
            for (int i = 0; i < KMAX; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) T[lidx(i, j) * LD] = scr[lidx(i, j) * 32] + scale;
        }
        const double tol = 1e-14 * scale;
        unsigned S = 0u, banned = 0u;
        bool ended = false, broke = !active;
        double sy = 0.0;
        double r[KMAX];
        if (active) {
            lane_pivot<KMAX>(T, start, S, 0.0); // G_ss + scale >= scale > 0
        }
#pragma unroll 1
        for (int sweep = 0; sweep < SWEEP_CAP; ++sweep) {
            const bool live = !ended && !broke;
            if (!__any_sync(CHB_FULL, live)) break;
            unsigned flip = 0u, dual = 0u;
            if (live) {
#pragma unroll
                for (int i = 0; i < KMAX; ++i) r[i] = 0.0;
#pragma unroll
                for (int i = 0; i < KMAX; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) {
                        const double v = T[lidx(i, j) * LD];
                        if ((S >> j) & 1u) r[i] += v;
                        if (j != i && ((S >> i) & 1u)) r[j] += v;
                    }
                sy = 0.0;
#pragma unroll
                for (int i = 0; i < KMAX; ++i)
                    if ((S >> i) & 1u) sy += r[i];
                if (!(sy > 0.0)) {
                    broke = true;
                } else {
                    const double thr = 1.0 - tol * sy;
                    unsigned neg = 0u;
#pragma unroll
                    for (int i = 0; i < KMAX; ++i) {
                        if (i < m) {
                            if ((S >> i) & 1u) {
                                if (r[i] < 0.0) neg |= 1u << i;
                            } else if (!((banned >> i) & 1u) && r[i] < thr) {
                                dual |= 1u << i;
                            }
                        }
                    }
                    flip = neg | dual;
                    if (!flip) ended = true;
                }
            }
            while (__any_sync(CHB_FULL, flip != 0u)) {
                if (flip) {
                    const int v = __ffs(flip) - 1;
                    flip &= flip - 1;
                    const bool entering = (dual >> v) & 1u;
                    if (lane_pivot<KMAX>(T, v, S, entering ? 2e-11 * scale : 0.0)) {
                        if (!entering) banned = 0u; // the face shrank: a vertex that depended on it may be independent now
                    } else if (entering) {
                        banned |= 1u << v; // affinely dependent on the face: same hull without it
                        status = CHB_QP_DEGENERATE;
                    } else {
                        broke = true; // a diagonal entry of an SPD inverse came out non-positive: drift
                        flip = 0u;
                    }
                }
            }
        }
        // KKT check of the face the lane ended on, against the parked G
        if (__any_sync(CHB_FULL, ended)) {
            if (ended) {
                const double isy = 1.0 / sy;
#pragma unroll
                for (int i = 0; i < KMAX; ++i) beta[i] = ((S >> i) & 1u) ? r[i] * isy : 0.0;
                double g[KMAX];
#pragma unroll
                for (int i = 0; i < KMAX; ++i) g[i] = 0.0;
#pragma unroll
                for (int i = 0; i < KMAX; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) {
                        const double v = scr[lidx(i, j) * 32];
                        g[i] = fma(v, beta[j], g[i]);
                        if (j != i) g[j] = fma(v, beta[i], g[j]);
                    }
                double f = 0.0;
#pragma unroll
                for (int i = 0; i < KMAX; ++i) f = fma(beta[i], g[i], f);
                bool ok = f == f;
#pragma unroll
                for (int i = 0; i < KMAX; ++i) {
                    if (i < m) {
                        if ((S >> i) & 1u)
                            ok = ok && beta[i] >= 0.0 && fabs(g[i] - f) <= 1e-9 * scale;
                        else
                            ok = ok && g[i] >= f - tol;
                    }
                }
                if (ok) {
                    best = f;
                    solved = true;
                }
            }
        }
        if (valid) {
            if (m <= 0) {
                a.dist[pair] = INFINITY;
                if (a.status) a.status[pair] = CHB_QP_EMPTY_BIN;
            } else {
                const bool exact_needed = !solved || !(best == best) || (scale > 0.0 && !(best > 1e-5 * scale) && best != 0.0);
                if (exact_needed) {
                    const int w = atomicAdd(fallback_count, 1);
                    fallback[w] = wk;
                } else {
                    a.dist[pair] = sqrt(fmax(best, 0.0));
                    if (a.status) a.status[pair] = status;
                    if (a.alpha) {
#pragma unroll
                        for (int i = 0; i < KMAX; ++i)
                            if (i < k) a.alpha[pair * k + i] = beta[i];
                    }
                }
            }
        }
        __syncwarp();
    }
}


template <int KMAX, int LANE_WARPS>
int launch_lane(chb_ctx *ctx, const chb_qp_args &a, int2 *fallback, int32_t *fallback_count)
{
    constexpr int NE = KMAX * (KMAX + 1) / 2;
    const size_t smem = sizeof(double) * (size_t)LANE_WARPS * NE * LD;
    CHB_CUDA(ctx, cudaFuncSetAttribute(qp_lane_solve_kernel<KMAX, LANE_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CHB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_lane_solve_kernel<KMAX, LANE_WARPS>, LANE_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    // The exact number of pairs (the rounds only keep it on the device; the list is a fraction of its bound n_work): the
    // chunking below follows it, at the price of one 4-byte read-back instead of dozens of empty launches
    int64_t total = a.n_work;
    if (a.work_count) {
        CHB_CUDA(ctx, cudaMemcpyAsync(&ctx->counters_host[15], a.work_count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        total = ctx->counters_host[15] < a.n_work ? ctx->counters_host[15] : a.n_work;
    }
    if (total <= 0) return CHB_OK;
    // The Gram scratch holds one chunk of pairs (NE doubles each), at most 1 GiB: the two kernels run chunk after chunk.
    // (Running the Gram kernel of the next chunk on a second stream beside the solve kernel was tried: the two compete for
    // CTA slots and the pair came out slower than back to back for k <= 16, within 7 % for 20 and 24.)
    int64_t chunk_cap = ((int64_t)1 << 30) / (int64_t)(sizeof(double) * NE);
    chunk_cap &= ~(int64_t)31;
    const int64_t chunk = total < chunk_cap ? ((total + 31) & ~(int64_t)31) : chunk_cap;
    const int64_t need = chunk * NE;
    if (ctx->qp_scratch_cap < need) {
        if (ctx->qp_scratch) cudaFree(ctx->qp_scratch);
        ctx->qp_scratch = nullptr;
        ctx->qp_scratch_cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->qp_scratch), sizeof(double) * (size_t)need);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return chb_fail(ctx, CHB_ENOMEM, "cudaMalloc of the QP Gram scratch (%lld MB) failed: %s", (long long)(need >> 17), cudaGetErrorString(e));
        }
        ctx->qp_scratch_cap = need;
    }
    for (int64_t item0 = 0; item0 < total; item0 += chunk) {
        const int64_t n = (total - item0 < chunk) ? total - item0 : chunk;
        int64_t gblocks = (n + GRAM_WARPS - 1) / GRAM_WARPS;
        const int64_t gcap = (int64_t)ctx->sm_count * 8;
        if (gblocks > gcap) gblocks = gcap;
        int64_t blocks = (n + LANE_WARPS * 32 - 1) / (LANE_WARPS * 32);
        const int64_t cap = (int64_t)ctx->sm_count * per_sm;
        if (blocks > cap) blocks = cap;
        {
            chb_stage_timer t(ctx, CHB_ST_QP);
            gram_dmma_kernel<KMAX><<<(unsigned)gblocks, GRAM_WARPS * 32, 0, ctx->stream>>>(a, item0, n, ctx->qp_scratch);
        }
        {
            chb_stage_timer t(ctx, CHB_ST_QP);
            qp_lane_solve_kernel<KMAX, LANE_WARPS><<<(unsigned)blocks, LANE_WARPS * 32, smem, ctx->stream>>>(a, item0, n, ctx->qp_scratch,
                                                                                                      fallback, fallback_count);
        }
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}

} // namespace

int chb_launch_qp_lane(chb_ctx *ctx, const chb_qp_args &a, int2 *fallback, int32_t *fallback_count)
{
    // warps per CTA chosen so that the per-lane tableaus (k (k + 1) / 2 entries x 32 x 8 bytes per warp) fill the SM's shared memory
    if (a.k <= 12) return launch_lane<12, 2>(ctx, a, fallback, fallback_count); // 20 KB per warp
    if (a.k <= 14) return launch_lane<14, 2>(ctx, a, fallback, fallback_count); // 27 KB per warp
    if (a.k <= 16) return launch_lane<16, 2>(ctx, a, fallback, fallback_count); // 35 KB per warp: 3 CTAs
    if (a.k <= 20) return launch_lane<20, 1>(ctx, a, fallback, fallback_count); // 54 KB per warp: 4 CTAs
    return launch_lane<24, 1>(ctx, a, fallback, fallback_count);                // 75 KB per warp: 3 CTAs fill the 228 KB exactly
}
