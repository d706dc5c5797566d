// qp.cu -- batched point-to-hull distance QPs, one warp per (query, bin) pair (sm_100a).
//
// Replaces, per pair, the whole chain
//   calculate_distance -> convex_hull_distance   /root/reference/ch_bin/core/clustering/hull_distance.py:90-108, 7-35
//   solve_qp -> _quadprog_solve_qp               /root/reference/ch_bin/core/clustering/solve_qp.py:96-132, 18-51
//   nearest_positive_definite                    /root/reference/ch_bin/core/clustering/positive_def.py:25-48
//   quadprog.solve_qp (third party, Goldfarb-Idnani)
// i.e.   minimise || x - V' alpha ||^2   s.t.  alpha >= 0, sum(alpha) = 1,   return || alpha V - x ||_2
// (the distance is recomputed from alpha in d dimensions exactly as hull_distance.py:34-35 does).
//
// Method.  With W = V - 1 x' the objective is alpha' G alpha, G = W W' (m x m, PSD): the minimum-norm point of
// the hull of the rows of W.  Phase 1 forms G in FP64 (lane i accumulates row i from a shared-memory tile of W);
// phase 2 finds the optimal face -- lane i owns vertex i -- by block principal pivoting on an exchanged tableau of
// M = G + s 11' (main launches) with Wolfe's finite active-set method as the safety net (and as the method of the
// fallback launches): the affine minimiser on a face S solves (M_SS) y = 1, rows distributed over lanes; affinely
// dependent vertices (duplicate contigs) show up as a vanishing pivot and are banned, which leaves the distance
// unchanged; phase 3 evaluates the residual norm in d dimensions from global memory.
// The QP is strictly convex on the affine hull, so its distance is unique: any exact solver agrees with
// quadprog to rounding; the parity bar is 1e-6 relative, this kernel is good to ~1e-12.
//
// Roofline: k neighbour rows (8kd bytes) per pair from L2/HBM against k(k+1)d + 4kd + 3d + 2k^3 FP64 flops.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int QP_WARPS = 4;
constexpr int TC = 32;        // feature columns per staged tile
constexpr int LDW = TC + 2;   // padded tile pitch (doubles), keeps double2 alignment

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CHB_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(CHB_FULL, v, o));
    return v;
}
// minimum with the lowest lane winning ties
__device__ __forceinline__ void warp_argmin(double &v, int &idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(CHB_FULL, v, o);
        const int oi = __shfl_xor_sync(CHB_FULL, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// Solve (G_SS + s 11') y = 1 on the vertices in smask; rows over lanes, Gauss-Jordan without pivoting (SPD).
// Returns the lane index of a vanishing pivot (affinely dependent vertex) or -1.  y is 0 outside S.
template <int KMAX>
__device__ __forceinline__ int affine_solve(const double *sG, double *sRow, unsigned smask, int m, double shift, int lane,
                                            double &y)
{
    const bool in = (lane < m) && ((smask >> lane) & 1u);
    double A[KMAX];
    double rhs = in ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
        const bool cin = (c < m) && ((smask >> c) & 1u);
        A[c] = (in && cin) ? sG[lane * LDW + c] + shift : ((c == lane) ? 1.0 : 0.0);
    }
    int bad = -1;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j >= m || !((smask >> j) & 1u)) continue; // warp-uniform
        if (lane == j) {
#pragma unroll
            for (int c = 0; c < KMAX; ++c) sRow[c] = A[c];
            sRow[KMAX] = rhs;
        }
        __syncwarp();
        const double piv = sRow[j];
        const double ref = sG[j * LDW + j] + shift;
        if (!(piv > 1e-11 * ref)) { bad = j; __syncwarp(); break; }
        if (lane != j) {
            const double f = A[j] / piv;
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c >= j) A[c] -= f * sRow[c];
            rhs -= f * sRow[KMAX];
        }
        __syncwarp();
    }
    double diag = 1.0;
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
        if (c == lane) diag = A[c];
    y = in ? rhs / diag : 0.0;
    return bad;
}

// Principal pivot (Gauss-Jordan exchange) of the tableau on vertex kv; lane i holds row i.  Starting from M = G + s 11' and
// exchanging the vertices of a corral S one by one leaves (M_SS)^-1 in the S block, whatever the order, and exchanging a
// vertex again takes it out: moving a vertex in or out of the corral costs ONE rank-1 update of the rows instead of a fresh
// factorisation of the corral.  The diagonal entry met when a vertex enters is its Schur complement (affine dependence shows
// up there); when it leaves it is a diagonal entry of an SPD inverse.  Returns false if the pivot is not above min_piv.
template <int KMAX>
__device__ __forceinline__ bool exchange(double (&A)[KMAX], double *sRow, int kv, int lane, double min_piv)
{
    if (lane == kv) {
#pragma unroll
        for (int c = 0; c < KMAX; ++c) sRow[c] = A[c];
    }
    __syncwarp();
    const double piv = sRow[kv];
    if (!(piv > min_piv)) {
        __syncwarp();
        return false;
    }
    const double rinv = 1.0 / piv;
    if (lane == kv) {
#pragma unroll
        for (int c = 0; c < KMAX; ++c) A[c] = (c == kv) ? rinv : -A[c] * rinv;
    } else {
        double aik = 0.0;
#pragma unroll
        for (int c = 0; c < KMAX; ++c)
            if (c == kv) aik = A[c];
        const double f = aik * rinv;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) A[c] = (c == kv) ? f : fma(-f, sRow[c], A[c]);
    }
    __syncwarp();
    return true;
}

template <int KMAX>
__global__ void __launch_bounds__(QP_WARPS * 32) qp_kernel(chb_qp_args a, int fast)
{
    chb_pdl_wait();
    __shared__ __align__(16) double sWall[QP_WARPS][KMAX * LDW];
    __shared__ __align__(16) double sRowAll[QP_WARPS][KMAX + 2];
    __shared__ __align__(16) double sAlphaAll[QP_WARPS][KMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sW = sWall[warp];
    double *sRow = sRowAll[warp];
    double *sAlpha = sAlphaAll[warp];
    const int64_t n_work = a.work_count ? (int64_t)*a.work_count : a.n_work;
    const int d = a.d, ldx = a.ldx, k = a.k, C = a.C;

    for (int64_t item = (int64_t)blockIdx.x * QP_WARPS + warp; item < n_work; item += (int64_t)gridDim.x * QP_WARPS) {
        const int2 wk = a.work[item];
        const int64_t pair = (int64_t)wk.x * C + wk.y;
        const int m = a.knn_cnt[pair];
        if (m <= 0) {
            if (lane == 0) {
                a.dist[pair] = INFINITY;
                if (a.status) a.status[pair] = CHB_QP_EMPTY_BIN;
            }
            continue;
        }
        const int jq = a.row_point[wk.x];
        const int myidx = lane < m ? a.knn_idx[pair * k + lane] : 0;
        const double *__restrict__ xq = a.X + (int64_t)jq * ldx;

        // ---------------- phase 1: G = W W', lane i accumulates row i; with at most 16 rows both half-warps work: lane i and
        // lane i + 16 take half of each tile's columns of row i
        constexpr bool SPLIT = KMAX <= 16;
        constexpr int TSPAN = SPLIT ? TC / 2 : TC;
        const int grow = SPLIT ? (lane & 15) : lane;
        const int tbeg = SPLIT ? (lane >> 4) * TSPAN : 0;
        double G[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) G[c] = 0.0;
        for (int t0 = 0; t0 < d; t0 += TC) {
            const int t = t0 + lane;
            const bool tin = t < d;
            const double xv = tin ? xq[t] : 0.0;
            // eight rows per batch, loads first: a load behind a per-row branch would expose one memory latency per row
            const int tc = tin ? t : 0;
#pragma unroll
            for (int r0 = 0; r0 < KMAX; r0 += 8) {
                if (r0 < m) { // warp-uniform; lanes >= m hold index 0, a valid row
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = a.X[(int64_t)__shfl_sync(CHB_FULL, myidx, r0 + u) * ldx + tc];
#pragma unroll
                    for (int u = 0; u < 8; ++u) sW[(r0 + u) * LDW + lane] = tin ? v[u] - xv : 0.0;
                }
            }
            __syncwarp();
            if (grow < m) {
#pragma unroll 4
                for (int tt = tbeg; tt < tbeg + TSPAN; tt += 2) {
                    const double2 own = *reinterpret_cast<const double2 *>(&sW[grow * LDW + tt]);
#pragma unroll
                    for (int c = 0; c < KMAX; ++c) {
                        if (c < m) {
                            const double2 wc = *reinterpret_cast<const double2 *>(&sW[c * LDW + tt]);
                            G[c] = fma(own.x, wc.x, G[c]);
                            G[c] = fma(own.y, wc.y, G[c]);
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (SPLIT) { // the two half-warps each summed half of every tile's columns for row (lane & 15)
#pragma unroll
            for (int c = 0; c < KMAX; ++c) G[c] += __shfl_xor_sync(CHB_FULL, G[c], 16);
        }
        // G to shared memory (the tile buffer is free now): sG[i][c] at sW[i*LDW + c]
        double gii = 0.0;
        if (lane < m) {
#pragma unroll
            for (int c = 0; c < KMAX; ++c) {
                if (c < m) sW[lane * LDW + c] = G[c];
                if (c == lane) gii = G[c];
            }
        }
        __syncwarp();
        const double *sG = sW;

        // ---------------- phase 2: active set
        const double scale = warp_max(lane < m ? gii : 0.0);
        double alpha = 0.0;
        int status = CHB_QP_OK;
        if (!(scale > 0.0)) {
            alpha = (lane == 0) ? 1.0 : 0.0; // every neighbour coincides with the query
        } else if (a.metric != CHB_METRIC_CONVEX) { // affine-qp / affine: no sign constraints
            unsigned smask = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
            for (int guard = 0; guard < m; ++guard) {
                double y;
                const int bad = affine_solve<KMAX>(sG, sRow, smask, m, scale, lane, y);
                if (bad < 0) {
                    const double sy = warp_sum(y);
                    alpha = y / sy;
                    break;
                }
                smask &= ~(1u << bad); // dependent vertex: same affine hull without it
                status = CHB_QP_DEGENERATE;
            }
        } else {
            // attempt 0 (fast, the main kernel for k > 10): block principal pivoting on an exchanged tableau of M = G + s 11'
            // (principal pivots on the vertices of a face S leave (M_SS)^-1 in the S block).  Per sweep,
            // beta = (M_SS)^-1 1 / (1' (M_SS)^-1 1) is read off the tableau's row sums, every vertex with a negative weight is
            // pivoted OUT and every excluded vertex with a negative multiplier (g_i < beta' G beta) is pivoted IN -- all at once.
            // A sweep costs one pivot per vertex that moves; the whole solve needs about (size of the optimal face + wrong
            // guesses) pivots and a handful of sweeps, where Wolfe's method spends a sweep per vertex.  The sweeps keep no
            // feasibility, so the result only counts if they END on the KKT conditions and a fresh solve on the final face
            // confirms it; otherwise (cycling, drift, a vanishing pivot on the way out) attempt 1 runs Wolfe's method with a
            // fresh Gauss-Jordan solve per minor cycle, as the fallback launches do.
            const double tol = 1e-14 * scale;
            bool solved = false;
            if (fast == 1) {
                double A[KMAX];
#pragma unroll
                for (int c = 0; c < KMAX; ++c) A[c] = (lane < m && c < m) ? sG[lane * LDW + c] + scale : ((c == lane) ? 1.0 : 0.0);
                unsigned S = 0u, banned = 0u;
                int st = CHB_QP_OK;
                {
                    // start from the vertex nearest to the query (as Wolfe's method does): for a query far from the bin the
                    // optimal face has one or two vertices and the first sweeps find it; a query inside the cloud gets most
                    // vertices added by the first sweep, all at once
                    double key = lane < m ? gii : DBL_MAX;
                    int start = lane;
                    warp_argmin(key, start);
                    exchange<KMAX>(A, sRow, start, lane, 0.0); // G_ss + scale >= scale > 0
                    S = 1u << start;
                }
                bool ended = false, broke = false;
                double beta = 0.0;
                for (int sweep = 0; sweep < 2 * m + 8 && !ended && !broke; ++sweep) {
                    const bool in = (S >> lane) & 1u;
                    double y = 0.0;
                    if (in) {
#pragma unroll
                        for (int c = 0; c < KMAX; ++c)
                            if ((S >> c) & 1u) y += A[c];
                    }
                    beta = y / warp_sum(y);
                    if (lane < KMAX) sAlpha[lane] = in ? beta : 0.0;
                    __syncwarp();
                    double g = 0.0;
                    if (lane < m) {
#pragma unroll
                        for (int c = 0; c < KMAX; ++c)
                            if (c < m) g = fma(sG[lane * LDW + c], sAlpha[c], g);
                    }
                    const double f = warp_sum(in ? beta * g : 0.0);
                    __syncwarp();
                    const unsigned neg = __ballot_sync(CHB_FULL, in && beta < 0.0);
                    const unsigned dual = __ballot_sync(CHB_FULL, lane < m && !in && !((banned >> lane) & 1u) && g < f - tol);
                    if (!neg && !dual) { ended = true; break; }
                    unsigned flip = neg | dual;
                    while (flip) {
                        const int v = __ffs(flip) - 1;
                        flip &= flip - 1;
                        const bool entering = (dual >> v) & 1u;
                        if (exchange<KMAX>(A, sRow, v, lane, entering ? 1e-11 * (sG[v * LDW + v] + scale) : 0.0)) {
                            S ^= 1u << v;
                            if (!entering) banned = 0u; // the face shrank: a vertex that depended on it may be independent now
                        } else if (entering) {
                            banned |= 1u << v;
                            st = CHB_QP_DEGENERATE;
                        } else {
                            broke = true; // a diagonal entry of an SPD inverse came out non-positive: drift
                            break;
                        }
                    }
                }
                if (ended) {
                    // confirm on a fresh solve of the final face (removes whatever rounding the pivots accumulated)
                    double y;
                    const int bad = affine_solve<KMAX>(sG, sRow, S, m, scale, lane, y);
                    const double b2 = y / warp_sum(y);
                    const bool in = (S >> lane) & 1u;
                    const unsigned negm = __ballot_sync(CHB_FULL, in && !(b2 >= 0.0));
                    if (bad < 0 && !negm) {
                        // ... and the multipliers of EVERY excluded vertex (banned ones included) on that fresh solution: the
                        // result is accepted only as a verified KKT point, whatever the tableau drifted to on the way
                        if (lane < KMAX) sAlpha[lane] = in ? b2 : 0.0;
                        __syncwarp();
                        double g = 0.0;
                        if (lane < m) {
#pragma unroll
                            for (int c = 0; c < KMAX; ++c)
                                if (c < m) g = fma(sG[lane * LDW + c], sAlpha[c], g);
                        }
                        const double f = warp_sum(in ? b2 * g : 0.0);
                        __syncwarp();
                        const unsigned viol = __ballot_sync(CHB_FULL, lane < m && !in && g < f - tol);
                        if (!viol) {
                            alpha = in ? b2 : 0.0;
                            status = st;
                            solved = true;
                        }
                    }
                }
            }
            // Wolfe's method: the main path for 25..32 neighbours (fast == 2: exchanged tableau kept across its iterations -- there
            // the block pivoting above moves too many vertices the wrong way to pay off), and the safety net of the block pivoting
            // (a fresh Gauss-Jordan solve per minor cycle, as in the fallback launches)
            for (int attempt = (fast == 2) ? 0 : 1; attempt < 2 && !solved; ++attempt) {
                const bool tab = attempt == 0;
                bool tab_fail = false;
                status = CHB_QP_OK;
                double key = lane < m ? gii : DBL_MAX;
                int start = lane;
                warp_argmin(key, start);
                alpha = (lane == start) ? 1.0 : 0.0;
                unsigned smask = 1u << start, banned = 0u;
                double A[KMAX];
                if (tab) {
#pragma unroll
                    for (int c = 0; c < KMAX; ++c) A[c] = (lane < m && c < m) ? sG[lane * LDW + c] + scale : ((c == lane) ? 1.0 : 0.0);
                    exchange<KMAX>(A, sRow, start, lane, 0.0); // G_ss + scale >= scale > 0
                }
                const int itmax = 3 * m + 8;
                int it = 0;
                for (; it < itmax && !tab_fail; ++it) {
                    if (lane < KMAX) sAlpha[lane] = alpha;
                    __syncwarp();
                    double g = 0.0;
                    if (lane < m) {
#pragma unroll
                        for (int c = 0; c < KMAX; ++c)
                            if (c < m) g = fma(sG[lane * LDW + c], sAlpha[c], g);
                    }
                    const double f = warp_sum(alpha * g);
                    const bool cand = (lane < m) && !((smask >> lane) & 1u) && !((banned >> lane) & 1u) && (g < f - tol);
                    double gk = cand ? g : DBL_MAX;
                    int jn = lane;
                    warp_argmin(gk, jn);
                    if (gk == DBL_MAX) break; // optimal
                    if (tab && !exchange<KMAX>(A, sRow, jn, lane, 1e-11 * (sG[jn * LDW + jn] + scale))) {
                        banned |= 1u << jn; // affinely dependent on the corral: same hull without it
                        status = CHB_QP_DEGENERATE;
                        continue;
                    }
                    smask |= 1u << jn;
                    for (int minor = 0; minor <= m; ++minor) {
                        double y;
                        if (tab) {
                            y = 0.0;
                            if ((smask >> lane) & 1u) {
#pragma unroll
                                for (int c = 0; c < KMAX; ++c)
                                    if ((smask >> c) & 1u) y += A[c];
                            }
                        } else {
                            const int bad = affine_solve<KMAX>(sG, sRow, smask, m, scale, lane, y);
                            if (bad >= 0) {
                                smask &= ~(1u << jn);
                                banned |= 1u << jn;
                                status = CHB_QP_DEGENERATE;
                                break;
                            }
                        }
                        const double beta = y / warp_sum(y);
                        const bool in = (smask >> lane) & 1u;
                        const bool neg = in && !(beta > 0.0);
                        const unsigned negm = __ballot_sync(CHB_FULL, neg);
                        if (!negm) {
                            alpha = in ? beta : 0.0;
                            break;
                        }
                        double th = DBL_MAX;
                        if (neg) th = (alpha > 0.0) ? alpha / (alpha - beta) : 0.0;
                        int lt = lane;
                        warp_argmin(th, lt);
                        th = fmin(fmax(th, 0.0), 1.0);
                        alpha = in ? alpha + th * (beta - alpha) : 0.0;
                        const bool drop = in && (lane == lt || !(alpha > 0.0));
                        unsigned dropm = __ballot_sync(CHB_FULL, drop);
                        smask &= ~dropm;
                        if (drop) alpha = 0.0;
                        if ((dropm >> jn) & 1u) banned |= 1u << jn; // the entering vertex bounced straight out
                        const double sa = warp_sum(alpha);
                        alpha = alpha / sa;
                        if (tab) {
                            while (dropm) { // take the dropped vertices out of the tableau again
                                const int dv = __ffs(dropm) - 1;
                                dropm &= dropm - 1;
                                if (!exchange<KMAX>(A, sRow, dv, lane, 0.0)) { tab_fail = true; break; }
                            }
                            if (tab_fail) break;
                        }
                        if (!((smask >> jn) & 1u)) break;
                    }
                }
                if (tab && (tab_fail || it >= itmax || !(warp_sum(alpha) > 0.5))) continue; // redo without the tableau
                if (it >= itmax) status = CHB_QP_ITER_CAP;
                if (tab) {
                    // polish: one fresh solve on the final corral removes whatever rounding the exchanges accumulated
                    double y;
                    const int bad = affine_solve<KMAX>(sG, sRow, smask, m, scale, lane, y);
                    const double beta = y / warp_sum(y);
                    const bool in = (smask >> lane) & 1u;
                    const unsigned negm = __ballot_sync(CHB_FULL, in && !(beta > 0.0));
                    if (bad < 0 && !negm) alpha = in ? beta : 0.0;
                }
                break;
            }
        }

        // ---------------- phase 3: || alpha V - x ||  in d dimensions (hull_distance.py:34-35)
        double ss = 0.0;
        for (int t0 = 0; t0 < d; t0 += TC) {
            const int t = t0 + lane;
            const bool tin = t < d;
            double pr = 0.0;
            const int tc = tin ? t : 0;
#pragma unroll
            for (int r0 = 0; r0 < KMAX; r0 += 8) {
                if (r0 < m) { // alpha is 0 in the lanes >= m, whose index 0 is a valid row
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = a.X[(int64_t)__shfl_sync(CHB_FULL, myidx, r0 + u) * ldx + tc];
#pragma unroll
                    for (int u = 0; u < 8; ++u) pr = fma(__shfl_sync(CHB_FULL, alpha, r0 + u), v[u], pr);
                }
            }
            const double df = tin ? pr - xq[t] : 0.0;
            ss = fma(df, df, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) {
            a.dist[pair] = sqrt(ss);
            if (a.status) a.status[pair] = status;
            if (status == CHB_QP_ITER_CAP && a.cap_count) atomicAdd(a.cap_count, 1); // feasible, possibly not optimal: reported
        }
        if (a.alpha && lane < k) a.alpha[pair * k + lane] = lane < m ? alpha : 0.0;
        __syncwarp();
    }
}

template <int KMAX>
int launch(chb_ctx *ctx, const chb_qp_args &a, int blocks_per_sm = 16, int fast = 0)
{
    int64_t blocks = (a.n_work + QP_WARPS - 1) / QP_WARPS;
    const int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    {
        chb_stage_timer t(ctx, CHB_ST_QP);
        CHB_PDL_LAUNCH(ctx, qp_kernel<KMAX>, (unsigned)blocks, QP_WARPS * 32, 0, a, fast);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}

} // namespace

int chb_launch_qp(chb_ctx *ctx, const chb_qp_args &a)
{
    const bool fb_zeroed = ctx->qp_fb_zeroed; // round_reset_kernel zeroed the fallback counter for THIS launch only
    ctx->qp_fb_zeroed = false;
    if (a.n_work <= 0) return CHB_OK;
    CHB_CHECK(ctx, a.k >= 1 && a.k <= CHB_KMAX, CHB_EINVAL, "num_neighbors must be in [1, %d]", CHB_KMAX);
    CHB_CHECK(ctx, a.metric == CHB_METRIC_CONVEX || a.metric == CHB_METRIC_AFFINE_QP || a.metric == CHB_METRIC_AFFINE,
              CHB_ENOTIMPL, "Metric %d not implemented", a.metric);
    const bool lane_path = a.k >= 11 && a.k <= 24 && !getenv("CHB_QP_NO_LANE");
    if ((a.k <= 10 || lane_path) && a.metric == CHB_METRIC_CONVEX) {
        // k <= 5: qp_small.cu; 6..10: 8-lane Gram + one lane per pair (qp_mid.cu); 11..24: tensor-core Gram + one lane per pair
        // (qp_lane.cu).  Pairs whose a'Ga is too small to trust, and pairs that did not end on a verified face, come back here
        if (ctx->fallback_cap < a.n_work) {
            if (ctx->fallback) cudaFree(ctx->fallback);
            ctx->fallback = nullptr;
            cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->fallback), sizeof(int2) * (size_t)a.n_work);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                ctx->fallback_cap = 0;
                return chb_fail(ctx, CHB_ENOMEM, "cudaMalloc of the QP fallback list failed: %s", cudaGetErrorString(e));
            }
            ctx->fallback_cap = a.n_work;
        }
        if (!fb_zeroed) CHB_CUDA(ctx, cudaMemsetAsync(&ctx->counters[3], 0, sizeof(int32_t), ctx->stream));
        int rc = a.k <= 5 ? chb_launch_qp_small(ctx, a, ctx->fallback, &ctx->counters[3])
                          : (a.k <= 10 ? chb_launch_qp_mid(ctx, a, ctx->fallback, &ctx->counters[3])
                                       : chb_launch_qp_lane(ctx, a, ctx->fallback, &ctx->counters[3]));
        if (rc != CHB_OK) return rc;
        if (getenv("CHB_QP_DEBUG")) { // development aid: how many pairs the first-line kernel handed back
            int32_t nfb = 0;
            cudaMemcpyAsync(&nfb, &ctx->counters[3], sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            fprintf(stderr, "[chbin_b200] qp k=%d: %lld pairs (upper bound), %d handed to the general kernel\n", a.k, (long long)a.n_work, nfb);
        }
        chb_qp_args b = a;
        b.work = ctx->fallback;
        b.work_count = &ctx->counters[3];
        if (a.k <= 5) return launch<8>(ctx, b, 1);
        if (a.k <= 10) return launch<16>(ctx, b, 1);
        // low-dimensional inputs put most queries INSIDE the hull (a'Ga = 0): the whole batch may come back
        const int fb_fast = getenv("CHB_QP_NO_TABLEAU") ? 0 : 1;
        return a.k <= 16 ? launch<16>(ctx, b, 16, fb_fast) : launch<24>(ctx, b, 16, fb_fast);
    }
    // main kernel for these neighbour counts: the exchanged tableau is kept across the iterations of a pair
    // fast: 0 = Wolfe with a fresh solve per minor cycle (test aid CHB_QP_NO_TABLEAU), 1 = block principal pivoting on the
    // exchanged tableau (<= 24 neighbours), 2 = Wolfe on the exchanged tableau (25..32 neighbours)
    const int fast = a.metric == CHB_METRIC_CONVEX && !getenv("CHB_QP_NO_TABLEAU");
    if (a.k <= 8) return launch<8>(ctx, a, 16, fast);
    if (a.k <= 16) return launch<16>(ctx, a, 16, fast);
    if (a.k <= 24) return launch<24>(ctx, a, 16, fast); // the cost follows the unrolled width, not k
    return launch<32>(ctx, a, 16, fast ? 2 : 0);
}
