// qp_mid.cu -- hull-distance QPs with 6..10 neighbours (BASELINE config #3 uses AlgoNumNeighbors = 10) (sm_100a).
//
// Same contract as qp.cu / qp_small.cu (hull_distance.py:7-35 + solve_qp.py:18-51 + quadprog).  Mapping:
//  * phase 1 (as qp_small.cu): 8 lanes per (query, bin) pair split the feature columns with 16-byte loads, keep
//    W = V - 1x' in registers and accumulate the 55 entries of G = W W' with DFMA; a transposed-halving reduction
//    leaves 7 entries per lane, which go to shared memory.  A warp does this for 32 pairs (8 sub-steps of 4).
//  * phase 2: ONE LANE PER PAIR works on its own G (shared memory, conflict-free [entry][lane] layout): block principal
//    pivoting from the full face (a handful of solves), and Wolfe's finite active-set method as the safety net.  The
//    affine minimiser on a face solves (G_SS + s 11') y = 1 by fully unrolled masked symmetric elimination in
//    registers; vanishing pivots (affinely dependent neighbours) ban the entering vertex.  32 independent solves per
//    warp instead of one warp-cooperative solve per pair (qp.cu).
//  * phase 3: distance = sqrt(a'Ga) when a'Ga > 1e-5 max G_ii, otherwise the pair is handed to qp.cu, which
//    recomputes || aV - x || in d dimensions exactly as hull_distance.py:34-35.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int M = 10;            // maximum neighbours handled here
constexpr int NP = M * (M + 1) / 2; // 55 Gram entries
constexpr int NPP = 56;          // padded to 7 per lane of an 8-lane group
constexpr int GL = 8;
constexpr int QPW = 4;           // pairs per warp and sub-step
constexpr int WARPS = 3;          // 3 x 14 KB of Gram staging stays under the 48 KB static shared-memory limit
constexpr int BPP_ITMAX = 12;

__device__ __forceinline__ constexpr int pidx(int i, int j) { return i * M - (i * (i - 1)) / 2 + (j - i); } // i <= j

// affine minimiser of a'Ga on the face `mask` via (G_SS + shift 11') y = 1; returns false if a pivot vanishes
__device__ __forceinline__ bool affine_min(const double *__restrict__ sg, unsigned mask, double shift, double (&beta)[M])
{
    double A[NP], rhs[M];
    bool in[M];
#pragma unroll
    for (int i = 0; i < M; ++i) in[i] = (mask >> i) & 1u;
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
        for (int j = i; j < M; ++j) A[pidx(i, j)] = (in[i] && in[j]) ? sg[pidx(i, j) * 32] + shift : (i == j ? 1.0 : 0.0);
        rhs[i] = in[i] ? 1.0 : 0.0;
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double p = A[pidx(j, j)];
        const double ref = in[j] ? sg[pidx(j, j) * 32] + shift : 1.0;
        ok = ok && (p > 1e-11 * ref);
        const double inv = 1.0 / p;
#pragma unroll
        for (int i = j + 1; i < M; ++i) {
            const double f = A[pidx(j, i)] * inv;
#pragma unroll
            for (int c = i; c < M; ++c) A[pidx(i, c)] = fma(-f, A[pidx(j, c)], A[pidx(i, c)]);
            rhs[i] = fma(-f, rhs[j], rhs[i]);
        }
        rhs[j] *= inv;
#pragma unroll
        for (int c = j + 1; c < M; ++c) A[pidx(j, c)] *= inv;
    }
    double y[M];
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
        double v = rhs[i];
#pragma unroll
        for (int c = i + 1; c < M; ++c) v = fma(-A[pidx(i, c)], y[c], v);
        y[i] = v;
    }
    double sy = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) sy += y[i];
    const double isy = 1.0 / sy;
#pragma unroll
    for (int i = 0; i < M; ++i) beta[i] = in[i] ? y[i] * isy : 0.0;
    return ok && (sy > 0.0);
}

__global__ void __launch_bounds__(WARPS * 32) qp_mid_kernel(chb_qp_args a, int2 *__restrict__ fallback,
                                                             int32_t *__restrict__ fallback_count)
{
    chb_pdl_wait();
    __shared__ __align__(16) double sG[WARPS][NPP * 32]; // [entry][pair-in-warp-batch]
    __shared__ int sI[WARPS][32][12];                     // per pair of the batch: m, query point, up to 10 neighbour points
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & (GL - 1), grp = lane / GL;
    const int64_t n_work = a.work_count ? (int64_t)*a.work_count : a.n_work;
    const int ldx = a.ldx, k = a.k, C = a.C;
    const int64_t stride = (int64_t)gridDim.x * WARPS * 32;
    double *sg_w = sG[warp];

    for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * 32; base < n_work; base += stride) {
        // ---------------- phase 0: lane l fetches the indices of pair l of the batch (work item -> neighbour count ->
        // neighbour points: three dependent loads, paid once per batch instead of once per sub-step) and keeps what phase 2
        // needs in its own registers
        const int64_t my_item = base + lane;
        const bool my_valid = my_item < n_work;
        int2 my_wk = make_int2(0, 0);
        int64_t my_pair = 0;
        int my_m = 0;
        {
            int qpt = 0;
            if (my_valid) {
                my_wk = a.work[my_item];
                my_pair = (int64_t)my_wk.x * C + my_wk.y;
                my_m = a.knn_cnt[my_pair];
                qpt = a.row_point[my_wk.x];
            }
            sI[warp][lane][0] = my_m;
            sI[warp][lane][1] = qpt;
#pragma unroll
            for (int r = 0; r < M; ++r) sI[warp][lane][2 + r] = (my_valid && r < my_m) ? a.knn_idx[my_pair * k + r] : qpt; // r >= m: W row = 0
        }
        __syncwarp();
        // ---------------- phase 1: Gram matrices of 32 pairs, 4 at a time
#pragma unroll 1
        for (int sub = 0; sub < 8; ++sub) {
            if (base + (int64_t)sub * QPW >= n_work) break; // warp-uniform
            const int slot = sub * QPW + grp; // pair index inside the warp batch = the lane that will solve it
            const int m = sI[warp][slot][0];
            const double *xq = a.X + (int64_t)sI[warp][slot][1] * ldx;
            const double *rows[M];
#pragma unroll
            for (int r = 0; r < M; ++r) rows[r] = a.X + (int64_t)sI[warp][slot][2 + r] * ldx;
            double acc[NPP];
#pragma unroll
            for (int i = 0; i < NPP; ++i) acc[i] = 0.0;
            for (int c0 = 0; c0 < ldx; c0 += 2 * GL) {
                // a column beyond the row is replaced by column 0 of the QUERY row for every row: w - x = 0 there, exactly as
                // for the rows r >= m, which point at the query row altogether
                const int col = c0 + 2 * g;
                const bool inb = col < ldx;
                const double2 xv = __ldg(reinterpret_cast<const double2 *>(xq + (inb ? col : 0)));
                double2 w[M];
#pragma unroll
                for (int r = 0; r < M; ++r) {
                    w[r] = __ldg(reinterpret_cast<const double2 *>((inb ? rows[r] : xq) + (inb ? col : 0)));
                    w[r].x -= xv.x;
                    w[r].y -= xv.y;
                }
#pragma unroll
                for (int p = 0; p < M; ++p)
#pragma unroll
                    for (int q = p; q < M; ++q) {
                        acc[pidx(p, q)] = fma(w[p].x, w[q].x, acc[pidx(p, q)]);
                        acc[pidx(p, q)] = fma(w[p].y, w[q].y, acc[pidx(p, q)]);
                    }
            }
            // transposed halving inside the 8-lane group: 56 -> 28 -> 14 -> 7 entries per lane; lane g ends with 7g..7g+6
            double v28[28], v14[14], v7[7];
#pragma unroll
            for (int i = 0; i < 28; ++i) {
                const bool hi = g & 4;
                const double send = hi ? acc[i] : acc[i + 28];
                const double keep = hi ? acc[i + 28] : acc[i];
                v28[i] = keep + __shfl_xor_sync(CHB_FULL, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                const bool hi = g & 2;
                const double send = hi ? v28[i] : v28[i + 14];
                const double keep = hi ? v28[i + 14] : v28[i];
                v14[i] = keep + __shfl_xor_sync(CHB_FULL, send, 2);
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const bool hi = g & 1;
                const double send = hi ? v14[i] : v14[i + 7];
                const double keep = hi ? v14[i + 7] : v14[i];
                v7[i] = keep + __shfl_xor_sync(CHB_FULL, send, 1);
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) sg_w[(7 * g + i) * 32 + slot] = v7[i];
        }
        __syncwarp();

        // ---------------- phase 2: one lane per pair
        const bool valid = my_valid;
        const int m = my_m;
        const int64_t pair = my_pair;
        const int2 wk = my_wk;
        const double *sg = sg_w + lane;
        double alpha[M];
#pragma unroll
        for (int i = 0; i < M; ++i) alpha[i] = 0.0;
        double best = 0.0, scale = 0.0;
        int status = CHB_QP_OK;
        if (m > 0) {
            int start = 0;
            double dmin = DBL_MAX;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const double gii = sg[pidx(i, i) * 32];
                if (i < m) {
                    scale = fmax(scale, gii);
                    if (gii < dmin) { dmin = gii; start = i; }
                }
            }
            if (scale > 0.0) {
                const double tol = 1e-14 * scale;
                // Block principal pivoting from the FULL face first: solve on S, take out every vertex with a negative weight,
                // put back every excluded vertex with a negative multiplier, repeat.  It ends on the KKT conditions of the
                // simplex problem after a handful of solves whatever m is, where Wolfe's method below adds one vertex per
                // major cycle.  No feasibility is kept on the way, so it is only trusted when it ends by itself within the
                // cap; otherwise (cycling, or a vanishing pivot: affinely dependent neighbours) Wolfe's method runs.
                bool done = false;
                {
                    const unsigned full = (1u << m) - 1u;
                    unsigned S = full;
#pragma unroll 1
                    for (int bit = 0; bit < BPP_ITMAX && !done; ++bit) {
                        double beta[M];
                        if (!affine_min(sg, S, scale, beta)) break;
                        double gr[M];
#pragma unroll
                        for (int i = 0; i < M; ++i) gr[i] = 0.0;
#pragma unroll
                        for (int i = 0; i < M; ++i)
#pragma unroll
                            for (int j = i; j < M; ++j) {
                                const double gij = sg[pidx(i, j) * 32];
                                gr[i] = fma(gij, beta[j], gr[i]);
                                if (j != i) gr[j] = fma(gij, beta[i], gr[j]);
                            }
                        double f = 0.0;
#pragma unroll
                        for (int i = 0; i < M; ++i) f = fma(beta[i], gr[i], f);
                        unsigned neg = 0u, dual = 0u;
#pragma unroll
                        for (int i = 0; i < M; ++i) {
                            if ((S >> i) & 1u) {
                                if (beta[i] < 0.0) neg |= 1u << i;
                            } else if (((full >> i) & 1u) && gr[i] < f - tol) {
                                dual |= 1u << i;
                            }
                        }
                        if (!neg && !dual) {
#pragma unroll
                            for (int i = 0; i < M; ++i) alpha[i] = beta[i];
                            done = true;
                        } else {
                            S = (S & ~neg) | dual;
                        }
                    }
                }
                if (!done) {
#pragma unroll
                    for (int i = 0; i < M; ++i) alpha[i] = (i == start) ? 1.0 : 0.0;
                }
                unsigned smask = 1u << start, banned = 0u;
                const int itmax = done ? 0 : 3 * m + 8;
                int it = 0;
                for (; it < itmax; ++it) {
                    // gradient g = G alpha, f = alpha' g
                    double gr[M];
#pragma unroll
                    for (int i = 0; i < M; ++i) gr[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int j = i; j < M; ++j) {
                            const double gij = sg[pidx(i, j) * 32];
                            gr[i] = fma(gij, alpha[j], gr[i]);
                            if (j != i) gr[j] = fma(gij, alpha[i], gr[j]);
                        }
                    double f = 0.0;
#pragma unroll
                    for (int i = 0; i < M; ++i) f = fma(alpha[i], gr[i], f);
                    best = f;
                    int jn = -1;
                    double gmin = f - tol;
#pragma unroll
                    for (int i = 0; i < M; ++i)
                        if (i < m && !((smask >> i) & 1u) && !((banned >> i) & 1u) && gr[i] < gmin) { gmin = gr[i]; jn = i; }
                    if (jn < 0) break; // optimal
                    smask |= 1u << jn;
                    for (int minor = 0; minor <= m; ++minor) {
                        double beta[M];
                        if (!affine_min(sg, smask, scale, beta)) {
                            smask &= ~(1u << jn);
                            banned |= 1u << jn;
                            status = CHB_QP_DEGENERATE;
                            break;
                        }
                        bool anyneg = false;
#pragma unroll
                        for (int i = 0; i < M; ++i) anyneg = anyneg || (((smask >> i) & 1u) && !(beta[i] > 0.0));
                        if (!anyneg) {
#pragma unroll
                            for (int i = 0; i < M; ++i) alpha[i] = beta[i];
                            break;
                        }
                        double th = 1.0;
                        int lt = -1;
#pragma unroll
                        for (int i = 0; i < M; ++i)
                            if (((smask >> i) & 1u) && !(beta[i] > 0.0)) {
                                const double t = (alpha[i] > 0.0) ? alpha[i] / (alpha[i] - beta[i]) : 0.0;
                                if (t < th || lt < 0) { th = t; lt = i; }
                            }
                        th = fmin(fmax(th, 0.0), 1.0);
                        double sa = 0.0;
#pragma unroll
                        for (int i = 0; i < M; ++i) {
                            const bool ins = (smask >> i) & 1u;
                            double v = ins ? alpha[i] + th * (beta[i] - alpha[i]) : 0.0;
                            if (ins && (i == lt || !(v > 0.0))) {
                                v = 0.0;
                                smask &= ~(1u << i);
                                if (i == jn) banned |= 1u << jn;
                            }
                            alpha[i] = v;
                            sa += v;
                        }
                        const double isa = 1.0 / sa;
#pragma unroll
                        for (int i = 0; i < M; ++i) alpha[i] *= isa;
                        if (!((smask >> jn) & 1u)) break;
                    }
                }
                if (!done && it >= itmax) status = CHB_QP_ITER_CAP;
                // objective at the final alpha
                double o = 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    double r = 0.5 * sg[pidx(i, i) * 32] * alpha[i];
#pragma unroll
                    for (int j = i + 1; j < M; ++j) r = fma(sg[pidx(i, j) * 32], alpha[j], r);
                    o = fma(alpha[i], r, o);
                }
                best = 2.0 * o;
            } else {
                alpha[0] = 1.0; // every neighbour coincides with the query
                best = 0.0;
            }
        }
        if (valid) {
            if (m <= 0) {
                a.dist[pair] = INFINITY;
                if (a.status) a.status[pair] = CHB_QP_EMPTY_BIN;
            } else {
                const bool exact_needed = !(best == best) || (scale > 0.0 && !(best > 1e-5 * scale) && best != 0.0) ||
                                          status == CHB_QP_ITER_CAP;
                if (exact_needed) {
                    const int w = atomicAdd(fallback_count, 1);
                    fallback[w] = wk;
                } else {
                    a.dist[pair] = sqrt(fmax(best, 0.0));
                    if (a.status) a.status[pair] = status;
                    if (a.alpha) {
#pragma unroll
                        for (int i = 0; i < M; ++i)
                            if (i < k) a.alpha[pair * k + i] = alpha[i];
                    }
                }
            }
        }
        __syncwarp();
    }
}

} // namespace

int chb_launch_qp_mid(chb_ctx *ctx, const chb_qp_args &a, int2 *fallback, int32_t *fallback_count)
{
    int64_t blocks = (a.n_work + WARPS * 32 - 1) / (WARPS * 32);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    {
        chb_stage_timer t(ctx, CHB_ST_QP);
        CHB_PDL_LAUNCH(ctx, qp_mid_kernel, (unsigned)blocks, WARPS * 32, 0, a, fallback, fallback_count);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}
