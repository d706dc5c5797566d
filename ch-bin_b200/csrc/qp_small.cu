// qp_small.cu -- hull-distance QPs with at most 5 neighbours (AlgoNumNeighbors <= 5, the reference's default,
// /root/reference/config/default.ini:16) (sm_100a).
//
// Same contract as qp.cu (hull_distance.py:7-35 + solve_qp.py:18-51 + quadprog), different mapping:
//  * phase 1: 8 lanes per (query, bin) pair, 4 pairs per warp and sub-step.  The 8 lanes split the d feature columns
//    (16-byte loads, one full 128-byte line per neighbour row and step), keep W = V - 1x' in registers and accumulate
//    the 15 entries of G = W W' with DFMA; a transposed-halving reduction (28 SHFL per warp) leaves two entries per
//    lane, which go to shared memory in an [entry][pair] layout.  A warp repeats this for up to 8 sub-steps (32 pairs);
//    the number of sub-steps follows the size of the work list so that small lists still fill the machine.
//  * phase 2: ONE LANE PER PAIR.  The simplex-constrained minimum of a'Ga is attained on the face S whose affine
//    minimiser is non-negative and whose excluded vertices have non-negative multipliers (KKT).  Block principal
//    pivoting from the full face: solve (G_SS + s 11') y = 1 by fully unrolled masked elimination, take out every
//    vertex with a negative weight, put back every excluded vertex whose multiplier is negative, repeat.  On contig
//    data the full face is already optimal for ~97 % of the pairs (high-dimensional noise keeps the projection inside
//    the simplex), one more solve settles nearly all others.  Whatever the path, a'Ga is EVALUATED on a feasible point,
//    so it can never undershoot the optimum.  If the pivoting does not settle within its cap, or meets a vanishing
//    pivot (affinely dependent neighbours: duplicate contigs), the lane falls back to enumerating all <= 31 faces
//    (each rejected if a pivot vanishes or a weight is negative) -- exact, no iteration.
//  * phase 3: the distance is sqrt(a'Ga) when that is well conditioned (a'Ga > 1e-5 max G_ii, error ~1e-11
//    relative); otherwise the pair is handed to the general kernel, which recomputes ||aV - x|| in d
//    dimensions exactly as hull_distance.py:34-35.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int GL = 8;        // lanes per QP in phase 1
constexpr int QPW = 32 / GL; // QPs per warp and sub-step
constexpr int MAXSUB = 8;    // sub-steps per warp batch: up to 32 pairs, one per lane in phase 2
constexpr int WARPS = 4;
constexpr int NCH = 3;                // double2 loads per row per chunk and lane
constexpr int CHUNK_COLS = GL * 2 * NCH; // 48 columns per chunk
constexpr int BPP_ITMAX = 8;

__device__ __forceinline__ constexpr int pidx(int i, int j) { return i * 5 - (i * (i - 1)) / 2 + (j - i); } // i <= j

// affine minimiser on the face `mask`: beta (0 outside the face), obj = beta' G beta.  Returns false if a pivot vanishes
// (affinely dependent vertices) or the solve breaks down; `feasible` says whether every weight is >= 0.
__device__ __forceinline__ bool eval_face(const double (&G)[15], unsigned mask, double shift, double (&beta)[5], double &obj,
                                          bool &feasible)
{
    double A[15], rhs[5];
    bool in[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) in[i] = (mask >> i) & 1u;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = i; j < 5; ++j) A[pidx(i, j)] = (in[i] && in[j]) ? G[pidx(i, j)] + shift : (i == j ? 1.0 : 0.0);
        rhs[i] = in[i] ? 1.0 : 0.0;
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double p = A[pidx(j, j)];
        const double ref = in[j] ? G[pidx(j, j)] + shift : 1.0;
        ok = ok && (p > 1e-11 * ref);
        const double inv = 1.0 / p;
#pragma unroll
        for (int i = j + 1; i < 5; ++i) {
            const double f = A[pidx(j, i)] * inv;
#pragma unroll
            for (int c = i; c < 5; ++c) A[pidx(i, c)] = fma(-f, A[pidx(j, c)], A[pidx(i, c)]);
            rhs[i] = fma(-f, rhs[j], rhs[i]);
        }
        rhs[j] *= inv; // row j scaled: A[j][c] / p is applied in the back substitution below
#pragma unroll
        for (int c = j + 1; c < 5; ++c) A[pidx(j, c)] *= inv;
    }
    // back substitution on the unit-diagonal upper factor
    double y[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        double v = rhs[i];
#pragma unroll
        for (int c = i + 1; c < 5; ++c) v = fma(-A[pidx(i, c)], y[c], v);
        y[i] = v;
    }
    const double sy = y[0] + y[1] + y[2] + y[3] + y[4];
    const double isy = 1.0 / sy;
    ok = ok && (sy > 0.0);
    feasible = true;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        beta[i] = in[i] ? y[i] * isy : 0.0;
        feasible = feasible && (beta[i] >= 0.0);
    }
    // objective evaluated on the point itself
    double o = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double r = 0.5 * G[pidx(i, i)] * beta[i];
#pragma unroll
        for (int j = i + 1; j < 5; ++j) r = fma(G[pidx(i, j)], beta[j], r);
        o = fma(beta[i], r, o);
    }
    obj = 2.0 * o;
    return ok && (obj == obj);
}

__global__ void __launch_bounds__(WARPS * 32, 4) qp_small_kernel(chb_qp_args a, int2 *__restrict__ fallback,
                                                               int32_t *__restrict__ fallback_count)
{
    chb_pdl_wait();
    __shared__ __align__(16) double sG[WARPS][16 * 32]; // [entry][pair of the warp batch]
    __shared__ int sI[WARPS][32][8];                     // per pair of the batch: m, query point, 5 neighbour points
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & (GL - 1), grp = lane / GL;
    const int64_t n_work = a.work_count ? (int64_t)*a.work_count : a.n_work;
    const int ldx = a.ldx, k = a.k, C = a.C;
    double *sg_w = sG[warp];
    // pairs per warp batch: as many sub-steps as keep every warp of the grid busy (phase 2 then runs on 4 * nsub lanes)
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    int nsub = (int)((n_work + nwarps * QPW - 1) / (nwarps * QPW));
    nsub = nsub < 1 ? 1 : (nsub >= MAXSUB ? MAXSUB : (nsub >= 4 ? 4 : (nsub >= 2 ? 2 : 1)));
    const int per_batch = nsub * QPW;
    const int64_t stride = nwarps * per_batch;

    for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * per_batch; base < n_work; base += stride) {
        // ---------------- phase 0: lane l fetches the indices of pair l of the batch (work item -> neighbour count ->
        // neighbour points: three dependent loads, paid once per batch instead of once per sub-step) and keeps what phase 2
        // needs in its own registers
        const int64_t my_item = base + lane;
        const bool my_valid = lane < per_batch && my_item < n_work;
        int2 my_wk = make_int2(0, 0);
        int64_t my_pair = 0;
        int my_m = 0;
        {
            int qpt = 0, nb[5] = {0, 0, 0, 0, 0};
            if (my_valid) {
                my_wk = a.work[my_item];
                my_pair = (int64_t)my_wk.x * C + my_wk.y;
                my_m = a.knn_cnt[my_pair];
                qpt = a.row_point[my_wk.x];
#pragma unroll
                for (int r = 0; r < 5; ++r) nb[r] = (r < my_m) ? a.knn_idx[my_pair * k + r] : qpt; // r >= m: the query itself, W row = 0
            }
            sI[warp][lane][0] = my_m;
            sI[warp][lane][1] = qpt;
#pragma unroll
            for (int r = 0; r < 5; ++r) sI[warp][lane][2 + r] = nb[r];
        }
        __syncwarp();
        // ---------------- phase 1: Gram matrices of 4 * nsub pairs, 4 at a time
#pragma unroll 1
        for (int sub = 0; sub < nsub; ++sub) {
            if (base + (int64_t)sub * QPW >= n_work) break; // warp-uniform
            const int slot = sub * QPW + grp; // pair index inside the warp batch = the lane that solves it
            const int m = sI[warp][slot][0];
            const double *xq = a.X + (int64_t)sI[warp][slot][1] * ldx;
            const double *rows[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) rows[r] = a.X + (int64_t)sI[warp][slot][2 + r] * ldx;

            double acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.0;
            for (int c0 = 0; c0 < ldx; c0 += CHUNK_COLS) {
                double2 xv[NCH], w[5][NCH];
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    // a column beyond the row (last chunk) is replaced by column 0 of the QUERY row for every row: w - x = 0
                    // there, exactly as for the rows r >= m, which point at the query row altogether
                    const int col = c0 + i * (2 * GL) + 2 * g;
                    const bool inb = col < ldx;
                    xv[i] = __ldg(reinterpret_cast<const double2 *>(xq + (inb ? col : 0)));
#pragma unroll
                    for (int r = 0; r < 5; ++r) w[r][i] = __ldg(reinterpret_cast<const double2 *>((inb ? rows[r] : xq) + (inb ? col : 0)));
                }
#pragma unroll
                for (int i = 0; i < NCH; ++i)
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        w[r][i].x -= xv[i].x;
                        w[r][i].y -= xv[i].y;
                    }
#pragma unroll
                for (int i = 0; i < NCH; ++i)
#pragma unroll
                    for (int p = 0; p < 5; ++p)
#pragma unroll
                        for (int q = p; q < 5; ++q) {
                            acc[pidx(p, q)] = fma(w[p][i].x, w[q][i].x, acc[pidx(p, q)]);
                            acc[pidx(p, q)] = fma(w[p][i].y, w[q][i].y, acc[pidx(p, q)]);
                        }
            }
            // transposed halving reduction inside the 8-lane group: lane g ends with entries 2g, 2g+1
            double v8[8], v4[4], v2[2];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool hi = g & 4;
                const double send = hi ? acc[i] : acc[i + 8];
                const double keep = hi ? acc[i + 8] : acc[i];
                v8[i] = keep + __shfl_xor_sync(CHB_FULL, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool hi = g & 2;
                const double send = hi ? v8[i] : v8[i + 4];
                const double keep = hi ? v8[i + 4] : v8[i];
                v4[i] = keep + __shfl_xor_sync(CHB_FULL, send, 2);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool hi = g & 1;
                const double send = hi ? v4[i] : v4[i + 2];
                const double keep = hi ? v4[i + 2] : v4[i];
                v2[i] = keep + __shfl_xor_sync(CHB_FULL, send, 1);
            }
            sg_w[(2 * g) * 32 + slot] = v2[0];
            sg_w[(2 * g + 1) * 32 + slot] = v2[1];
        }
        __syncwarp();

        // ---------------- phase 2: one lane per pair
        if (my_valid) {
            const int2 wk = my_wk;
            const int64_t pair = my_pair;
            const int m = my_m;
            if (m <= 0) {
                a.dist[pair] = INFINITY;
                if (a.status) a.status[pair] = CHB_QP_EMPTY_BIN;
            } else {
                double G[15];
#pragma unroll
                for (int i = 0; i < 15; ++i) G[i] = sg_w[i * 32 + lane];
                double scale = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) scale = fmax(scale, G[pidx(i, i)]);
                double best = DBL_MAX, bbeta[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                if (scale > 0.0) {
                    const unsigned full = (1u << m) - 1u;
                    const double tol = 1e-14 * scale;
                    // block principal pivoting from the full face
                    unsigned S = full;
                    bool done = false;
#pragma unroll 1
                    for (int it = 0; it < BPP_ITMAX && !done; ++it) {
                        double beta[5], obj;
                        bool feas;
                        if (!eval_face(G, S, scale, beta, obj, feas)) break; // vanishing pivot: enumerate
                        unsigned neg = 0u, dual = 0u;
                        // multipliers of the excluded vertices: (G beta)_i - beta' G beta >= 0
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if ((S >> i) & 1u) {
                                if (beta[i] < 0.0) neg |= 1u << i;
                            } else if ((full >> i) & 1u) {
                                double gi = 0.0;
#pragma unroll
                                for (int j = 0; j < 5; ++j) gi = fma(G[i <= j ? pidx(i, j) : pidx(j, i)], beta[j], gi);
                                if (gi < obj - tol) dual |= 1u << i;
                            }
                        }
                        if (!neg && !dual) {
                            best = obj;
#pragma unroll
                            for (int i = 0; i < 5; ++i) bbeta[i] = beta[i];
                            done = true;
                        } else {
                            S = (S & ~neg) | dual;
                        }
                    }
                    if (!done) {
                        // exact fallback: every face; feasible candidates can only overshoot, the optimal face attains
#pragma unroll 1
                        for (unsigned mask = 1; mask <= full; ++mask) {
                            double beta[5], obj;
                            bool feas;
                            if (eval_face(G, mask, scale, beta, obj, feas) && feas && obj < best) {
                                best = obj;
#pragma unroll
                                for (int i = 0; i < 5; ++i) bbeta[i] = beta[i];
                            }
                        }
                    }
                } else {
                    best = 0.0; // every neighbour coincides with the query
                    bbeta[0] = 1.0;
                }
                const bool exact_needed = !(best < DBL_MAX) || (scale > 0.0 && !(best > 1e-5 * scale));
                if (exact_needed && best != 0.0) {
                    const int w = atomicAdd(fallback_count, 1);
                    fallback[w] = wk;
                } else {
                    a.dist[pair] = sqrt(fmax(best, 0.0));
                    if (a.status) a.status[pair] = CHB_QP_OK;
                    if (a.alpha) {
                        for (int i = 0; i < k; ++i) a.alpha[pair * k + i] = i < 5 ? bbeta[i] : 0.0;
                    }
                }
            }
        }
        __syncwarp();
    }
}

} // namespace

int chb_launch_qp_small(chb_ctx *ctx, const chb_qp_args &a, int2 *fallback, int32_t *fallback_count)
{
    int64_t blocks = (a.n_work + WARPS * QPW - 1) / (WARPS * QPW);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    {
        chb_stage_timer t(ctx, CHB_ST_QP);
        CHB_PDL_LAUNCH(ctx, qp_small_kernel, (unsigned)blocks, WARPS * 32, 0, a, fallback, fallback_count);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}
