// qp_small.cu -- hull-distance QPs with at most 5 neighbours (AlgoNumNeighbors <= 5, the reference's default,
// /root/reference/config/default.ini:16), 8 lanes per (query, bin) pair, 4 pairs per warp (sm_100a).
//
// Same contract as qp.cu (hull_distance.py:7-35 + solve_qp.py:18-51 + quadprog), different mapping:
//  * phase 1: the 8 lanes of a group split the d feature columns (16-byte loads, one full 128-byte line per
//    neighbour row and step), keep W = V - 1x' in registers and accumulate the 15 entries of G = W W' with
//    DFMA; a transposed-halving reduction (28 SHFL per warp) plus one shared-memory broadcast gives every
//    lane the whole G.
//  * phase 2: the simplex-constrained minimum of a'Ga is attained on a face whose affine minimiser is
//    non-negative (KKT).  With m <= 5 there are at most 31 faces: each lane takes 4 of them, solves
//    (G_SS + s 11') y = 1 by fully unrolled masked elimination, normalises, rejects faces with a negative
//    weight or a vanishing pivot (affinely dependent neighbours), and EVALUATES a'Ga directly -- every
//    surviving candidate is a feasible point, so the minimum over candidates can never undershoot the true
//    optimum, and the optimal face attains it.  No iteration, no divergence.
//  * phase 3: the distance is sqrt(a'Ga) when that is well conditioned (a'Ga > 1e-5 max G_ii, error ~1e-11
//    relative); otherwise the pair is handed to the general kernel, which recomputes ||aV - x|| in d
//    dimensions exactly as hull_distance.py:34-35.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int GL = 8;        // lanes per QP
constexpr int QPW = 32 / GL; // QPs per warp
constexpr int WARPS = 4;
constexpr int NCH = 3;                // double2 loads per row per chunk and lane
constexpr int CHUNK_COLS = GL * 2 * NCH; // 48 columns per chunk

__device__ __forceinline__ constexpr int pidx(int i, int j) { return i * 5 - (i * (i - 1)) / 2 + (j - i); } // i <= j

// affine minimiser on the face `mask`, returns false if the face is rejected
__device__ __forceinline__ bool eval_face(const double (&G)[15], unsigned mask, double shift, double (&beta)[5], double &obj)
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
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        beta[i] = in[i] ? y[i] * isy : 0.0;
        ok = ok && (beta[i] >= 0.0);
    }
    // objective evaluated on the feasible point itself
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

__global__ void __launch_bounds__(WARPS * 32) qp_small_kernel(chb_qp_args a, int2 *__restrict__ fallback,
                                                               int32_t *__restrict__ fallback_count)
{
    __shared__ __align__(16) double sG[WARPS][QPW][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & (GL - 1), grp = lane / GL;
    const int64_t n_work = a.work_count ? (int64_t)*a.work_count : a.n_work;
    const int ldx = a.ldx, k = a.k, C = a.C;
    const int64_t stride = (int64_t)gridDim.x * WARPS * QPW;

    for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * QPW; base < n_work; base += stride) {
        const int64_t item = base + grp;
        const bool valid = item < n_work;
        int2 wk = make_int2(0, 0);
        int m = 0;
        int64_t pair = 0;
        if (valid) {
            wk = a.work[item];
            pair = (int64_t)wk.x * C + wk.y;
            m = a.knn_cnt[pair];
        }
        const double *rows[5];
        const double *xq = a.X;
        if (m > 0) xq = a.X + (int64_t)a.row_point[wk.x] * ldx;
#pragma unroll
        for (int r = 0; r < 5; ++r) rows[r] = (r < m) ? a.X + (int64_t)a.knn_idx[pair * k + r] * ldx : a.X;

        // ---------------- phase 1
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0;
        for (int c0 = 0; c0 < ldx; c0 += CHUNK_COLS) {
            double2 xv[NCH], w[5][NCH];
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int col = c0 + i * (2 * GL) + 2 * g;
                const bool inb = (col < ldx) && (m > 0);
                xv[i] = inb ? __ldg(reinterpret_cast<const double2 *>(xq + col)) : make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < 5; ++r)
                    w[r][i] = (inb && r < m) ? __ldg(reinterpret_cast<const double2 *>(rows[r] + col)) : xv[i];
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
        *reinterpret_cast<double2 *>(&sG[warp][grp][2 * g]) = make_double2(v2[0], v2[1]);
        __syncwarp();
        double G[15];
#pragma unroll
        for (int i = 0; i < 14; i += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(&sG[warp][grp][i]);
            G[i] = t.x;
            G[i + 1] = t.y;
        }
        G[14] = sG[warp][grp][14];
        __syncwarp();

        // ---------------- phase 2: faces g+1, g+9, g+17, g+25 (< 2^m)
        double scale = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) scale = fmax(scale, G[pidx(i, i)]);
        double best = DBL_MAX, bbeta[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        const unsigned nfaces = 1u << m;
        if (scale > 0.0) {
#pragma unroll 1
            for (unsigned mask = g + 1; mask < nfaces; mask += GL) {
                double beta[5], obj;
                if (eval_face(G, mask, scale, beta, obj) && obj < best) {
                    best = obj;
#pragma unroll
                    for (int i = 0; i < 5; ++i) bbeta[i] = beta[i];
                }
            }
        } else if (m > 0 && g == 0) {
            best = 0.0; // every neighbour coincides with the query
            bbeta[0] = 1.0;
        }
        // group argmin (lowest lane wins ties)
        double bv = best;
        int bl = g;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(CHB_FULL, bv, o);
            const int ol = __shfl_xor_sync(CHB_FULL, bl, o);
            if (ov < bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
        }
        if (!valid) continue;
        if (m <= 0) {
            if (g == 0) {
                a.dist[pair] = INFINITY;
                if (a.status) a.status[pair] = CHB_QP_EMPTY_BIN;
            }
            continue;
        }
        const bool exact_needed = !(bv < DBL_MAX) || (scale > 0.0 && !(bv > 1e-5 * scale));
        if (g == bl) {
            if (exact_needed && bv != 0.0) {
                const int w = atomicAdd(fallback_count, 1);
                fallback[w] = wk;
            } else {
                a.dist[pair] = sqrt(fmax(bv, 0.0));
                if (a.status) a.status[pair] = CHB_QP_OK;
                if (a.alpha) {
                    for (int i = 0; i < k; ++i) a.alpha[pair * k + i] = i < 5 ? bbeta[i] : 0.0;
                }
            }
        }
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
        qp_small_kernel<<<(unsigned)blocks, WARPS * 32, 0, ctx->stream>>>(a, fallback, fallback_count);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    return CHB_OK;
}
