/*
 * oracle/fit_cluster_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Sequential CPU restatement of the reference's clustering hot path:
 *   chb_oracle_cdist            <- /root/reference/ch_bin/core/clustering/distance_matrix.py:33-44 (scipy cdist)
 *   chb_oracle_find_nearest     <- distance_matrix.py:47-62
 *   chb_oracle_convex_hull_...  <- hull_distance.py:7-35 + solve_qp.py:18-51,96-132
 *   chb_oracle_fit_cluster      <- algorithm.py:12-76
 *
 * Build with -ffp-contract=off: the cdist recipe is "subtract, multiply, add, left to right, then sqrt"
 * with no fused multiply-add (checked bit-for-bit against scipy in tests/test_oracle_golden.py).
 *
 * Deviations from the reference, all documented in DESIGN.md:
 *  - nearest_positive_definite (positive_def.py:25-48) is not applied: it perturbs a well-conditioned Gram
 *    matrix by <=3e-15 relative; the verbatim reference flow used to generate tests/golden
 *    DOES apply it, and the two agree to <=1e-10 relative in tests/test_oracle_qp.py.
 *  - when GI reports "not positive definite"/"inconsistent" the reference falls back to cvxopt
 *    (solve_qp.py:126-129); here the fallback is oracle/minnorm.c.
 *  - OpenMP may be used across the C bins of ONE sequential step (legal: they are independent).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

static inline double row_dist(const double *a, const double *b, int32_t d)
{
    double s = 0.0;
    for (int32_t t = 0; t < d; ++t) {
        double df = a[t] - b[t];
        s += df * df;
    }
    return sqrt(s);
}

void chb_oracle_cdist(const double *X, int64_t n, int32_t d, double *D)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) D[i * n + j] = row_dist(X + i * d, X + j * d, d);
}

void chb_oracle_cdist_rows(const double *X, int64_t n, int32_t d, const int64_t *rows, int64_t nrows, double *D)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < nrows; ++r) {
        const double *a = X + rows[r] * d;
        for (int64_t j = 0; j < n; ++j) D[r * n + j] = row_dist(a, X + j * d, d);
    }
}

/* keep the m smallest (distance, index) pairs, sorted ascending */
static inline void topm_insert(double *bd, int64_t *bi, int32_t *cnt, int32_t m, double dv, int64_t iv)
{
    int32_t c = *cnt;
    if (c == m) {
        if (dv > bd[m - 1] || (dv == bd[m - 1] && iv > bi[m - 1])) return;
        c = m - 1;
    }
    int32_t p = c;
    while (p > 0 && (bd[p - 1] > dv || (bd[p - 1] == dv && bi[p - 1] > iv))) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
    }
    bd[p] = dv;
    bi[p] = iv;
    *cnt = c + 1;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

int32_t chb_oracle_find_nearest(int32_t c, const int64_t *curr_bins, int64_t n, const double *distance_row, int32_t m,
                                int64_t *idx_out)
{
    double *bd = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    int32_t cnt = 0;
    int64_t members = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (curr_bins[i] != c) continue;
        ++members;
        if (m > 0) topm_insert(bd, idx_out, &cnt, m, distance_row[i], i);
    }
    free(bd);
    if (members <= m) qsort(idx_out, (size_t)cnt, sizeof(int64_t), cmp_i64);
    return cnt;
}

static double hull_distance_impl(const double *query, const double *points, int32_t m, int32_t d, int with_ineq,
                                 double *alpha_out, int32_t *status)
{
    enum { KM = 64 };
    double P[KM * KM], qa[KM], Cm[KM * (KM + 1)], bv[KM + 1], alpha[KM];
    if (status) *status = 0;
    if (m <= 0 || m > KM) { if (status) *status = 3; return INFINITY; }
    for (int32_t i = 0; i < m; ++i) {
        for (int32_t j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int32_t t = 0; t < d; ++t) s += points[(size_t)i * d + t] * points[(size_t)j * d + t];
            P[i * m + j] = P[j * m + i] = 2.0 * s; /* hull_distance.py:30 */
        }
        double s = 0.0;
        for (int32_t t = 0; t < d; ++t) s += query[t] * points[(size_t)i * d + t];
        qa[i] = 2.0 * s; /* qp_a = -vec_q = 2 V x  (hull_distance.py:31, solve_qp.py:46) */
    }
    int q = with_ineq ? m + 1 : 1;
    /* qp_C = -[A; G]' = [-1 | +I],  qp_b = -[b; h] = [-1, 0..0], meq = 1  (solve_qp.py:47-49) */
    for (int32_t i = 0; i < m; ++i) {
        Cm[i * q + 0] = -1.0;
        if (with_ineq)
            for (int32_t j = 0; j < m; ++j) Cm[i * q + 1 + j] = (i == j) ? 1.0 : 0.0;
    }
    bv[0] = -1.0;
    for (int32_t j = 1; j < q; ++j) bv[j] = 0.0;
    int rc = chb_oracle_gi_solve(m, P, qa, q, Cm, bv, 1, alpha, NULL, NULL, NULL, NULL, NULL);
    if (rc != 0) {
        if (!with_ineq) {
            /* affine metrics with affinely dependent vertices (duplicate contigs): the affine hull is the same without
             * them and the distance stays well defined -- hull_distance.py:64-87 gets it from an SVD basis
             * (scipy.linalg.orth).  Restated rank-revealing: modified Gram-Schmidt over the edge vectors v_i - v_0,
             * skipping the ones that are already in the span, then the residual of x - v_0. */
            enum { DM = 1024 };
            if (d > DM) { if (status) *status = 2; return NAN; }
            static _Thread_local double Q[KM][DM];
            double u[DM], r[DM];
            int nb = 0;
            for (int32_t i = 1; i < m; ++i) {
                double un2 = 0.0, rn2 = 0.0;
                for (int32_t t = 0; t < d; ++t) { u[t] = points[(size_t)i * d + t] - points[t]; un2 += u[t] * u[t]; }
                for (int pass = 0; pass < 2; ++pass) /* twice is enough */
                    for (int b = 0; b < nb; ++b) {
                        double dot = 0.0;
                        for (int32_t t = 0; t < d; ++t) dot += Q[b][t] * u[t];
                        for (int32_t t = 0; t < d; ++t) u[t] -= dot * Q[b][t];
                    }
                for (int32_t t = 0; t < d; ++t) rn2 += u[t] * u[t];
                if (!(rn2 > 1e-22 * un2)) continue;
                const double inv = 1.0 / sqrt(rn2);
                for (int32_t t = 0; t < d; ++t) Q[nb][t] = u[t] * inv;
                ++nb;
            }
            for (int32_t t = 0; t < d; ++t) r[t] = query[t] - points[t];
            for (int pass = 0; pass < 2; ++pass)
                for (int b = 0; b < nb; ++b) {
                    double dot = 0.0;
                    for (int32_t t = 0; t < d; ++t) dot += Q[b][t] * r[t];
                    for (int32_t t = 0; t < d; ++t) r[t] -= dot * Q[b][t];
                }
            double ss = 0.0;
            for (int32_t t = 0; t < d; ++t) ss += r[t] * r[t];
            if (status) *status = 1;
            return sqrt(ss);
        }
        double qv[KM];
        for (int32_t i = 0; i < m; ++i) qv[i] = -qa[i];
        chb_oracle_simplex_qp(m, P, qv, alpha);
        if (status) *status = 1;
    }
    /* proj = alpha @ points ; norm(proj - x)   (hull_distance.py:34-35) */
    double ss = 0.0;
    for (int32_t t = 0; t < d; ++t) {
        double pt = 0.0;
        for (int32_t i = 0; i < m; ++i) pt += alpha[i] * points[(size_t)i * d + t];
        double df = pt - query[t];
        ss += df * df;
    }
    if (alpha_out) memcpy(alpha_out, alpha, sizeof(double) * (size_t)m);
    return sqrt(ss);
}

double chb_oracle_convex_hull_distance(const double *query, const double *points, int32_t m, int32_t d,
                                       double *alpha_out, int32_t *status)
{
    return hull_distance_impl(query, points, m, d, 1, alpha_out, status);
}

double chb_oracle_affine_hull_distance_qp(const double *query, const double *points, int32_t m, int32_t d,
                                          int32_t *status)
{
    return hull_distance_impl(query, points, m, d, 0, NULL, status);
}

int chb_oracle_fit_cluster(const double *X, int64_t n, int32_t d, int32_t num_clusters, const int64_t *initial_bins,
                           const double *dist, int32_t num_neighbors, int32_t max_iterations, int32_t metric,
                           const int64_t *perms, int64_t U, int32_t threads, int64_t max_steps, int64_t *labels_out,
                           int32_t *iters_run, int32_t *converged, int64_t *changed, int64_t *qps_solved)
{
    const int32_t C = num_clusters, k = num_neighbors;
    if (k < 1 || k > 64 || C < 1) return -1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    int64_t *prev = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *curr = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    double *row = dist ? NULL : (double *)malloc(sizeof(double) * (size_t)n);
    double *bd = (double *)malloc(sizeof(double) * (size_t)C * k);
    int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)C * k);
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)C);
    int64_t *size = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    double *dc = (double *)malloc(sizeof(double) * (size_t)C);
    double *vbuf = (double *)malloc(sizeof(double) * (size_t)C * k * d);
    memcpy(prev, initial_bins, sizeof(int64_t) * (size_t)n);
    memcpy(curr, initial_bins, sizeof(int64_t) * (size_t)n);
    int64_t steps = 0, nqp = 0;
    int32_t it = 0, conv = 0, stop = 0;
    for (it = 0; it < max_iterations && !stop; ++it) {
        const int64_t *perm = perms + (int64_t)it * U;
        for (int64_t s = 0; s < U; ++s) {
            if (max_steps >= 0 && steps >= max_steps) { stop = 1; break; }
            ++steps;
            int64_t j = perm[s];
            int64_t best_c = curr[j]; /* algorithm.py:48 */
            double best = INFINITY;
            curr[j] = -1; /* algorithm.py:50 */
            const double *drow;
            if (dist) {
                drow = dist + j * n;
            } else {
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < n; ++i) row[i] = row_dist(X + j * d, X + i * d, d);
                drow = row;
            }
            for (int32_t c = 0; c < C; ++c) { cnt[c] = 0; size[c] = 0; }
            for (int64_t i = 0; i < n; ++i) {
                int64_t c = curr[i];
                if (c < 0 || c >= C) continue;
                ++size[c];
                topm_insert(bd + (size_t)c * k, bi + (size_t)c * k, &cnt[c], k, drow[i], i);
            }
#pragma omp parallel for schedule(dynamic, 1)
            for (int32_t c = 0; c < C; ++c) {
                int32_t m = cnt[c];
                if (m == 0) { dc[c] = INFINITY; continue; }
                int64_t *ids = bi + (size_t)c * k;
                if (size[c] <= k) qsort(ids, (size_t)m, sizeof(int64_t), cmp_i64); /* distance_matrix.py:58-59 */
                double *V = vbuf + (size_t)c * k * d;
                for (int32_t a = 0; a < m; ++a) memcpy(V + (size_t)a * d, X + ids[a] * d, sizeof(double) * (size_t)d);
                int32_t st;
                dc[c] = hull_distance_impl(X + j * d, V, m, d, metric == 0, NULL, &st);
            }
            nqp += C;
            for (int32_t c = 0; c < C; ++c)
                if (best > dc[c]) { best = dc[c]; best_c = c; } /* algorithm.py:57-58: strict, NaN never wins */
            curr[j] = best_c; /* algorithm.py:60 */
        }
        if (stop) break;
        int64_t nch = 0;
        for (int64_t i = 0; i < n; ++i) nch += (prev[i] != curr[i]);
        if (changed) changed[it] = nch;
        if (nch == 0) { conv = 1; ++it; break; } /* algorithm.py:63-66 */
        memcpy(prev, curr, sizeof(int64_t) * (size_t)n); /* algorithm.py:71-72 */
    }
    memcpy(labels_out, curr, sizeof(int64_t) * (size_t)n);
    if (iters_run) *iters_run = it;
    if (converged) *converged = conv;
    if (qps_solved) *qps_solved = nqp;
    free(prev); free(curr); free(row); free(bd); free(bi); free(cnt); free(size); free(dc); free(vbuf);
    return 0;
}
