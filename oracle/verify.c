/*
 * oracle/verify.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Position-parallel restatement of ONE step of the reference's assignment loop
 * (/root/reference/ch_bin/core/clustering/algorithm.py:46-60), used to check a claimed result of a whole
 * iteration at sizes where the sequential oracle (fit_cluster_ref.c) would run for many minutes.
 *
 * The sequential loop defines, for the permutation position p (point j = perm[p]),
 *     new[j] = assign(j | labels of the positions < p as AFTER this iteration,
 *                         labels of the positions > p and of the seed contigs as BEFORE it, j itself = -1)
 * (algorithm.py:50: the point leaves its bin; :51-58: C hull distances, strict '<'; :60: visible to the next point).
 * A vector `new` satisfies this equation at EVERY position iff it is the sequential result (induction over p:
 * position 0 sees only old labels; position p sees positions < p, equal by hypothesis).  Each position's equation can
 * be evaluated independently of the others, so the check parallelises over positions -- with exactly the arithmetic of
 * the sequential oracle: scipy's cdist recipe (sequential sum, no FMA), the (distance, index) ranking of
 * distance_matrix.py:47-62 and hull_distance_impl() (hull_distance.py:7-35 through gi_qp.c).
 *
 * Queries are processed in blocks of QB so that one pass over X serves QB positions (the sequential oracle is bound by
 * streaming X once per step); every query's squared-distance sum still runs strictly left to right over the features,
 * the block dimension is only the SIMD lane.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

#define QB 8

/* squared distances of QB queries (transposed: qt[t * QB + lane]) to one point; per lane: ((0 + d0^2) + d1^2) + ... */
__attribute__((target_clones("avx512f", "avx2", "default"))) static void
block_sqdist(const double *restrict qt, const double *restrict x, int32_t d, double *restrict out)
{
    double acc[QB];
    for (int l = 0; l < QB; ++l) acc[l] = 0.0;
    for (int32_t t = 0; t < d; ++t) {
        const double xv = x[t];
        for (int l = 0; l < QB; ++l) {
            const double df = qt[(size_t)t * QB + l] - xv;
            acc[l] += df * df;
        }
    }
    for (int l = 0; l < QB; ++l) out[l] = acc[l];
}

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

/* see oracle.h */
int64_t chb_oracle_verify_positions(const double *X, int64_t n, int32_t d, int32_t C, const int64_t *old_labels,
                                    const int64_t *new_labels, const int64_t *perm, int64_t U, const int64_t *positions,
                                    int64_t npos, int32_t k, int32_t metric, int32_t threads, int64_t *label_out,
                                    double *best_out, double *second_out, double *dist_out)
{
    if (k < 1 || k > 64 || C < 1) return -1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (!pos) return -2;
    for (int64_t i = 0; i < n; ++i) pos[i] = -1; /* seed contigs: never permuted, label fixed (algorithm.py:38) */
    for (int64_t p = 0; p < U; ++p) pos[perm[p]] = (int32_t)p;
    const int64_t nblocks = (npos + QB - 1) / QB;
    int64_t mismatches = 0;
#pragma omp parallel reduction(+ : mismatches)
    {
        double *qt = (double *)malloc(sizeof(double) * (size_t)d * QB);
        double *bd = (double *)malloc(sizeof(double) * (size_t)QB * C * k);
        int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)QB * C * k);
        int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)QB * C);
        int64_t *size = (int64_t *)malloc(sizeof(int64_t) * (size_t)QB * C);
        double *V = (double *)malloc(sizeof(double) * (size_t)k * d);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < nblocks; ++b) {
            const int64_t q0 = b * QB;
            const int nq = (int)((npos - q0) < QB ? (npos - q0) : QB);
            int64_t pq[QB], jq[QB];
            for (int l = 0; l < QB; ++l) {
                pq[l] = positions[q0 + (l < nq ? l : 0)];
                jq[l] = perm[pq[l]];
                for (int32_t t = 0; t < d; ++t) qt[(size_t)t * QB + l] = X[jq[l] * d + t];
            }
            for (int64_t e = 0; e < (int64_t)QB * C; ++e) { cnt[e] = 0; size[e] = 0; }
            for (int64_t i = 0; i < n; ++i) {
                const int32_t pi = pos[i];
                const int64_t lo = old_labels[i], ln = new_labels[i];
                if (lo < 0 && ln < 0) continue; /* visible to nobody */
                double sq[QB];
                block_sqdist(qt, X + i * d, d, sq);
                for (int l = 0; l < nq; ++l) {
                    /* algorithm.py:46-60: earlier positions already carry this iteration's label, later ones last
                     * iteration's; the query itself was taken out of its bin (:50) */
                    const int64_t lab = (pi < 0) ? lo : (pi < pq[l] ? ln : (pi > pq[l] ? lo : -1));
                    if (lab < 0 || lab >= C) continue;
                    const size_t e = (size_t)l * C + (size_t)lab;
                    ++size[e];
                    topm_insert(bd + e * k, bi + e * k, &cnt[e], k, sqrt(sq[l]), i);
                }
            }
            for (int l = 0; l < nq; ++l) {
                double best = INFINITY, second = INFINITY;
                int64_t best_c = old_labels[jq[l]]; /* algorithm.py:48 */
                for (int32_t c = 0; c < C; ++c) {
                    const size_t e = (size_t)l * C + (size_t)c;
                    const int32_t m = cnt[e];
                    double dc = INFINITY;
                    if (m > 0) {
                        int64_t *ids = bi + e * k;
                        if (size[e] <= k) qsort(ids, (size_t)m, sizeof(int64_t), cmp_i64); /* distance_matrix.py:58-59 */
                        for (int32_t a = 0; a < m; ++a) memcpy(V + (size_t)a * d, X + ids[a] * d, sizeof(double) * (size_t)d);
                        int32_t st;
                        dc = (metric == 0) ? chb_oracle_convex_hull_distance(X + jq[l] * d, V, m, d, NULL, &st)
                                           : chb_oracle_affine_hull_distance_qp(X + jq[l] * d, V, m, d, &st);
                    }
                    if (dist_out) dist_out[(q0 + l) * C + c] = dc;
                    if (best > dc) { second = best; best = dc; best_c = c; } /* algorithm.py:57-58 */
                    else if (second > dc) second = dc;
                }
                if (label_out) label_out[q0 + l] = best_c;
                if (best_out) best_out[q0 + l] = best;
                if (second_out) second_out[q0 + l] = second;
                mismatches += (best_c != new_labels[jq[l]]);
            }
        }
        free(qt); free(bd); free(bi); free(cnt); free(size); free(V);
    }
    free(pos);
    return mismatches;
}

/* see oracle.h: hull_distance.py:90-108 for a batch of (query, neighbour list) pairs, OpenMP over the pairs */
void chb_oracle_hull_distance_batch(const double *X, int32_t d, const int64_t *queries, const int64_t *idx, const int32_t *m,
                                    int64_t npairs, int32_t k, int32_t metric, int32_t threads, double *dist_out,
                                    int32_t *status_out)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        double *V = (double *)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1) * d);
#pragma omp for schedule(dynamic, 64)
        for (int64_t p = 0; p < npairs; ++p) {
            const int32_t mm = m[p];
            int32_t st = 3;
            double dv = INFINITY;
            if (mm > 0) {
                for (int32_t a = 0; a < mm; ++a) memcpy(V + (size_t)a * d, X + idx[p * k + a] * d, sizeof(double) * (size_t)d);
                dv = (metric == 0) ? chb_oracle_convex_hull_distance(X + queries[p] * d, V, mm, d, NULL, &st)
                                   : chb_oracle_affine_hull_distance_qp(X + queries[p] * d, V, mm, d, &st);
            }
            dist_out[p] = dv;
            if (status_out) status_out[p] = st;
        }
        free(V);
    }
}
