/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-callable CPU restatement of the reference's clustering hot path
 * (/root/reference/ch_bin/core/clustering/{algorithm,distance_matrix,hull_distance,solve_qp}.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (ch-bin_b200/) never does.
 */
#ifndef CHB_ORACLE_H
#define CHB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* gi_qp.c -- quadprog.solve_qp convention (solve_qp.py:51). 0 ok, 1 inconsistent, 2 not PD, 3 OOM */
int chb_oracle_gi_solve(int n, const double *G, const double *a, int q, const double *Cm, const double *b, int meq,
                        double *x, double *obj, double *lagr, int *iact_out, int *nact_out, int *iters);

/* minnorm.c -- independent exact solver (Wolfe 1976 minimum-norm point) on the same simplex QP.
 * P (m*m row-major, = 2 V V'), qv (m, = -2 V x): minimise 1/2 a'Pa + qv'a, a>=0, sum a = 1. */
int chb_oracle_simplex_qp(int m, const double *P, const double *qv, double *alpha);

/* fit_cluster_ref.c */
/* distance_matrix.py:33-44 -- scipy cdist 'euclidean' recipe: sequential sum of squared differences, sqrt */
void chb_oracle_cdist(const double *X, int64_t n, int32_t d, double *D);
void chb_oracle_cdist_rows(const double *X, int64_t n, int32_t d, const int64_t *rows, int64_t nrows, double *D);

/* distance_matrix.py:47-62 -- returns the count written to idx_out (<= m); canonical order (distance, index)
 * when the bin has more than m members, ascending index otherwise */
int32_t chb_oracle_find_nearest(int32_t c, const int64_t *curr_bins, int64_t n, const double *distance_row, int32_t m,
                                int64_t *idx_out);

/* hull_distance.py:7-35 with solve_qp.py:18-51.  status: 0 GI, 1 GI failed -> min-norm fallback */
double chb_oracle_convex_hull_distance(const double *query, const double *points /* m*d */, int32_t m, int32_t d,
                                       double *alpha_out /* m or NULL */, int32_t *status);
/* hull_distance.py:38-66 (equality-constrained QP only) */
double chb_oracle_affine_hull_distance_qp(const double *query, const double *points, int32_t m, int32_t d,
                                          int32_t *status);

/* algorithm.py:12-76.  perms: max_iterations rows of U int64 (host-drawn np.random.permutation per iteration).
 * dist: optional n*n matrix (NULL = rows computed on the fly with the same recipe).
 * metric: 0 convex, 1 affine-qp.  Returns 0 or a negative error.  iters_run = number of executed iterations,
 * converged = 1 if stopped on "no changes", changed[it] = number of points whose label differs from the
 * previous iteration's (algorithm.py:63-69).  max_steps >=0 bounds the number of sequential steps executed
 * (for timing a bounded sample); -1 = unbounded. */
int chb_oracle_fit_cluster(const double *X, int64_t n, int32_t d, int32_t num_clusters, const int64_t *initial_bins,
                           const double *dist, int32_t num_neighbors, int32_t max_iterations, int32_t metric,
                           const int64_t *perms, int64_t U, int32_t threads, int64_t max_steps, int64_t *labels_out,
                           int32_t *iters_run, int32_t *converged, int64_t *changed, int64_t *qps_solved);

/* verify.c -- position-parallel check of ONE iteration of algorithm.py:46-60.  For each listed permutation position p
 * (point j = perm[p]) evaluates assign(j | new_labels for positions < p, old_labels for positions > p and for seeds) with
 * the arithmetic of chb_oracle_fit_cluster, and compares with new_labels[j].  `new_labels` is the sequential result iff the
 * check passes at every position.  label_out / best_out / second_out (npos, optional): the oracle's label, its smallest
 * and second smallest hull distance; dist_out (npos * C, optional): every hull distance.  Returns the number of
 * positions whose label differs (negative: error). */
int64_t chb_oracle_verify_positions(const double *X, int64_t n, int32_t d, int32_t C, const int64_t *old_labels,
                                    const int64_t *new_labels, const int64_t *perm, int64_t U, const int64_t *positions,
                                    int64_t npos, int32_t k, int32_t metric, int32_t threads, int64_t *label_out,
                                    double *best_out, double *second_out, double *dist_out);

/* verify.c -- calculate_distance (hull_distance.py:90-108) for npairs (query point, neighbour list) pairs: queries (npairs)
 * point indices, idx (npairs * k) neighbour point indices of which the first m[p] count; metric 0 convex, 1 affine-qp.
 * dist_out (npairs; +inf where m == 0), status_out (npairs, optional: 0 GI, 1 fallback solver, 3 empty). */
void chb_oracle_hull_distance_batch(const double *X, int32_t d, const int64_t *queries, const int64_t *idx, const int32_t *m,
                                    int64_t npairs, int32_t k, int32_t metric, int32_t threads, double *dist_out,
                                    int32_t *status_out);

#ifdef __cplusplus
}
#endif
#endif
