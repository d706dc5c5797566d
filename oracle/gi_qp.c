/*
 * oracle/gi_qp.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the dual active-set method of Goldfarb & Idnani (1983) for
 * strictly convex QPs, with the calling convention of the third-party package the
 * reference calls at /root/reference/ch_bin/core/clustering/solve_qp.py:51
 * (`quadprog.solve_qp(G, a, C, b, meq)`, quadprog==0.1.8 pinned in
 * /root/reference/requirements.txt:7).  quadprog's source is NOT vendored in the
 * reference and the wheel is absent from this image, so this file restates the
 * published algorithm (B. Turlach's qpgen2 formulation of GI: Cholesky start at the
 * unconstrained minimiser, most-violated-constraint selection weighted by the column
 * norm, partial/full steps, Givens updates of J and R).
 *
 *      minimise   1/2 x'Gx - a'x     subject to   C'x >= b,
 *      the first `meq` constraints being equalities.
 *
 * PARITY STATUS: "unpinned" against quadprog itself (it cannot be imported here);
 * pinned against KKT certificates, closed forms and an independent solver
 * (oracle/minnorm.c, SLSQP) in tests/test_oracle_qp.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* Smallest v (by doubling from 1e-60) with 1 + 0.1 v > 1 and 1 + 0.2 v > 1:
 * the "vsmall" guard the qpgen2 formulation uses for every is-it-zero test. */
static double gi_vsmall(void)
{
    static double cached = 0.0;
    if (cached == 0.0) {
        volatile double v = 1e-60, ta, tb;
        do {
            v = v + v;
            ta = 1.0 + 0.1 * v;
            tb = 1.0 + 0.2 * v;
        } while (ta <= 1.0 || tb <= 1.0);
        cached = v;
    }
    return cached;
}

/* Givens pair (c,s) that maps (p,q) -> (r,0); returns r. */
static inline double givens(double p, double q, double *c, double *s)
{
    double h = hypot(p, q);
    if (h == 0.0) { *c = 1.0; *s = 0.0; return 0.0; }
    *c = p / h;
    *s = q / h;
    return h;
}

/* rotate columns (j0,j1) of the n x n column-major matrix J */
static inline void rot_cols(double *J, int n, int j0, int j1, double c, double s)
{
    double *u = J + (size_t)j0 * n, *v = J + (size_t)j1 * n;
    for (int i = 0; i < n; ++i) {
        double a = u[i], b = v[i];
        u[i] = c * a + s * b;
        v[i] = -s * a + c * b;
    }
}

/*
 * n      : number of variables
 * G      : n*n, row-major, symmetric; NOT modified
 * a      : n
 * q      : number of constraints
 * Cm     : n*q row-major, column j is the normal of constraint j (quadprog's `C`)
 * b      : q
 * x      : out, n
 * obj    : out, objective value at x (may be NULL)
 * lagr   : out, q multipliers (may be NULL)
 * iact   : out, up to q active constraint ids (may be NULL); nact out (may be NULL)
 * iters  : out, [0]=constraints added, [1]=constraints dropped (may be NULL)
 * return : 0 ok, 1 constraints inconsistent, 2 G not positive definite
 */
int chb_oracle_gi_solve(int n, const double *G, const double *a, int q, const double *Cm, const double *b, int meq,
                        double *x, double *obj, double *lagr, int *iact_out, int *nact_out, int *iters)
{
    const double vsmall = gi_vsmall();
    int rc = 0;
    size_t nn = (size_t)n * n;
    double *mem = (double *)malloc(sizeof(double) * (nn * 2 + (size_t)n * q + 4 * (size_t)n + 4 * (size_t)q + 8));
    int *A = (int *)malloc(sizeof(int) * (q + 1));
    if (!mem || !A) { free(mem); free(A); return 3; }
    double *J = mem;                  /* n x n column-major: J J' = G^-1 */
    double *R = J + nn;               /* n x n column-major upper triangular, R(i,j) at R[j*n+i] */
    double *N = R + nn;               /* q normals, each contiguous (n) -- sign may flip for equalities */
    double *d = N + (size_t)n * q;    /* n */
    double *z = d + n;                /* n */
    double *r = z + n;                /* n */
    double *u = r + n;                /* n+1: multipliers of the active set (+ candidate) */
    double *bv = u + n + 1;           /* q */
    double *s = bv + q;               /* q slacks */
    double *nrm = s + q;              /* q column norms */
    int nact = 0, n_add = 0, n_drop = 0;

    /* --- Cholesky G = L L' (lower, stored column-major in R temporarily) --- */
    /* A pivot at rounding level means G is singular to working precision (exact duplicate contigs among the neighbours make
     * two rows of 2 V V' identical): whether such a pivot comes out as +1e-17 or -1e-17 is noise, and factoring through it
     * returns garbage (1e32-size "solutions") with a success code.  The reference's designed route for a non-PD matrix is
     * "ValueError -> fallback solver" (solve_qp.py:110-123); the restatement takes that route deterministically: a pivot
     * below n * 64 * eps * max|diag| counts as not positive definite. */
    double *L = R;
    double gmax = 0.0;
    for (int j = 0; j < n; ++j) gmax = fmax(gmax, fabs(G[(size_t)j * n + j]));
    const double ptol = (double)n * 64.0 * 2.220446049250313e-16 * gmax;
    for (int j = 0; j < n; ++j) {
        double dj = G[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) dj -= L[(size_t)k * n + j] * L[(size_t)k * n + j];
        if (!(dj > ptol)) { rc = 2; goto done; }
        dj = sqrt(dj);
        L[(size_t)j * n + j] = dj;
        for (int i = j + 1; i < n; ++i) {
            double v = G[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) v -= L[(size_t)k * n + i] * L[(size_t)k * n + j];
            L[(size_t)j * n + i] = v / dj;
        }
    }
    /* J = L^-T (upper triangular): column j of J solves L' J(:,j) = e_j */
    memset(J, 0, sizeof(double) * nn);
    for (int j = 0; j < n; ++j) {
        double *col = J + (size_t)j * n;
        for (int i = j; i >= 0; --i) {
            double v = (i == j) ? 1.0 : 0.0;
            for (int k = i + 1; k <= j; ++k) v -= L[(size_t)i * n + k] * col[k];
            col[i] = v / L[(size_t)i * n + i];
        }
    }
    /* x = J J' a  (unconstrained minimiser) */
    for (int j = 0; j < n; ++j) {
        double t = 0.0;
        const double *col = J + (size_t)j * n;
        for (int i = 0; i <= j; ++i) t += col[i] * a[i];
        d[j] = t;
    }
    for (int i = 0; i < n; ++i) x[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double *col = J + (size_t)j * n;
        for (int i = 0; i <= j; ++i) x[i] += col[i] * d[j];
    }
    double crval = 0.0;
    for (int i = 0; i < n; ++i) crval -= 0.5 * a[i] * x[i];
    memset(R, 0, sizeof(double) * nn);

    for (int j = 0; j < q; ++j) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) {
            double v = Cm[(size_t)i * q + j];
            N[(size_t)j * n + i] = v;
            t += v * v;
        }
        nrm[j] = sqrt(t);
        bv[j] = b[j];
    }
    for (int i = 0; i <= n; ++i) u[i] = 0.0;

    for (;;) {
        /* --- slack of every constraint at x; equalities count as violated either way --- */
        for (int j = 0; j < q; ++j) {
            double *nj = N + (size_t)j * n;
            double t = -bv[j];
            for (int i = 0; i < n; ++i) t += nj[i] * x[i];
            if (fabs(t) < vsmall) t = 0.0;
            if (j >= meq) {
                s[j] = t;
            } else {
                s[j] = -fabs(t);
                if (t > 0.0) {
                    for (int i = 0; i < n; ++i) nj[i] = -nj[i];
                    bv[j] = -bv[j];
                }
            }
        }
        for (int i = 0; i < nact; ++i) s[A[i]] = 0.0;
        int p = -1;
        double worst = 0.0;
        for (int j = 0; j < q; ++j)
            if (s[j] < worst * nrm[j]) { p = j; worst = s[j] / nrm[j]; }
        if (p < 0) break; /* all constraints satisfied: optimal */

        for (;;) { /* step 2(a): repeated after each drop, with the same p */
            const double *np = N + (size_t)p * n;
            for (int j = 0; j < n; ++j) {
                const double *col = J + (size_t)j * n;
                double t = 0.0;
                for (int i = 0; i < n; ++i) t += col[i] * np[i];
                d[j] = t;
            }
            for (int i = 0; i < n; ++i) z[i] = 0.0;
            for (int j = nact; j < n; ++j) {
                const double *col = J + (size_t)j * n;
                for (int i = 0; i < n; ++i) z[i] += col[i] * d[j];
            }
            /* r = R^-1 d_1; the blocking candidate is the active INEQUALITY minimising u/r over r>0 */
            int l = -1;
            double t1 = INFINITY;
            for (int i = nact - 1; i >= 0; --i) {
                double t = d[i];
                for (int k = i + 1; k < nact; ++k) t -= R[(size_t)k * n + i] * r[k];
                r[i] = t / R[(size_t)i * n + i];
            }
            for (int i = 0; i < nact; ++i) {
                if (A[i] < meq || !(r[i] > 0.0)) continue;
                double t = u[i] / r[i];
                if (t < t1) { t1 = t; l = i; }
            }
            double zz = 0.0;
            for (int i = 0; i < n; ++i) zz += z[i] * z[i];
            int full_step = 0;
            if (fabs(zz) <= vsmall) {
                /* no primal direction: dual step only, then drop l */
                if (l < 0) { rc = 1; goto done; }
                for (int i = 0; i < nact; ++i) u[i] -= t1 * r[i];
                u[nact] += t1;
            } else {
                double zn = 0.0;
                for (int i = 0; i < n; ++i) zn += z[i] * np[i];
                double t = -s[p] / zn;
                full_step = 1;
                if (l >= 0 && t1 < t) { t = t1; full_step = 0; }
                for (int i = 0; i < n; ++i) x[i] += t * z[i];
                crval += t * zn * (0.5 * t + u[nact]);
                for (int i = 0; i < nact; ++i) u[i] -= t * r[i];
                u[nact] += t;
            }
            if (full_step) {
                /* add p: new column of R is d after rotating d[nact+1..n-1] into d[nact] */
                for (int j = n - 1; j > nact; --j) {
                    if (d[j] == 0.0) continue;
                    double c, sn;
                    d[j - 1] = givens(d[j - 1], d[j], &c, &sn);
                    d[j] = 0.0;
                    rot_cols(J, n, j - 1, j, c, sn);
                }
                for (int i = 0; i <= nact; ++i) R[(size_t)nact * n + i] = d[i];
                A[nact] = p;
                ++nact;
                ++n_add;
                break; /* back to the violated-constraint search */
            }
            /* partial step: recompute the slack of p, drop active constraint l, retry p */
            if (fabs(zz) > vsmall) {
                double t = -bv[p];
                double *npm = N + (size_t)p * n;
                for (int i = 0; i < n; ++i) t += npm[i] * x[i];
                if (p >= meq) {
                    s[p] = t;
                } else {
                    s[p] = -fabs(t);
                    if (t > 0.0) {
                        for (int i = 0; i < n; ++i) npm[i] = -npm[i];
                        bv[p] = -bv[p];
                    }
                }
            }
            /* remove column l of R, restore triangular form with row rotations mirrored on J's columns */
            for (int j = l; j < nact - 1; ++j) {
                double *dst = R + (size_t)j * n, *src = R + (size_t)(j + 1) * n;
                for (int i = 0; i <= j + 1; ++i) dst[i] = src[i];
                A[j] = A[j + 1];
                u[j] = u[j + 1];
            }
            u[nact - 1] = u[nact];
            u[nact] = 0.0;
            for (int i = 0; i < n; ++i) R[(size_t)(nact - 1) * n + i] = 0.0;
            --nact;
            ++n_drop;
            for (int j = l; j < nact; ++j) {
                double *col = R + (size_t)j * n;
                if (col[j + 1] == 0.0) continue;
                double c, sn;
                col[j] = givens(col[j], col[j + 1], &c, &sn);
                col[j + 1] = 0.0;
                for (int k = j + 1; k < nact; ++k) {
                    double *ck = R + (size_t)k * n;
                    double p0 = ck[j], p1 = ck[j + 1];
                    ck[j] = c * p0 + sn * p1;
                    ck[j + 1] = -sn * p0 + c * p1;
                }
                rot_cols(J, n, j, j + 1, c, sn);
            }
        }
    }

    if (obj) *obj = crval;
    if (lagr) {
        for (int j = 0; j < q; ++j) lagr[j] = 0.0;
        for (int i = 0; i < nact; ++i) lagr[A[i]] = u[i];
    }
    if (iact_out) for (int i = 0; i < nact; ++i) iact_out[i] = A[i];
    if (nact_out) *nact_out = nact;
done:
    if (iters) { iters[0] = n_add; iters[1] = n_drop; }
    free(mem);
    free(A);
    return rc;
}
