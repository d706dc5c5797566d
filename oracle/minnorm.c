/*
 * oracle/minnorm.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Independent exact solver for the one QP shape the reference ever builds
 * (/root/reference/ch_bin/core/clustering/hull_distance.py:19-33):
 *
 *      minimise 1/2 a'Pa + qv'a   subject to  a >= 0, 1'a = 1,   P = 2 V V', qv = -2 V x.
 *
 * On the simplex this equals a'Ha with H = P/2 + (qv 1' + 1 qv')/2, i.e. the squared norm of the
 * point sum a_i (v_i - x) up to the constant |x|^2, so it is Wolfe's (1976) minimum-norm-point
 * problem.  The implementation below is Wolfe's corral method with the affine minimiser obtained
 * in null-space form (differences to the first corral vertex), re-factored from scratch at every
 * minor cycle -- slow, simple and tolerant of affinely dependent points.
 *
 * Roles: (1) cross-check of oracle/gi_qp.c in tests; (2) stand-in for the reference's
 * `cvxopt` fallback (solve_qp.py:126-129) when GI reports a non-positive-definite matrix:
 * cvxopt 1.2.6 (requirements.txt:8) is absent from this image and is an interior-point code with
 * ~1e-7 accuracy, so this stand-in is MORE exact than what it replaces; stated in DESIGN.md.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

#define MN_MAX 64

/* affine minimiser of a'Ha on the support sup[0..ns): returns 0, or 1 if the reduced matrix is singular */
static int affine_min(int m, const double *H, const int *sup, int ns, double scale, double *beta)
{
    if (ns == 1) { beta[0] = 1.0; return 0; }
    int r = ns - 1;
    double M[MN_MAX * MN_MAX], rhs[MN_MAX];
    int i0 = sup[0];
    for (int a = 0; a < r; ++a) {
        int ia = sup[a + 1];
        for (int b = 0; b < r; ++b) {
            int ib = sup[b + 1];
            M[a * r + b] = H[ia * m + ib] - H[ia * m + i0] - H[i0 * m + ib] + H[i0 * m + i0];
        }
        rhs[a] = -(H[ia * m + i0] - H[i0 * m + i0]);
    }
    /* Cholesky in place (lower) */
    for (int j = 0; j < r; ++j) {
        double dj = M[j * r + j];
        for (int k = 0; k < j; ++k) dj -= M[j * r + k] * M[j * r + k];
        if (!(dj > 1e-13 * scale)) return 1;
        dj = sqrt(dj);
        M[j * r + j] = dj;
        for (int i = j + 1; i < r; ++i) {
            double v = M[i * r + j];
            for (int k = 0; k < j; ++k) v -= M[i * r + k] * M[j * r + k];
            M[i * r + j] = v / dj;
        }
    }
    for (int i = 0; i < r; ++i) {
        double v = rhs[i];
        for (int k = 0; k < i; ++k) v -= M[i * r + k] * rhs[k];
        rhs[i] = v / M[i * r + i];
    }
    for (int i = r - 1; i >= 0; --i) {
        double v = rhs[i];
        for (int k = i + 1; k < r; ++k) v -= M[k * r + i] * rhs[k];
        rhs[i] = v / M[i * r + i];
    }
    double s = 0.0;
    for (int a = 0; a < r; ++a) { beta[a + 1] = rhs[a]; s += rhs[a]; }
    beta[0] = 1.0 - s;
    return 0;
}

int chb_oracle_simplex_qp(int m, const double *P, const double *qv, double *alpha)
{
    if (m < 1 || m > MN_MAX) return 2;
    double H[MN_MAX * MN_MAX];
    double scale = 0.0;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            double h = 0.25 * (P[i * m + j] + P[j * m + i]) + 0.5 * (qv[i] + qv[j]);
            H[i * m + j] = h;
        }
    /* scale of the translation-invariant part: largest squared edge length |v_i - v_j|^2 and spread of the diagonal */
    double dmin = H[0];
    int imin = 0;
    for (int i = 0; i < m; ++i) {
        if (H[i * m + i] < dmin) { dmin = H[i * m + i]; imin = i; }
        for (int j = 0; j < i; ++j) {
            double e = H[i * m + i] + H[j * m + j] - 2.0 * H[i * m + j];
            if (e > scale) scale = e;
        }
    }
    for (int i = 0; i < m; ++i) {
        double e = fabs(H[i * m + i] - dmin);
        if (e > scale) scale = e;
    }
    if (scale == 0.0) scale = fabs(dmin) > 0 ? fabs(dmin) : 1.0;
    const double tol = 1e-14 * scale;

    int sup[MN_MAX], ns = 1, in_sup[MN_MAX];
    double beta[MN_MAX], g[MN_MAX];
    memset(in_sup, 0, sizeof(in_sup));
    for (int i = 0; i < m; ++i) alpha[i] = 0.0;
    alpha[imin] = 1.0;
    sup[0] = imin;
    in_sup[imin] = 1;

    for (int major = 0; major < 20 * m + 20; ++major) {
        double f = 0.0;
        for (int i = 0; i < m; ++i) {
            double t = 0.0;
            for (int j = 0; j < m; ++j) t += H[i * m + j] * alpha[j];
            g[i] = t;
        }
        for (int i = 0; i < m; ++i) f += alpha[i] * g[i];
        int jn = -1;
        double gmin = f - tol;
        for (int i = 0; i < m; ++i)
            if (!in_sup[i] && g[i] < gmin) { gmin = g[i]; jn = i; }
        if (jn < 0) return 0;
        sup[ns++] = jn;
        in_sup[jn] = 1;
        for (int minor = 0; minor < 2 * m + 4; ++minor) {
            if (affine_min(m, H, sup, ns, scale, beta)) {
                /* affinely dependent corral: discard the newest vertex and stop improving along it */
                in_sup[sup[ns - 1]] = 2; /* 2 = banned for this solve */
                --ns;
                break;
            }
            int ok = 1;
            for (int a = 0; a < ns; ++a)
                if (!(beta[a] > 0.0)) ok = 0;
            if (ok) {
                for (int i = 0; i < m; ++i) alpha[i] = 0.0;
                for (int a = 0; a < ns; ++a) alpha[sup[a]] = beta[a];
                break;
            }
            double theta = 1.0;
            for (int a = 0; a < ns; ++a) {
                double ai = alpha[sup[a]];
                if (!(beta[a] > 0.0)) {
                    double t = ai / (ai - beta[a]);
                    if (t < theta) theta = t;
                }
            }
            if (theta < 0.0) theta = 0.0;
            int w = 0, removed = 0;
            double amin = INFINITY;
            int amin_a = -1;
            for (int a = 0; a < ns; ++a) {
                double v = (1.0 - theta) * alpha[sup[a]] + theta * beta[a];
                alpha[sup[a]] = v;
                if (!(beta[a] > 0.0) && v < amin) { amin = v; amin_a = a; }
            }
            for (int a = 0; a < ns; ++a) {
                int i = sup[a];
                if (a == amin_a || alpha[i] <= 0.0) {
                    alpha[i] = 0.0;
                    in_sup[i] = 0;
                    ++removed;
                } else {
                    sup[w++] = i;
                }
            }
            ns = w;
            (void)removed;
            double s = 0.0;
            for (int a = 0; a < ns; ++a) s += alpha[sup[a]];
            for (int a = 0; a < ns; ++a) alpha[sup[a]] /= s;
        }
    }
    return 0; /* iteration cap: alpha is feasible, possibly not optimal to full precision */
}
