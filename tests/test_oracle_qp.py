"""
CPU: solver-independent known-answer tests for the oracle's QP: KKT certificate, closed forms,
simplex projection, an independent solver (Wolfe min-norm, oracle/minnorm.c) and scipy SLSQP, quadprog's calling
convention (return tuple, ValueError text), degenerate inputs.
"""
import numpy as np
import pytest

import oracle


def _problem(rng, m, d, scale=1.0):
    V = rng.dirichlet(np.ones(d), size=m) * scale
    x = rng.dirichlet(np.ones(d)) * scale
    return x, V


def _kkt_ok(x, V, alpha, tol=1e-10):
    g = V @ (V.T @ alpha) - V @ x          # gradient / 2 of ||x - V'a||^2
    assert abs(alpha.sum() - 1) <= 1e-12
    assert alpha.min() >= -1e-14
    act = alpha > 1e-12
    mu = g[act].mean()
    sc = np.abs(g).max() + 1e-300
    assert np.all(np.abs(g[act] - mu) <= tol * sc)
    assert np.all(g[~act] >= mu - tol * sc)


@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 10, 16, 32])
def test_kkt_certificate_and_cross_solvers(m):
    rng = np.random.default_rng(m)
    for _ in range(20):
        x, V = _problem(rng, m, 137)
        d, alpha, st = oracle.convex_hull_distance(x, V, return_alpha=True)
        assert st == 0
        _kkt_ok(x, V, alpha)
        assert abs(d - np.linalg.norm(alpha @ V - x)) <= 1e-15 + 1e-12 * d
        a2 = oracle.simplex_qp(2 * V @ V.T, -2 * V @ x)
        d2 = np.linalg.norm(a2 @ V - x)
        assert abs(d - d2) <= 1e-10 * d


def test_closed_forms():
    rng = np.random.default_rng(1)
    for _ in range(50):
        x, V = _problem(rng, 2, 40)
        assert abs(oracle.convex_hull_distance(x, V[:1]) - np.linalg.norm(x - V[0])) < 1e-15
        t = np.clip(np.dot(x - V[0], V[1] - V[0]) / np.dot(V[1] - V[0], V[1] - V[0]), 0, 1)
        ref = np.linalg.norm(V[0] + t * (V[1] - V[0]) - x)
        assert abs(oracle.convex_hull_distance(x, V) - ref) <= 1e-12 * ref


def test_point_inside_hull_and_permutation_invariance():
    rng = np.random.default_rng(2)
    for m in (3, 6, 12):
        _, V = _problem(rng, m, 30)
        w = rng.dirichlet(np.ones(m))
        x = w @ V
        assert oracle.convex_hull_distance(x, V) <= 1e-9 * np.linalg.norm(x)
        x2, _ = _problem(rng, 1, 30)
        d0 = oracle.convex_hull_distance(x2, V)
        p = rng.permutation(m)
        assert abs(oracle.convex_hull_distance(x2, V[p]) - d0) <= 1e-12 * d0


def test_simplex_projection():
    # V = I_k: the QP is the Euclidean projection of x onto the simplex (sort-and-threshold algorithm)
    rng = np.random.default_rng(4)
    for k in (3, 5, 9):
        x = rng.normal(size=k)
        u = np.sort(x)[::-1]
        css = np.cumsum(u)
        rho = np.nonzero(u * np.arange(1, k + 1) > (css - 1))[0][-1]
        theta = (css[rho] - 1) / (rho + 1.0)
        proj = np.maximum(x - theta, 0)
        d, alpha, _ = oracle.convex_hull_distance(x, np.eye(k), return_alpha=True)
        assert np.allclose(alpha, proj, atol=1e-12)
        assert abs(d - np.linalg.norm(proj - x)) < 1e-12


def test_against_slsqp():
    from scipy.optimize import minimize

    rng = np.random.default_rng(5)
    for m in (4, 7):
        x, V = _problem(rng, m, 25)
        f = lambda a: np.sum((a @ V - x) ** 2)
        r = minimize(f, np.full(m, 1 / m), method="SLSQP", bounds=[(0, 1)] * m,
                     constraints=[{"type": "eq", "fun": lambda a: a.sum() - 1}], options={"ftol": 1e-15, "maxiter": 500})
        d = oracle.convex_hull_distance(x, V)
        assert abs(d - np.sqrt(r.fun)) <= 1e-6 * d


def test_quadprog_convention():
    G = np.array([[4.0, 1.0], [1.0, 3.0]])
    a = np.array([1.0, 2.0])
    C = np.array([[1.0, 1.0, 0.0], [1.0, 0.0, 1.0]])
    b = np.array([1.0, 0.0, 0.0])
    x, f, xu, it, lagr, iact = oracle.quadprog_solve_qp(G, a, C, b, meq=1)
    assert np.allclose(xu, np.linalg.solve(G, a))
    assert abs(x.sum() - 1) < 1e-12 and np.all(x >= -1e-14)
    assert abs(f - (0.5 * x @ G @ x - a @ x)) < 1e-12
    with pytest.raises(ValueError, match="positive definite"):
        oracle.quadprog_solve_qp(np.array([[1.0, 2.0], [2.0, 1.0]]), a, C, b, 1)
    with pytest.raises(ValueError, match="inconsistent"):
        oracle.quadprog_solve_qp(G, a, np.array([[1.0, -1.0], [0.0, 0.0]]), np.array([1.0, 1.0]), 0)


def test_duplicates_fall_back():
    rng = np.random.default_rng(6)
    x, V = _problem(rng, 4, 20)
    V2 = np.vstack([V, V[1], V[1]])
    d0 = oracle.convex_hull_distance(x, V)
    d1, alpha, st = oracle.convex_hull_distance(x, V2, return_alpha=True)
    assert abs(d0 - d1) <= 1e-9 * d0
    assert abs(alpha.sum() - 1) < 1e-9 and alpha.min() >= -1e-12


@pytest.mark.parametrize("m,ndup", [(24, 4), (32, 6), (16, 2), (8, 1)])
def test_duplicates_never_factor_through_a_noise_pivot(m, ndup):
    """Exact duplicate vertices make 2 V V' singular; whether the Cholesky pivot comes out as +1e-17 or -1e-17 is rounding
    noise.  A factorisation that accepts the positive case returns garbage (distances of 1e32 were seen at m = 24) with a
    success code, so the restatement treats rounding-level pivots as "not positive definite" and takes the reference's
    designed route for that (ValueError -> fallback solver, solve_qp.py:110-123).  The distance is still the unique
    projection distance: equal to the one on the de-duplicated vertex set."""
    for seed in range(40):
        rng = np.random.default_rng(1000 * m + seed)
        V = rng.dirichlet(np.full(137, 8.0), size=m - ndup)
        x = rng.dirichlet(np.full(137, 8.0))
        src = rng.choice(m - ndup, ndup, replace=False)
        order = rng.permutation(m)
        V2 = np.vstack([V, V[src]])[order]
        d0 = oracle.convex_hull_distance(x, V)
        d1, alpha, st = oracle.convex_hull_distance(x, V2, return_alpha=True)
        assert st == 1  # degenerate: handled by the fallback solver
        assert abs(d0 - d1) <= 1e-9 * d0, (seed, d0, d1)
        assert abs(alpha.sum() - 1) < 1e-9 and alpha.min() >= -1e-12


def test_affine_metrics_with_dependent_vertices():
    """Duplicate contigs among the neighbours leave the affine hull unchanged; the QP form (hull_distance.py:38-61) is then
    singular and the restatement falls back to a rank-revealing Gram-Schmidt -- same value as the SVD form
    (hull_distance.py:64-87, restated in oracle.affine_hull_distance) and as the de-duplicated vertex set."""
    for seed in range(30):
        rng = np.random.default_rng(seed)
        m = int(rng.integers(2, 12))
        V = rng.dirichlet(np.full(137, 8.0), size=m)
        x = rng.dirichlet(np.full(137, 8.0))
        V2 = np.vstack([V, V[rng.integers(0, m, 3)]])[rng.permutation(m + 3)]
        d0 = oracle.affine_hull_distance_qp(x, V)
        d1 = oracle.affine_hull_distance_qp(x, V2)
        d2 = oracle.affine_hull_distance(x, V2)
        assert abs(d1 - d0) <= 1e-9 * d0 and abs(d2 - d0) <= 1e-9 * d0, (seed, d0, d1, d2)
