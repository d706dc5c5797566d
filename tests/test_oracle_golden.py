"""
CPU: the oracle (oracle/*.c) against the committed golden vectors produced by the reference's own modules
(tests/golden/make_golden.py), and against scipy/numpy directly -- scipy.cdist and numpy.argpartition ARE the
reference's implementation of the distance / kNN stage (distance_matrix.py:26,41,57-62).
"""
import os

import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_golden.npz"))


def test_golden_records_which_solvers_made_it(G, record_property):
    """quadprog / cvxopt are not in this image: the QP goldens come from the reference's own modules running over the
    stand-in solvers of oracle/ref_shim.py (scipy SLSQP / trust-constr, same problem, same tolerance class).  The file
    says which, so a regenerated file made WITH the real packages is recognisable."""
    flags = G["real_solvers"]
    assert flags.shape == (2,) and set(flags.tolist()) <= {0, 1}
    record_property("golden_quadprog_real", int(flags[0]))
    record_property("golden_cvxopt_real", int(flags[1]))


def test_cdist_recipe_bit_exact_vs_scipy():
    from scipy.spatial.distance import cdist

    rng = np.random.default_rng(0)
    for n, d in [(50, 137), (200, 146), (33, 5), (64, 156)]:
        X = rng.dirichlet(np.ones(d), size=n)
        D = cdist(X, X, metric="euclidean")
        Do = oracle.create_in_mem_distance_matrix(X)
        assert np.array_equal(D, Do)
        assert np.array_equal(Do, Do.T) and np.all(np.diag(Do) == 0.0)
        rows = rng.choice(n, 7, replace=False)
        assert np.array_equal(oracle.distance_rows(X, rows), D[rows])


def test_cdist_golden(G):
    assert np.array_equal(oracle.create_in_mem_distance_matrix(G["knn_X"]), G["knn_D"])


def test_find_nearest_golden(G):
    X, D, labels, queries, k = G["knn_X"], G["knn_D"], G["knn_labels"], G["knn_queries"], int(G["knn_k"])
    it = iter(G["knn_sets"])
    for q in queries:
        lab = labels.copy()
        lab[q] = -1
        for c in range(4):
            ref = next(it)
            ref = ref[ref >= 0]
            got = oracle.find_nearest_from_cluster(c, lab, D[q], k)
            assert np.array_equal(np.sort(got), ref)


def test_find_nearest_vs_numpy_argpartition():
    # distance_matrix.py:57-62 restated inline with numpy itself
    rng = np.random.default_rng(3)
    n = 500
    row = rng.random(n)
    bins = rng.integers(-1, 6, size=n)
    for c in range(6):
        for m in (1, 4, 5, 17, 200):
            members = np.where(bins == c)[0]
            if len(members) <= m:
                ref = members
            else:
                ref = members[np.argpartition(row[members], kth=m - 1)[:m]]
            got = oracle.find_nearest_from_cluster(c, bins, row, m)
            assert np.array_equal(np.sort(got), np.sort(ref))
            if len(members) <= m:
                assert np.array_equal(got, members)  # ascending index order, as the reference returns it


def test_hull_distance_golden(G):
    X = G["hull_X"]
    for q, idx, m, dref, aref in zip(G["hull_q"], G["hull_idx"], G["hull_m"], G["hull_dist"], G["hull_affine_qp"]):
        pts = X[idx[:m]]
        d = oracle.convex_hull_distance(X[q], pts)
        assert abs(d - dref) <= 1e-9 * dref + 1e-13, (m, d, dref)  # C restatement skips nearest-PD: <= 1e-9 apart
        if m >= 2 and np.isfinite(aref):
            da = oracle.affine_hull_distance_qp(X[q], pts)
            if np.isfinite(da):
                assert abs(da - aref) <= 1e-7 * aref + 1e-12, (m, da, aref)


def test_affine_svd_form_golden(G):
    """hull_distance.py:69-87 (scipy.linalg.orth projection) against the reference's own outputs, and its agreement
    with the QP form on these affinely independent neighbour sets (the GPU path evaluates both metrics with one solve)."""
    X = G["hull_X"]
    for q, idx, m, sref, aref in zip(G["hull_q"], G["hull_idx"], G["hull_m"], G["hull_affine"], G["hull_affine_qp"]):
        if m < 2:
            continue
        pts = X[idx[:m]]
        ds = oracle.affine_hull_distance(X[q], pts)
        assert abs(ds - sref) <= 1e-12 * sref + 1e-15, (m, ds, sref)
        assert abs(ds - oracle.affine_hull_distance_qp(X[q], pts)) <= 1e-7 * sref + 1e-12
        assert abs(sref - aref) <= 1e-9 * sref
    # affinely dependent neighbours: orth drops the null direction, the distance is that of the reduced hull
    pts = X[[3, 9, 3, 20]]
    assert abs(oracle.affine_hull_distance(X[50], pts) - oracle.affine_hull_distance(X[50], X[[3, 9, 20]])) <= 1e-12


def test_fit_cluster_affine_golden(G):
    X, bins, ref = G["fit_affine_X"], G["fit_affine_bins"], G["fit_affine_labels"]
    C, k, iters = (int(v) for v in G["fit_affine_params"])
    perms = oracle.draw_permutations(bins, iters, seed=0)
    got = oracle.fit_cluster(X, C, bins, None, k, iters, metric="affine", perms=perms, threads=2)
    assert np.array_equal(got, ref), f"{np.sum(got != ref)} labels differ from the verbatim reference run (metric=affine)"


@pytest.mark.parametrize("name", ["easy", "hard", "k10", "smallbins"])
def test_fit_cluster_golden(G, name):
    X, bins, ref = G[f"fit_{name}_X"], G[f"fit_{name}_bins"], G[f"fit_{name}_labels"]
    C, k, iters = (int(v) for v in G[f"fit_{name}_params"])
    perms = oracle.draw_permutations(bins, iters, seed=0)
    got = oracle.fit_cluster(X, C, bins, None, k, iters, perms=perms, threads=2)
    assert np.array_equal(got, ref), f"{np.sum(got != ref)} labels differ from the verbatim reference run"
    D = oracle.create_in_mem_distance_matrix(X)
    got2 = oracle.fit_cluster(X, C, bins, D, k, iters, perms=perms, threads=1)
    assert np.array_equal(got2, ref)


def test_five_genomes_like_golden(G):
    X, bins, ref = G["fg_X"], G["fg_bins"], G["fg_labels"]
    C, k, iters = (int(v) for v in G["fg_params"])
    assert X.shape == (735, 137) and int(np.sum(bins == -1)) == 246
    assert list(np.bincount(bins[bins >= 0])) == [118, 112, 100, 86, 73]
    perms = oracle.draw_permutations(bins, iters, seed=0)
    got = oracle.fit_cluster(X, C, bins, None, k, iters, perms=perms)
    assert np.array_equal(got, ref)
    assert not np.any(got < 0)  # cli/clustering.py:79-80
