"""GPU: the CUDA path (through the C-ABI) against the committed golden vectors of the reference's own modules."""
import os

import numpy as np
import pytest

import chbin_b200
from chbin_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_golden.npz"))


@pytest.mark.parametrize("name", ["easy", "hard", "k10", "smallbins", "fg"])
def test_fit_cluster_matches_reference_run(G, name):
    pre = "fg" if name == "fg" else f"fit_{name}"
    X, bins, ref = G[f"{pre}_X"], G[f"{pre}_bins"], G[f"{pre}_labels"]
    C, k, iters = (int(v) for v in G[f"{pre}_params"])
    np.random.seed(0)  # ch_bin/ch_bin.py:22
    got = chbin_b200.fit_cluster(np.asfortranarray(X), C, bins, None, k, iters)
    assert np.array_equal(got, ref), f"{np.sum(got != ref)} labels differ from the reference's fit_cluster"


def test_fit_cluster_affine_matches_reference_run(G):
    """metric="affine" (hull_distance.py:69-87) against the verbatim reference loop."""
    X, bins, ref = G["fit_affine_X"], G["fit_affine_bins"], G["fit_affine_labels"]
    C, k, iters = (int(v) for v in G["fit_affine_params"])
    for metric in ("affine", "affine-qp"):
        np.random.seed(0)
        got = chbin_b200.fit_cluster(X, C, bins, None, k, iters, metric=metric)
        assert np.array_equal(got, ref), f"{metric}: {np.sum(got != ref)} labels differ from the reference's fit_cluster"


@pytest.mark.parametrize("dist_mode", [0, 1])
def test_knn_sets_match_reference(G, dist_mode):
    X, labels, queries, k = G["knn_X"], G["knn_labels"], G["knn_queries"], int(G["knn_k"])
    ctx = capi.Context(0)
    ctx.set_features(X)
    ctx.set_labels(np.full(len(X), -1), 4)
    ctx.set_params(k, "convex")
    ctx.set_distance_mode(dist_mode)
    ctx.build_distance_matrix(True)
    idx, m = ctx.knn_per_bin(labels, queries)
    rows = ctx.get_distance_rows(0, len(X))
    ctx.close()
    assert np.array_equal(rows, G["knn_D"]), "distance rows are not bit-identical to scipy cdist"
    it = iter(G["knn_sets"])
    for qi in range(len(queries)):
        for c in range(4):
            ref = next(it)
            ref = ref[ref >= 0]
            assert np.array_equal(np.sort(idx[qi, c, : m[qi, c]]), ref)


def test_hull_distances_match_reference(G):
    X = G["hull_X"]
    hq, hidx, hm = G["hull_q"], G["hull_idx"], G["hull_m"]
    kmax = hidx.shape[1]
    for metric, key, tol in (("convex", "hull_dist", 1e-6), ("affine-qp", "hull_affine_qp", 1e-6), ("affine", "hull_affine", 1e-6)):
        ctx = capi.Context(0)
        ctx.set_features(X)
        ctx.set_labels(np.full(len(X), -1), 1)
        ctx.set_params(kmax, metric)
        ctx.build_distance_matrix(False)
        dist, status = ctx.hull_distance_batch(hq, hidx.reshape(len(hq), 1, kmax), hm.reshape(-1, 1).astype(np.int32))
        ctx.close()
        ref = G[key]
        for i in range(len(hq)):
            if not np.isfinite(ref[i]):
                continue
            assert abs(dist[i, 0] - ref[i]) <= tol * ref[i] + 1e-12 * np.linalg.norm(X[hq[i]]), (metric, i, hm[i])
    # k <= 5 kernel on the subset with m <= 5
    sel = np.where(hm <= 5)[0]
    ctx = capi.Context(0)
    ctx.set_features(X)
    ctx.set_labels(np.full(len(X), -1), 1)
    ctx.set_params(5, "convex")
    ctx.build_distance_matrix(False)
    dist, status = ctx.hull_distance_batch(hq[sel], hidx[sel, :5].reshape(len(sel), 1, 5), hm[sel].reshape(-1, 1).astype(np.int32))
    ctx.close()
    ref = G["hull_dist"][sel]
    assert np.all(np.abs(dist[:, 0] - ref) <= 1e-6 * ref + 1e-12)
