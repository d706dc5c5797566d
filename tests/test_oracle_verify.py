"""
CPU tests of the position-parallel oracle (oracle/verify.c) that the at-size GPU parity tests rely on: it must agree with
the sequential oracle (oracle/fit_cluster_ref.c, algorithm.py:12-76) wherever both can run, and it must notice a wrong
label.  Also the batched hull-distance entry point against the one-pair call.
"""
import numpy as np
import pytest

import oracle
from chbin_b200 import synth


@pytest.mark.parametrize("n,C,S,n_seed,k,conc,metric", [
    (1500, 8, 1, 30, 5, 60.0, "convex"),      # overlapping bins: labels keep changing for several iterations
    (900, 4, 3, 3, 7, 500.0, "convex"),       # bins smaller than k at the start
    (800, 5, 10, 20, 10, 300.0, "affine-qp"),
])
def test_verify_iteration_agrees_with_the_sequential_oracle(n, C, S, n_seed, k, conc, metric):
    X, bins, _ = synth.make_contig_features(n, C, S, n_seed, seed=7, concentration=conc)
    X[100:110] = X[100]  # duplicates: ties by index
    perms = oracle.draw_permutations(bins, 4, seed=0)
    prev = bins.copy()
    for it in range(4):
        cur = oracle.fit_cluster(X, C, prev, None, k, 1, metric=metric, perms=perms[it:it + 1].copy(), threads=2)
        res = oracle.verify_iteration(X, C, prev, cur, perms[it], k, metric=metric, threads=4, return_distances=True)
        assert res["mismatches"] == 0 and np.array_equal(res["labels"], cur[perms[it]])
        assert res["distances"].shape == (perms.shape[1], C)
        assert np.all(res["best"] <= res["second"])
        # a wrong label at one position is found at that position (and possibly at later ones that see it)
        bad = cur.copy()
        p = perms.shape[1] // 3
        j = perms[it][p]
        bad[j] = (bad[j] + 1) % C
        res2 = oracle.verify_iteration(X, C, prev, bad, perms[it], k, metric=metric, threads=4)
        assert res2["mismatches"] >= 1 and res2["labels"][p] != bad[j]
        # ... and a sample of positions checks exactly those positions
        pos = np.array([0, p, perms.shape[1] - 1], dtype=np.int64)
        res3 = oracle.verify_iteration(X, C, prev, bad, perms[it], k, positions=pos, metric=metric, threads=2)
        assert list(res3["positions"]) == list(pos) and res3["labels"][1] != bad[j]
        if np.array_equal(cur, prev):
            break
        prev = cur


def test_hull_distance_batch_matches_single_calls():
    X, bins, _ = synth.make_contig_features(600, 4, 2, 20, seed=3)
    rng = np.random.default_rng(1)
    P, k = 200, 6
    q = rng.integers(0, len(X), P)
    idx = rng.integers(0, len(X), (P, k))
    m = rng.integers(0, k + 1, P).astype(np.int32)
    for metric in ("convex", "affine-qp"):
        dist, status = oracle.hull_distance_batch(X, q, idx, m, metric=metric, threads=3)
        for p in range(P):
            if m[p] == 0:
                assert np.isinf(dist[p]) and status[p] == 3
                continue
            pts = X[idx[p, : m[p]]]
            ref = oracle.convex_hull_distance(X[q[p]], pts) if metric == "convex" else oracle.affine_hull_distance_qp(X[q[p]], pts)
            assert dist[p] == ref
