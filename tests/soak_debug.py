#!/usr/bin/env python
"""Reduces a failing case of tests/soak_parity.py: finds the first query (in assignment order) whose label differs from the
oracle's, rebuilds the label state at that step and prints, per bin, the oracle's neighbour set and hull distance (GI route
and min-norm route) next to the library's (chb_knn_per_bin + chb_hull_distance_batch).  usage: python tests/soak_debug.py SEED   |   soak_debug.py K SEED (a soak_k case)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import chbin_b200  # noqa: E402
import oracle  # noqa: E402
from chbin_b200 import capi  # noqa: E402
from soak_parity import case  # noqa: E402

if len(sys.argv) > 2:  # soak_k case: K SEED
    import soak_k

    cfg, X, bins = soak_k.case(int(sys.argv[1]), int(sys.argv[2]))
else:
    seed = int(sys.argv[1])
    cfg, X, bins = case(seed)
print(cfg)
C, k, iters, metric = cfg["C"], cfg["k"], cfg["iters"], cfg["metric"]
perms = oracle.draw_permutations(bins, iters, seed=0)
U = perms.shape[1]
first = None
for nit in range(1, iters + 1):
    ref = oracle.fit_cluster(X, C, bins, None, k, nit, metric=metric, perms=perms[:nit], threads=8)
    np.random.seed(0)
    got = chbin_b200.fit_cluster(X, C, bins, None, k, nit, metric=metric, distance_mode=cfg["mode"])
    bad = np.where(got != ref)[0]
    print("iterations", nit, "mismatches", bad.tolist(), "gpu", got[bad].tolist(), "oracle", ref[bad].tolist())
    if len(bad) and first is None:
        posn = {int(p): i for i, p in enumerate(perms[nit - 1])}
        q = min(bad.tolist(), key=lambda b: posn[b])
        first = (nit - 1, posn[q], q, int(got[q]), int(ref[q]))
        break
if first is None:
    sys.exit(0)
it, p, q, lg, lr = first
step = it * U + p
state = oracle.fit_cluster(X, C, bins, None, k, iters, metric=metric, perms=perms, max_steps=step, threads=8)
state[q] = -1
row = oracle.distance_rows(X, np.array([q]))[0]
np.set_printoptions(linewidth=220, precision=17)
with capi.Context(0) as ctx:
    ctx.set_features(X); ctx.set_labels(bins, C); ctx.set_params(k, metric); ctx.set_distance_mode(0); ctx.build_distance_matrix(False)
    gidx, gm = ctx.knn_per_bin(state, np.array([q], dtype=np.int64))
    gdist, gstat = ctx.hull_distance_batch(np.array([q], dtype=np.int64), gidx, gm)
    print("query", q, "iteration", it + 1, "position", p, "gpu label", lg, "oracle label", lr)
    for c in range(C):
        oi = np.sort(oracle.find_nearest_from_cluster(c, state, row, k))
        gi = np.sort(gidx[0, c, :gm[0, c]])
        V = X[oi]
        if metric == "convex":
            dg, alpha, st = oracle.convex_hull_distance(X[q], V, return_alpha=True)
            a2 = oracle.simplex_qp(2 * V @ V.T, -2 * V @ X[q])
            dm = float(np.linalg.norm(a2 @ V - X[q]))
        else:
            dg, st, dm = oracle.affine_hull_distance_qp(X[q], V), -1, oracle.affine_hull_distance(X[q], V)
        ndup = len(oi) - len(np.unique(V, axis=0))
        print("  bin %2d m=%2d sets equal %s dupes %d | oracle GI %.16g (status %d) other %.16g | gpu %.16g (status %d)"
              % (c, len(oi), np.array_equal(oi, gi), ndup, dg, st, dm, gdist[0, c], gstat[0, c]))
