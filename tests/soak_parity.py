#!/usr/bin/env python
"""
Randomised parity soak (GPU box): random small workloads -- sizes, bin counts, coverage samples, seed counts, neighbour
counts 1..32, metrics, distance modes, injected duplicate contigs -- each compared label for label with the oracle's
sequential fit_cluster.  Not collected by pytest (a failing random case must be reduced first); the cases it found live
on as regular tests.  usage: python tests/soak_parity.py [first_seed] [count]

Known benign differences (3 in seeds 0..759: 198, 657, 701; reduce with tests/soak_debug.py): a query that has an exact
duplicate in TWO bins is at distance 0 from both hulls.  The library returns 0.0 for both and the strict comparison of
algorithm.py:57 keeps the first bin; the oracle's (and quadprog's) solvers return 0.0 for one and 1e-16-size rounding noise
for the other, so their choice is decided by that noise.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chbin_b200  # noqa: E402
import oracle  # noqa: E402
from chbin_b200 import synth  # noqa: E402


def case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(600, 4000))
    C = int(rng.integers(2, 13))
    S = int(rng.choice([1, 2, 5, 10, 20]))
    k = int(rng.choice([1, 2, 3, 4, 5, 5, 5, 6, 7, 8, 10, 10, 11, 12, 13, 14, 16, 20, 24, 32]))
    n_seed = int(rng.integers(max(2, k // 2), 60))
    conc = float(rng.choice([60.0, 250.0, 1000.0, 4000.0]))
    metric = str(rng.choice(["convex"] * 6 + ["affine-qp", "affine"]))
    mode = int(rng.choice([2, 2, 2, 1, 0]))
    iters = int(rng.integers(1, 5))
    X, bins, _ = synth.make_contig_features(n, C, S, n_seed, seed=seed, concentration=conc)
    ndup = int(rng.choice([0, 0, 3, 40]))
    if ndup:
        src = rng.integers(0, n, ndup)
        dst = rng.integers(0, n, ndup)
        X[dst] = X[src]
    return dict(n=n, C=C, S=S, k=k, n_seed=n_seed, conc=conc, metric=metric, mode=mode, iters=iters, ndup=ndup), X, bins


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    bad = 0
    t0 = time.time()
    for seed in range(first, first + count):
        cfg, X, bins = case(seed)
        perms = oracle.draw_permutations(bins, cfg["iters"], seed=0)
        try:
            ref = oracle.fit_cluster(X, cfg["C"], bins, None, cfg["k"], cfg["iters"], metric=cfg["metric"], perms=perms, threads=8)
            np.random.seed(0)
            got = chbin_b200.fit_cluster(X, cfg["C"], bins, None, cfg["k"], cfg["iters"], metric=cfg["metric"],
                                         distance_mode=cfg["mode"])
            nbad = int((got != ref).sum())
        except Exception as e:  # noqa: BLE001
            nbad = -1
            print("seed", seed, "EXCEPTION", repr(e)[:300])
        if nbad:
            bad += 1
        print("seed %d %s mismatches %d" % (seed, cfg, nbad), flush=True)
    print("soak: %d cases, %d with mismatches, %.0f s" % (count, bad, time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
