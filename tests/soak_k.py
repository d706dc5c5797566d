#!/usr/bin/env python
"""Parity soak at fixed neighbour counts (default 14, 15: the fused path with no spare list entries), with and without
duplicate contigs, compact and uncompacted work items.  usage: python tests/soak_k.py [k ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chbin_b200  # noqa: E402
import oracle  # noqa: E402
from chbin_b200 import synth  # noqa: E402


def case(k, seed):
    rng = np.random.default_rng(1000 * k + seed)
    n, C, S = int(rng.integers(800, 6000)), int(rng.integers(2, 10)), int(rng.choice([1, 2, 10]))
    n_seed, conc = int(rng.integers(6, 50)), float(rng.choice([60.0, 250.0, 1000.0, 4000.0]))
    X, bins, _ = synth.make_contig_features(n, C, S, n_seed, seed=seed, concentration=conc)
    ndup = int(rng.choice([0, 3, 40, 200]))
    if ndup:
        X[rng.integers(0, n, ndup)] = X[rng.integers(0, n, ndup)]
    return dict(n=n, C=C, S=S, k=k, n_seed=n_seed, conc=conc, ndup=ndup, metric="convex", mode=2, iters=3), X, bins


def main():
    ks = [int(a) for a in sys.argv[1:]] or [14, 15]
    bad = 0
    for k in ks:
        for seed in range(24):
            cfg, X, bins = case(k, seed)
            perms = oracle.draw_permutations(bins, 3, seed=0)
            ref = oracle.fit_cluster(X, cfg["C"], bins, None, k, 3, perms=perms, threads=8)
            np.random.seed(0)
            got, info = chbin_b200.fit_cluster(X, cfg["C"], bins, None, k, 3, return_info=True)
            nb = int((got != ref).sum())
            bad += nb > 0
            print("k", k, "seed", seed, cfg, "gram launches", info["timers"]["launches_gram"], "mismatches", nb, flush=True)
    print("soak_k: cases with mismatches:", bad)


if __name__ == "__main__":
    main()
