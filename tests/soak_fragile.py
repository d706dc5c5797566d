#!/usr/bin/env python
"""CPU-side reading of a mismatching case of tests/soak_parity.py: lists the positions whose decision the ORACLE itself takes
on rounding noise -- the two smallest hull distances both ~1e-16 (a query with an exact duplicate in two bins; the library
returns 0.0 for both and keeps the lower bin, tests/test_gpu_parity.py::test_query_duplicated_in_two_bins_keeps_the_lower_bin).
A soak mismatch count equal to the number of such queries is the documented benign tie.  usage: python tests/soak_fragile.py SEED..."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle
from soak_parity import case
def fragile(seed):
    cfg, X, bins = case(seed)
    C,k,iters,metric = cfg["C"],cfg["k"],cfg["iters"],cfg["metric"]
    perms = oracle.draw_permutations(bins, iters, seed=0)
    prev = bins; out=[]
    for it in range(iters):
        lab = oracle.fit_cluster(X, C, prev, None, k, 1, metric=metric, perms=perms[it:it+1], threads=8)
        res = oracle.verify_iteration(X, C, prev, lab, perms[it], k, metric=metric, threads=8, return_distances=True)
        assert res["mismatches"]==0
        b, s = res["best"], res["second"]
        fr = np.where((s - b <= 1e-9*np.maximum(b,1e-300)) | (s < 1e-10))[0]
        for p in fr:
            out.append((it+1, int(p), int(perms[it][p]), float(b[p]), float(s[p])))
        if np.array_equal(lab, prev): break
        prev = lab
    return cfg, out
for seed in [int(a) for a in sys.argv[1:]]:
    cfg, out = fragile(seed)
    print(seed, cfg['k'], cfg['metric'], cfg['ndup'], 'fragile positions:', len(out), out[:6])
