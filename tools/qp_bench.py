#!/usr/bin/env python
"""BASELINE config #5, QP kernel isolated: fixed precomputed neighbour indices (chb_knn_per_bin once), then
chb_hull_distance_batch timed by the library's own CUDA-event timers, per AlgoNumNeighbors.  Reports QP/s, FP64 TFLOP/s by
the algorithmic count F(k, d) of SURVEY 8(d) and the per-pair status histogram (0 ok, 1 degenerate, 2 iteration cap).
usage: python tools/qp_bench.py [k ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chbin_b200 import capi, synth  # noqa: E402


def flops(k, d):
    return k * (k + 1) * d + 2 * k * d + 2 * k * d + 3 * d + 2 * k ** 3


def main():
    ks = [int(a) for a in sys.argv[1:]] or [3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32]
    n, C, nq = 50_000, 25, 40_000
    X, bins, truth = synth.make_contig_features(n, C, 1, 50, seed=0)
    queries = np.where(bins == -1)[0][:nq].astype(np.int64)
    with capi.Context(0) as ctx:
        ctx.set_features(X)
        ctx.set_labels(bins, C)
        for k in ks:
            ctx.set_params(k, "convex")
            ctx.set_distance_mode(0)
            ctx.build_distance_matrix(False)
            idx, m = ctx.knn_per_bin(truth, queries)
            for rep in range(2):
                ctx.reset_timers()
                dist, status = ctx.hull_distance_batch(queries, idx, m)
                t = ctx.timers()
            nqp = dist.size
            ms = t["ms_qp"]
            hist = np.bincount(status.ravel(), minlength=4)
            print("k=%2d: %d QPs in %.2f ms = %.1f M QP/s, %.2f TFLOP/s FP64 (algorithmic), status ok/degenerate/cap/empty = %s"
                  % (k, nqp, ms, nqp / ms / 1e3, nqp * flops(k, X.shape[1]) / ms / 1e9, hist.tolist()), flush=True)


if __name__ == "__main__":
    main()
