"""
TEST INFRASTRUCTURE: a CPU stand-in for chbin_b200.GpuEngine with the same round interface, built on the oracle, so that
the host-side logic (speculate/repair driver, query sharding, the per-round label exchange) is testable without a GPU and
with world_size > 1 over gloo.  It restates what csrc/api.cu + knn.cu + qp*.cu compute per round.
"""
import numpy as np
import torch

import oracle

UNOWNED = -(2**31)


class OracleEngine:
    def __init__(self, X, bins, C, k, u0=0, u1=None, window=0):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.n = len(X)
        self.C, self.k = int(C), int(k)
        self.old = np.asarray(bins, dtype=np.int64).copy()
        self.tent_pt = self.old.copy()
        self.qpoint = np.where(self.old == -1)[0]
        self.U = len(self.qpoint)
        self.u0, self.u1 = u0, (self.U if u1 is None else u1)
        self.slot = np.full(self.n, -1, dtype=np.int64)
        self.slot[self.qpoint] = np.arange(self.U)
        self._window = window
        self.D = oracle.create_in_mem_distance_matrix(self.X)
        self.qps = 0

    def window(self):
        return self._window or self.U

    def iteration_begin(self, perm):
        self.perm = np.asarray(perm, dtype=np.int64)
        self.pos = np.full(self.n, -1, dtype=np.int64)
        self.pos[self.perm] = np.arange(len(self.perm))
        self.tent_pt = self.old.copy()

    def _assign(self, p):
        j = self.perm[p]
        eff = np.where(self.pos < p, self.tent_pt, self.old)
        eff[j] = -1
        best, bc = np.inf, self.old[j]
        for c in range(self.C):
            idx = oracle.find_nearest_from_cluster(c, eff, self.D[j], self.k)
            if len(idx) == 0:
                continue
            d = oracle.convex_hull_distance(self.X[j], self.X[idx])
            self.qps += 1
            if best > d:
                best, bc = d, c
        return bc

    def round_run(self, lo, hi):
        tent = torch.full((hi - lo,), UNOWNED, dtype=torch.int32)
        for p in range(lo, hi):
            s = self.slot[self.perm[p]]
            if self.u0 <= s < self.u1:
                tent[p - lo] = int(self._assign(p))
        return tent

    def round_commit(self, lo, hi, tent):
        t = tent.numpy().astype(np.int64)
        assert np.all(t != UNOWNED), "a position was left un-owned after the exchange"
        pts = self.perm[lo:hi]
        changed = np.where(self.tent_pt[pts] != t)[0]
        self.tent_pt[pts] = t
        return -1 if len(changed) == 0 else int(lo + changed[0])

    def iteration_end(self):
        n_changed = int(np.sum(self.old != self.tent_pt))
        self.old = self.tent_pt.copy()
        return n_changed

    def get_labels(self):
        return self.old.copy()
