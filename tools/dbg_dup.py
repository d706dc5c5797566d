import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, chbin_b200
from chbin_b200 import synth, capi
X, bins, _ = synth.make_contig_features(1200, 5, 1, 20, seed=17, concentration=300.0)
X[300:340] = X[300]; X[700:712] = X[700]
perms = oracle.draw_permutations(bins, 10, seed=0)
D = oracle.create_in_mem_distance_matrix(X)
pts = np.where(bins == -1)[0]
ctx = capi.Context(0); ctx.set_features(X); ctx.set_params(5, "convex"); ctx.set_distance_mode(2); ctx.set_labels(bins, 5); ctx.build_distance_matrix(True)
cur = bins.copy()
for it in range(3):
    lab, nch = ctx.fit_iteration(perms[it])
    o = oracle.fit_cluster(X, 5, cur, None, 5, 1, perms=perms[it:it+1], threads=4)
    idx, cnt, dist = ctx.get_pair_cache(0, len(pts))
    # verify the cache against the oracle for the FINAL state of this iteration (as each query saw it)
    pos = np.full(len(X), -1); pos[perms[it]] = np.arange(len(perms[it]))
    bad = 0
    for u, j in enumerate(pts):
        eff = np.where(pos < pos[j], lab, cur); eff[j] = -1
        for c in range(5):
            refset = np.sort(oracle.find_nearest_from_cluster(c, eff, D[j], 5))
            got = np.sort(idx[u, c, :cnt[u, c]])
            if not np.array_equal(refset, got):
                bad += 1
                if bad <= 5:
                    print("  it", it, "query", j, "bin", c, "ref", refset, D[j][refset], "got", got, D[j][got], "dist", dist[u, c])
    print(f"iter {it}: changed={nch} label diffs {np.sum(lab != o)} bad sets {bad}")
    cur = o
ctx.close()
