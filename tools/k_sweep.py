"""BASELINE config #5: AlgoNumNeighbors sweep at 200k contigs (one clustering stage per k, labels checked against the
generator's ground truth).  usage: python tools/k_sweep.py [k ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chbin_b200 import capi, synth
ks = [int(a) for a in sys.argv[1:]] or [3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32]
X, bins, truth, cfg = synth.make_config("200k", seed=0)
pts = np.where(bins == -1)[0]
np.random.seed(0)
perms = np.stack([np.random.permutation(pts) for _ in range(3)]).astype(np.int64)
ctx = capi.Context(0)
ctx.set_features(X)
for k in ks:
    ctx.set_params(k, "convex"); ctx.set_distance_mode(2)
    for rep in range(2):
        ctx.reset_timers(); t0 = time.perf_counter()
        ctx.set_labels(bins, cfg["C"]); ctx.build_distance_matrix(True)
        labels, iters, conv, changed = ctx.fit(perms, 3)
        dt = time.perf_counter() - t0
    t = ctx.timers()
    print("200k k=%d: stage %.1f ms, %d iterations, acc %.4f, QPs solved %d of %d, gram %.1f ms, selection %.1f ms, qp %.1f ms"
          % (k, dt * 1e3, iters, float(np.mean(labels == truth)), t["qps_solved"], t["qps_reference"], t["ms_gram"], t["ms_knn"], t["ms_qp"]), flush=True)
ctx.close()
