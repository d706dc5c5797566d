"""One clustering stage through the C-ABI (for ncu launch lists at large n). usage: python tools/one_stage.py [workload] [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chbin_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "20k"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
X, bins, truth, cfg = synth.make_config(wl, seed=0, n=n)
pts = np.where(bins == -1)[0]
np.random.seed(0)
perms = np.stack([np.random.permutation(pts) for _ in range(3)]).astype(np.int64)
ctx = capi.Context(0)
ctx.set_features(X); ctx.set_params(cfg["k"], "convex"); ctx.set_distance_mode(2)
ctx.set_labels(bins, cfg["C"]); ctx.build_distance_matrix(True)
labels, iters, conv, changed = ctx.fit(perms, 3)
print("iterations", iters, "acc", float(np.mean(labels == truth)), ctx.timers())
ctx.close()
