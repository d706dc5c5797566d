import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chbin_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "100k"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
X, bins, truth, cfg = synth.make_config(wl, seed=0, n=n)
pts = np.where(bins == -1)[0]
np.random.seed(0)
perms = np.stack([np.random.permutation(pts) for _ in range(4)]).astype(np.int64)
for rep in range(3):
    ctx = capi.Context(0)
    ctx.set_features(X); ctx.set_params(cfg["k"], "convex"); ctx.set_distance_mode(2); ctx.set_labels(bins, cfg["C"]); ctx.build_distance_matrix(True)
    for it in range(3):
        ctx.reset_timers()
        lab, nch = ctx.fit_iteration(perms[it])
        t = ctx.timers()
        print(f"rep {rep} it {it+1}: changed={nch} rounds={t['rounds']} acc={np.mean(lab == truth):.5f} knn_launches={t['launches_knn']} gram_ms={t['ms_gram']:.2f} tiles={t['gram_tiles']} qps={t['qps_solved']}", flush=True)
    ctx.close()
