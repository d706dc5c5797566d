"""Where the end-to-end time of chbin_b200.fit_cluster goes (host wall clock per phase, pinned input as in bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from chbin_b200 import capi, synth
X, bins, truth, cfg = synth.make_config("20k", seed=0)
Xp = torch.empty(X.shape, dtype=torch.float64, pin_memory=True); Xp.copy_(torch.from_numpy(X)); Xh = Xp.numpy()
pts = np.where(bins == -1)[0]
T = time.perf_counter
ctx = capi.Context(0)
for rep in range(5):
    np.random.seed(0)
    t0 = T(); ctx.reset_timers(); ctx.set_features(Xh, asynchronous=True); t1 = T()
    ctx.set_labels(bins, cfg["C"]); ctx.set_params(5, "convex"); t2 = T()
    perm = np.random.permutation(pts).astype(np.int64); t3 = T()
    ctx.build_distance_matrix(True); t4 = T()
    its = []
    for it in range(10):
        ta = T(); ctx.iteration_begin(perm); tb = T(); ctx.round_run(0, len(perm)); tc = T()
        nxt = np.random.permutation(pts).astype(np.int64); td = T()
        first = ctx.round_commit(0, len(perm)); te = T(); nch = ctx.iteration_end(); tf = T()
        its.append(tuple(round(1e3 * v, 3) for v in (tb - ta, tc - tb, td - tc, te - td, tf - te)))
        perm = nxt
        if nch == 0: break
    t5 = T(); lab = ctx.get_labels(); t6 = T()
    print(f"features(async) {1e3*(t1-t0):.3f} labels {1e3*(t2-t1):.3f} draw {1e3*(t3-t2):.3f} build(wait) {1e3*(t4-t3):.3f} "
          f"iters[begin,run,draw,commit,end] {its} labels_out {1e3*(t6-t5):.3f} total {1e3*(t6-t0):.3f}")
ctx.close()
