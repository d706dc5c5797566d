"""One clustering stage of a BASELINE workload through the C-ABI, for ncu captures (no timing claims)."""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import chbin_b200
from chbin_b200 import capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="20k")
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--stages", type=int, default=1)
ap.add_argument("--mode", type=int, default=2)
ap.add_argument("--window", type=int, default=0)
args = ap.parse_args()
X, bins, truth, cfg = synth.make_config(args.workload, seed=0, n=args.n)
pts = np.where(bins == -1)[0]
np.random.seed(0)
perms = np.stack([np.random.permutation(pts) for _ in range(10)]).astype(np.int64)
ctx = capi.Context(0)
ctx.set_features(X)
ctx.set_params(cfg["k"], "convex")
ctx.set_distance_mode(args.mode)
ctx.set_window(args.window)
for s in range(args.stages):
    ctx.reset_timers()
    ctx.set_labels(bins, cfg["C"])
    ctx.build_distance_matrix(True)
    labels, iters, conv, changed = ctx.fit(perms, 10)
    t = ctx.timers()
    print(f"stage {s}: iterations={iters} converged={conv} changed={list(changed)} acc={np.mean(labels == truth):.4f}")
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()})
ctx.close()
# steady-state probe: two more iterations after convergence (nothing changes): cost of a pure warm scan
ctx = capi.Context(0)
ctx.set_features(X); ctx.set_params(cfg["k"], "convex"); ctx.set_distance_mode(args.mode); ctx.set_labels(bins, cfg["C"]); ctx.build_distance_matrix(True)
for it in range(4):
    ctx.reset_timers()
    _, nch = ctx.fit_iteration(perms[it], want_labels=False)
    t = ctx.timers()
    print(f"iteration {it+1}: changed={nch} rounds={t['rounds']} knn_ms={t['ms_knn']:.3f} ({t['launches_knn']} launches) qp_ms={t['ms_qp']:.3f} qps_solved={t['qps_solved']}")
ctx.close()
