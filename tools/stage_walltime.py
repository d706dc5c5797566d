"""Wall-clock time of every C-ABI call of one clustering stage (synchronised after each call): where a stage's time goes
on the host side.  usage: python tools/stage_walltime.py [workload] [n]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from chbin_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "20k"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
X, bins, truth, cfg = synth.make_config(wl, seed=0, n=n)
pts = np.where(bins == -1)[0]
np.random.seed(0)
perms = np.stack([np.random.permutation(pts) for _ in range(3)]).astype(np.int64)
U = len(pts)
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
tent = torch.empty(U, dtype=torch.int32, device=dev)
def T(label, fn):
    t0 = time.perf_counter(); r = fn(); ctx.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"   {label:<22s} {dt:9.3f} ms")
    return r
for rep in range(3):
    print(f"rep {rep}")
    T("set_features", lambda: ctx.set_features(X))
    T("set_labels", lambda: ctx.set_labels(bins, cfg["C"]))
    ctx.set_params(cfg["k"], "convex"); ctx.set_distance_mode(2)
    T("build_distance_matrix", lambda: ctx.build_distance_matrix(True))
    for it in range(3):
        T(f"it{it+1} iteration_begin", lambda: ctx.iteration_begin(perms[it]))
        lo = 0
        while lo < U:
            T(f"it{it+1} round_run", lambda: ctx.round_run(lo, U, tent.data_ptr()))
            first = T(f"it{it+1} round_commit", lambda: ctx.round_commit(lo, U, tent.data_ptr()))
            lo = U if first < 0 else first + 1
        nch = T(f"it{it+1} iteration_end", lambda: ctx.iteration_end())
        if nch == 0: break
    lab = T("get_labels", lambda: ctx.get_labels())
    print("   acc", float(np.mean(lab == truth)), ctx.timers())
ctx.close()
