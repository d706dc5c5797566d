import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chbin_b200 import capi, synth
X, bins, truth, cfg = synth.make_config("20k", seed=0)
pts = np.where(bins == -1)[0]
def T(): return time.perf_counter()
for rep in range(3):
    np.random.seed(0)
    t0=T(); ctx = capi.Context(0); t1=T()
    ctx.set_features(X); t2=T()
    ctx.set_labels(bins, cfg["C"]); ctx.set_params(5,"convex"); t3=T()
    ctx.build_distance_matrix(True); t4=T()
    for it in range(10):
        perm = np.random.permutation(pts).astype(np.int64)
        _, nch = ctx.fit_iteration(perm, want_labels=False)
        if nch == 0: break
    t5=T(); lab = ctx.get_labels(); t6=T(); tm = ctx.timers(); ctx.close(); t7=T()
    print(f"create {1e3*(t1-t0):.1f} features {1e3*(t2-t1):.1f} labels {1e3*(t3-t2):.1f} build {1e3*(t4-t3):.1f} fit {1e3*(t5-t4):.1f} get {1e3*(t6-t5):.1f} close {1e3*(t7-t6):.1f} total {1e3*(t7-t0):.1f} | kernels {tm['ms_distance']+tm['ms_knn']+tm['ms_qp']+tm['ms_commit']:.1f}")
