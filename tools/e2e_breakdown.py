import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chbin_b200 import capi, synth
X, bins, truth, cfg = synth.make_config("20k", seed=0)
pts = np.where(bins == -1)[0]
def T(): return time.perf_counter()
ctx = capi.Context(0)
for rep in range(4):
    np.random.seed(0)
    t1=T()
    ctx.reset_timers()
    ctx.set_features(X); t2=T()
    ctx.set_labels(bins, cfg["C"]); ctx.set_params(5,"convex"); t3=T()
    ctx.build_distance_matrix(True); t4=T()
    its=[]
    for it in range(10):
        ta=T()
        perm = np.random.permutation(pts).astype(np.int64)
        tb=T()
        _, nch = ctx.fit_iteration(perm, want_labels=False)
        its.append((1e3*(tb-ta), 1e3*(T()-tb)))
        if nch == 0: break
    t5=T(); lab = ctx.get_labels(); t6=T(); tm = ctx.timers()
    print(f"features {1e3*(t2-t1):.2f} labels {1e3*(t3-t2):.2f} build {1e3*(t4-t3):.2f} fit {1e3*(t5-t4):.2f} {[(round(a,2),round(b,2)) for a,b in its]} get {1e3*(t6-t5):.2f} total {1e3*(t6-t1):.2f} | kernels {tm['ms_distance']+tm['ms_knn']+tm['ms_qp']+tm['ms_commit']:.2f} rounds {tm['rounds']}")
ctx.close()
