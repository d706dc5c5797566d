import sys, os, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import chbin_b200
from chbin_b200 import synth
X, bins, truth, cfg = synth.make_config("20k", seed=0)
Xp = torch.empty(X.shape, dtype=torch.float64, pin_memory=True); Xp.copy_(torch.from_numpy(X)); Xh = Xp.numpy()
for _ in range(3):
    np.random.seed(0); chbin_b200.fit_cluster(Xh, cfg["C"], bins, None, 5, 10)
ts=[]
for _ in range(20):
    np.random.seed(0); t0=time.perf_counter(); chbin_b200.fit_cluster(Xh, cfg["C"], bins, None, 5, 10, return_info=True); ts.append(time.perf_counter()-t0)
print("fit_cluster ms: min %.3f median %.3f max %.3f" % (1e3*min(ts), 1e3*np.median(ts), 1e3*max(ts)))
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    np.random.seed(0); chbin_b200.fit_cluster(Xh, cfg["C"], bins, None, 5, 10, return_info=True)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
