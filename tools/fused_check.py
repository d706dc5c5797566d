import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, chbin_b200
from chbin_b200 import synth
n, C, S, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
X, bins, truth = synth.make_contig_features(n, C, S, 20, seed=5, concentration=2000.0)
perms = oracle.draw_permutations(bins, 10, seed=0)
np.random.seed(0)
got, info = chbin_b200.fit_cluster(X, C, bins, None, k, 10, distance_mode=2, return_info=True)
print("gpu done", info["iterations"], info["timers"]["ms_knn"], info["timers"]["ms_qp"])
if n <= 20000:
    ref = oracle.fit_cluster(X, C, bins, None, k, 10, perms=perms, threads=16)
    print("equal:", np.array_equal(got, ref))
