import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from chbin_b200 import capi, synth
for (n, S, conc) in [(3000, 1, 4000.0), (3000, 10, 300.0), (2000, 20, 60.0)]:
    X, bins, _ = synth.make_contig_features(n, 8, S, 10, seed=3, concentration=conc)
    pts = np.where(bins == -1)[0]
    D = oracle.create_in_mem_distance_matrix(X)[pts]
    for eng in (0, 1):
        ctx = capi.Context(0); ctx.set_features(X); ctx.set_labels(bins, 8); ctx.set_params(5, "convex")
        ctx.set_gram_engine(eng); ctx.build_distance_matrix(True)
        A, eps, nrm = ctx.get_candidate_rows(0, len(pts)); ctx.close()
        err = np.abs(A.astype(np.float64) - D * D)
        scale = nrm[pts].astype(np.float64)[:, None] + float(nrm.max())
        print(f"n={n} d={X.shape[1]} engine={eng}: eps_rel={eps:.3e} max err/scale={np.max(err/scale):.3e} ratio to bound={np.max(err/scale)/eps:.4f}  rms={np.sqrt(np.mean((err/scale)**2)):.2e}")
