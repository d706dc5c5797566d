"""
BASELINE.md section 3, baselines B1 and B3, measured in the BUILD CONTAINER (the only place /root/reference exists; the
GPU box cannot run them, so bench.py quotes the file this script writes):

  B1  reference-verbatim: /root/reference/ch_bin/core/clustering/algorithm.py:fit_cluster imported UNMODIFIED
      (oracle/ref_shim.py), single thread like the reference, on the first STEPS permuted queries of iteration 1 of a
      BASELINE workload.  quadprog / cvxopt are not installed in this image: the solver behind solve_qp.py:51 is the
      restated Goldfarb-Idnani solver (oracle/gi_qp.c) -- "reference flow + restated solver".
  B3  scipy cdist: create_in_mem_distance_matrix (distance_matrix.py:33-44), single thread, full n (n^2 * 8 B must fit).

usage: python oracle/baseline_b1.py [workload=20k] [steps=2000]   ->  profiles/r2_baseline_b1_b3_container.json
"""
import itertools
import json
import os
import platform
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from chbin_b200 import synth
from oracle import ref_shim


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "20k"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    X, bins, _, cfg = synth.make_config(wl, seed=0)
    n, d = X.shape
    C, k = cfg["C"], cfg["k"]
    ref = ref_shim.load()
    out = {"workload": wl, "n": n, "d": d, "C": C, "k": k, "threads": 1, "host": platform.processor() or platform.machine(),
           "cpu_count": os.cpu_count(), "real_solvers": dict(ref_shim.REAL_SOLVERS),
           "cvxopt": "real" if ref_shim.REAL_SOLVERS["cvxopt"] else "unavailable in image",
           "quadprog": "real" if ref_shim.REAL_SOLVERS["quadprog"] else "unavailable in image (restated GI solver, oracle/gi_qp.c)"}
    # B3
    t0 = time.perf_counter()
    D = ref.distance_matrix.create_in_mem_distance_matrix(X)
    dt = time.perf_counter() - t0
    out["b3_cdist"] = {"seconds": dt, "pairs": n * n, "pairs_per_s": n * n / dt, "matrix_bytes": int(D.nbytes)}
    print("B3", out["b3_cdist"], flush=True)
    # B1: the loop of algorithm.py:46 runs over tqdm(sample_perm): hand it a tqdm that stops after `steps` samples
    orig = ref.algorithm.tqdm
    ref.algorithm.tqdm = lambda it, **kw: itertools.islice(it, steps)
    import logging

    logging.getLogger("ch_bin").setLevel(logging.WARNING)
    try:
        # warm the numba JIT of positive_def.py outside the timed region
        np.random.seed(0)
        ref.algorithm.tqdm = lambda it, **kw: itertools.islice(it, 3)
        ref.algorithm.fit_cluster(samples=X, num_clusters=C, initial_bins=bins, distance_matrix=D, num_neighbors=k,
                                  max_iterations=1, metric="convex", qp_solver="quadprog")
        ref.algorithm.tqdm = lambda it, **kw: itertools.islice(it, steps)
        np.random.seed(0)
        t0 = time.perf_counter()
        ref.algorithm.fit_cluster(samples=X, num_clusters=C, initial_bins=bins, distance_matrix=D, num_neighbors=k,
                                  max_iterations=1, metric="convex", qp_solver="quadprog")
        dt = time.perf_counter() - t0
    finally:
        ref.algorithm.tqdm = orig
    out["b1_reference_verbatim"] = {"steps": steps, "qps": steps * C, "seconds": dt, "qps_per_s_per_core": steps * C / dt,
                                    "what": "algorithm.fit_cluster unmodified, first %d permuted queries of iteration 1, "
                                            "matrix in RAM, single thread" % steps}
    print("B1", out["b1_reference_verbatim"], flush=True)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_baseline_b1_b3_container.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
