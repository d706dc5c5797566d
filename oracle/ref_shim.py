"""
oracle/ref_shim.py -- TEST INFRASTRUCTURE (build container only).

Imports the reference's OWN clustering modules, unmodified, from /root/reference:
    ch_bin/core/clustering/algorithm.py       (fit_cluster)
    ch_bin/core/clustering/distance_matrix.py (cdist matrix, find_nearest_from_cluster)
    ch_bin/core/clustering/hull_distance.py   (convex_hull_distance, affine variants)
    ch_bin/core/clustering/solve_qp.py        (solver dispatch, quadprog->cvxopt fallback)
    ch_bin/core/clustering/positive_def.py    (numba nearest-PD)
The two third-party solver packages they import (`quadprog==0.1.8`, `cvxopt==1.2.6`,
/root/reference/requirements.txt:7-8) are NOT installed in this image and cannot be (no network), so
stand-in modules are injected into sys.modules first:
    quadprog.solve_qp   -> oracle.quadprog_solve_qp   (oracle/gi_qp.c, Goldfarb-Idnani restatement)
    cvxopt.solvers.qp   -> oracle.simplex_qp          (oracle/minnorm.c) -- only shape the reference uses
If the real packages ARE importable they are used instead and `REAL_SOLVERS` says so.

This module cannot travel to the GPU box (/root/reference does not exist there); it is used here to
validate the C restatement and by tests/golden/make_golden.py to generate the committed fixtures.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("CHB_REFERENCE_ROOT", "/root/reference")
REAL_SOLVERS = {"quadprog": False, "cvxopt": False}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ch_bin", "core", "clustering"))


def _install_quadprog():
    try:
        import quadprog  # noqa: F401

        REAL_SOLVERS["quadprog"] = True
        return
    except ImportError:
        pass
    import oracle

    mod = types.ModuleType("quadprog")
    mod.solve_qp = oracle.quadprog_solve_qp
    mod.__doc__ = "stand-in: oracle/gi_qp.c"
    sys.modules["quadprog"] = mod


def _install_cvxopt():
    try:
        import cvxopt  # noqa: F401

        REAL_SOLVERS["cvxopt"] = True
        return
    except ImportError:
        pass
    import oracle

    mod = types.ModuleType("cvxopt")
    solvers = types.ModuleType("cvxopt.solvers")
    solvers.options = {}

    def matrix(a):
        return np.array(a, dtype=np.float64)

    def qp(P, q, G=None, h=None, A=None, b=None):
        P = np.asarray(P, dtype=np.float64)
        q = np.asarray(q, dtype=np.float64).reshape(-1)
        m = P.shape[0]
        # the only shapes the reference ever passes (hull_distance.py:19-23, 50-54)
        if A is None or np.asarray(A).shape != (1, m) or not np.all(np.asarray(A) == 1.0) or float(np.asarray(b).reshape(-1)[0]) != 1.0:
            raise NotImplementedError("cvxopt stand-in: only 1'a = 1 supported")
        G = np.asarray(G)
        if G.shape == (m, m) and np.array_equal(G, -np.eye(m)) and not np.any(np.asarray(h)):
            x = oracle.simplex_qp(P, q)
            return {"status": "optimal", "x": x.reshape(m, 1)}
        if G.shape[0] == 0:
            kkt = np.block([[P, np.ones((m, 1))], [np.ones((1, m)), np.zeros((1, 1))]])
            sol = np.linalg.lstsq(kkt, np.concatenate([-q, [1.0]]), rcond=None)[0]
            return {"status": "optimal", "x": sol[:m].reshape(m, 1)}
        raise NotImplementedError("cvxopt stand-in: unsupported inequality block")

    mod.matrix = matrix
    solvers.qp = qp
    mod.solvers = solvers
    sys.modules["cvxopt"] = mod
    sys.modules["cvxopt.solvers"] = solvers


_loaded = {}


def load():
    """Returns a namespace with the reference's modules: .algorithm .distance_matrix .hull_distance .solve_qp"""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_quadprog()
    _install_cvxopt()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("distance_matrix", "positive_def", "solve_qp", "hull_distance", "algorithm"):
        _loaded[name] = importlib.import_module(f"ch_bin.core.clustering.{name}")
    return types.SimpleNamespace(**_loaded)


def fit_cluster_reference(samples, num_clusters, initial_bins, num_neighbors, max_iterations, metric="convex",
                          qp_solver="quadprog", seed=0, distance_matrix=None):
    """Runs the reference's fit_cluster exactly as ch_bin/cli/clustering.py:56-76 does, with np.random.seed(seed)
    (ch_bin/ch_bin.py:22) immediately before."""
    ref = load()
    import logging

    logging.getLogger("ch_bin").setLevel(logging.WARNING)
    if distance_matrix is None:
        distance_matrix = ref.distance_matrix.create_in_mem_distance_matrix(samples)
    if seed is not None:
        np.random.seed(seed)
    # silence tqdm
    import tqdm as _tqdm

    orig = ref.algorithm.tqdm
    ref.algorithm.tqdm = lambda it, **kw: it
    try:
        return ref.algorithm.fit_cluster(
            samples=samples, num_clusters=num_clusters, initial_bins=initial_bins, distance_matrix=distance_matrix,
            num_neighbors=num_neighbors, max_iterations=max_iterations, metric=metric, qp_solver=qp_solver,
        )
    finally:
        ref.algorithm.tqdm = orig
        del _tqdm


def load_cli_clustering():
    """The reference's own ch_bin.cli.clustering module (perform_clustering / run_perform_clustering), imported with a
    stand-in for Biopython's SeqIO: step 05 (dump_bins, FASTA output -- outside the hot path) becomes a no-op, steps
    01-04 (cli/clustering.py:47-92) execute verbatim."""
    load()
    if "ch_bin.cli.clustering" in sys.modules:
        return sys.modules["ch_bin.cli.clustering"]
    if "Bio" not in sys.modules:
        try:
            import Bio  # noqa: F401
        except ImportError:
            bio = types.ModuleType("Bio")
            seqio = types.ModuleType("Bio.SeqIO")
            seqio.parse = lambda *a, **k: iter(())
            bio.SeqIO = seqio
            sys.modules["Bio"] = bio
            sys.modules["Bio.SeqIO"] = seqio
    mod = importlib.import_module("ch_bin.cli.clustering")
    mod.dump_bins = lambda df_bins, contig_fasta, operating_dir: None
    return mod
