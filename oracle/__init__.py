"""
oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of the CPU restatement of the reference's clustering hot path (oracle/*.c).  Function names
follow the reference (`/root/reference/ch_bin/core/clustering/*.py`) so that parity tests read like tests of
the reference.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; `ch-bin_b200/` never does.

Parity status (see DESIGN.md "Oracle"):
  * distance matrix / kNN: PINNED -- bit-for-bit against scipy.cdist and numpy.argpartition, which ARE the
    reference's implementation of that stage (distance_matrix.py:26,41,57-62).
  * QP: quadprog 0.1.8 / cvxopt 1.2.6 are absent from this image => "parity unpinned" against those two
    packages; pinned instead against the verbatim reference flow (oracle/ref_shim.py), KKT certificates,
    closed forms, and two independent solvers.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib: Optional[ctypes.CDLL] = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile oracle/*.c into oracle/liboracle.so (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("gi_qp.c", "minnorm.c", "fit_cluster_ref.c", "verify.c", "oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-s", "-C", _HERE, "-B"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.chb_oracle_gi_solve.restype = ctypes.c_int
        L.chb_oracle_gi_solve.argtypes = [
            ctypes.c_int, _f64p, _f64p, ctypes.c_int, _f64p, _f64p, ctypes.c_int,
            _f64p, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p, ctypes.c_void_p,
            ctypes.POINTER(ctypes.c_int), ctypes.c_void_p,
        ]
        L.chb_oracle_simplex_qp.restype = ctypes.c_int
        L.chb_oracle_simplex_qp.argtypes = [ctypes.c_int, _f64p, _f64p, _f64p]
        L.chb_oracle_cdist.restype = None
        L.chb_oracle_cdist.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int32, _f64p]
        L.chb_oracle_cdist_rows.restype = None
        L.chb_oracle_cdist_rows.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int32, _i64p, ctypes.c_int64, _f64p]
        L.chb_oracle_find_nearest.restype = ctypes.c_int32
        L.chb_oracle_find_nearest.argtypes = [ctypes.c_int32, _i64p, ctypes.c_int64, _f64p, ctypes.c_int32, _i64p]
        L.chb_oracle_convex_hull_distance.restype = ctypes.c_double
        L.chb_oracle_convex_hull_distance.argtypes = [
            _f64p, _f64p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)
        ]
        L.chb_oracle_affine_hull_distance_qp.restype = ctypes.c_double
        L.chb_oracle_affine_hull_distance_qp.argtypes = [
            _f64p, _f64p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)
        ]
        L.chb_oracle_fit_cluster.restype = ctypes.c_int
        L.chb_oracle_fit_cluster.argtypes = [
            _f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _i64p, ctypes.c_void_p, ctypes.c_int32,
            ctypes.c_int32, ctypes.c_int32, _i64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _i64p,
            ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _i64p, ctypes.POINTER(ctypes.c_int64),
        ]
        L.chb_oracle_verify_positions.restype = ctypes.c_int64
        L.chb_oracle_verify_positions.argtypes = [
            _f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _i64p, _i64p, _i64p, ctypes.c_int64, _i64p, ctypes.c_int64,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _i64p, _f64p, _f64p, ctypes.c_void_p,
        ]
        L.chb_oracle_hull_distance_batch.restype = None
        L.chb_oracle_hull_distance_batch.argtypes = [
            _f64p, ctypes.c_int32, _i64p, _i64p, _i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _f64p, _i32p,
        ]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------------------------
# quadprog.solve_qp stand-in (solve_qp.py:51)
# ----------------------------------------------------------------------------------------------------------
def quadprog_solve_qp(G, a, C=None, b=None, meq: int = 0):
    """Same signature, return tuple and ValueError behaviour as quadprog.solve_qp (quadprog==0.1.8)."""
    G = np.ascontiguousarray(G, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = G.shape[0]
    if C is None:
        C = np.zeros((n, 0))
        b = np.zeros(0)
    C = np.ascontiguousarray(C, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    q = C.shape[1]
    x = np.empty(n)
    lagr = np.zeros(max(q, 1))
    iact = np.zeros(max(q, 1), dtype=np.int32)
    iters = np.zeros(2, dtype=np.int32)
    f = ctypes.c_double(0.0)
    nact = ctypes.c_int(0)
    Cc = C if q > 0 else np.zeros((n, 1))
    bc = b if q > 0 else np.zeros(1)
    rc = lib().chb_oracle_gi_solve(
        n, G, a, q, Cc, bc, int(meq), x, ctypes.byref(f), lagr.ctypes.data, iact.ctypes.data,
        ctypes.byref(nact), iters.ctypes.data,
    )
    if rc == 1:
        raise ValueError("constraints are inconsistent, no solution")
    if rc == 2:
        raise ValueError("matrix G is not positive definite")
    if rc != 0:
        raise MemoryError("oracle GI solver: allocation failed")
    xu = np.linalg.solve(G, a)
    return x, f.value, xu, iters.copy(), lagr[:q].copy(), (iact[: nact.value] + 1).copy()


def simplex_qp(P, q) -> np.ndarray:
    """Independent exact solver (Wolfe min-norm point) for: min 1/2 a'Pa + q'a, a >= 0, sum a = 1."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    alpha = np.empty(P.shape[0])
    lib().chb_oracle_simplex_qp(P.shape[0], P, q, alpha)
    return alpha


# ----------------------------------------------------------------------------------------------------------
# distance_matrix.py
# ----------------------------------------------------------------------------------------------------------
def create_in_mem_distance_matrix(arr: np.ndarray) -> np.ndarray:
    """distance_matrix.py:33-44 (scipy cdist 'euclidean' recipe, bit-exact)."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    n, d = arr.shape
    out = np.empty((n, n))
    lib().chb_oracle_cdist(arr, n, d, out)
    return out


def distance_rows(arr: np.ndarray, rows: np.ndarray) -> np.ndarray:
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    n, d = arr.shape
    out = np.empty((len(rows), n))
    lib().chb_oracle_cdist_rows(arr, n, d, rows, len(rows), out)
    return out


def find_nearest_from_cluster(c: int, curr_bins: np.ndarray, distance_row: np.ndarray, m: int) -> np.ndarray:
    """distance_matrix.py:47-62; canonical (distance, index) order instead of argpartition's unspecified one."""
    curr_bins = np.ascontiguousarray(curr_bins, dtype=np.int64)
    distance_row = np.ascontiguousarray(distance_row, dtype=np.float64)
    out = np.empty(max(m, 1), dtype=np.int64)
    cnt = lib().chb_oracle_find_nearest(int(c), curr_bins, len(curr_bins), distance_row, int(m), out)
    return out[:cnt].copy()


# ----------------------------------------------------------------------------------------------------------
# hull_distance.py
# ----------------------------------------------------------------------------------------------------------
def convex_hull_distance(query: np.ndarray, points: np.ndarray, return_alpha: bool = False):
    """hull_distance.py:7-35 through the GI restatement (min-norm fallback replaces cvxopt)."""
    query = np.ascontiguousarray(query, dtype=np.float64)
    points = np.ascontiguousarray(points, dtype=np.float64)
    m, d = points.shape
    st = ctypes.c_int32(0)
    alpha = np.empty(max(m, 1))
    dist = lib().chb_oracle_convex_hull_distance(query, points, m, d, alpha.ctypes.data, ctypes.byref(st))
    if return_alpha:
        return dist, alpha[:m], st.value
    return dist


def affine_hull_distance_qp(query: np.ndarray, points: np.ndarray) -> float:
    """hull_distance.py:38-66."""
    query = np.ascontiguousarray(query, dtype=np.float64)
    points = np.ascontiguousarray(points, dtype=np.float64)
    m, d = points.shape
    st = ctypes.c_int32(0)
    return lib().chb_oracle_affine_hull_distance_qp(query, points, m, d, ctypes.byref(st))


def affine_hull_distance(query: np.ndarray, points: np.ndarray) -> float:
    """hull_distance.py:69-87: project (query - mean) onto the orthogonal complement of span(points - mean).  The basis
    is scipy.linalg.orth's: left singular vectors whose singular value exceeds max(shape) * eps * s_max."""
    query = np.asarray(query, dtype=np.float64)
    points = np.asarray(points, dtype=np.float64)
    mean_vec = points.mean(axis=0)
    rel = (points - mean_vec).T  # d x m
    u, sv, _ = np.linalg.svd(rel, full_matrices=False)
    tol = max(rel.shape) * np.finfo(np.float64).eps * (sv.max() if sv.size else 0.0)
    basis = u[:, sv > tol]
    dq = query - mean_vec
    proj = basis @ np.linalg.solve(basis.T @ basis, basis.T @ dq) if basis.shape[1] else np.zeros_like(dq)
    return float(np.linalg.norm(dq - proj))


def hull_distance_batch(samples: np.ndarray, queries: np.ndarray, idx: np.ndarray, m: np.ndarray, metric: str = "convex",
                        threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """calculate_distance (hull_distance.py:90-108) for many (query, neighbour list) pairs at once: queries (P,) point
    indices, idx (P, k) neighbour indices, m (P,) how many of them count.  Returns (distances (P,), status (P,))."""
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    m = np.ascontiguousarray(m, dtype=np.int32)
    P, k = idx.shape
    dist = np.empty(P)
    status = np.empty(P, dtype=np.int32)
    met = {"convex": 0, "affine-qp": 1, "affine": 1}[metric]
    lib().chb_oracle_hull_distance_batch(samples, samples.shape[1], queries, idx, m, P, k, met, int(threads), dist, status)
    return dist, status


def calculate_distance(x: np.ndarray, mat_p: np.ndarray, qp_solver: str = "quadprog", metric: str = "convex") -> float:
    """hull_distance.py:90-108."""
    if metric == "convex":
        return convex_hull_distance(x, mat_p)
    if metric == "affine":
        return affine_hull_distance(x, mat_p)
    if metric == "affine-qp":
        return affine_hull_distance_qp(x, mat_p)
    raise NotImplementedError(f"Metric {metric} not implemented")


# ----------------------------------------------------------------------------------------------------------
# algorithm.py
# ----------------------------------------------------------------------------------------------------------
def draw_permutations(initial_bins: np.ndarray, max_iterations: int, seed: Optional[int] = 0) -> np.ndarray:
    """The RNG contract of the reference: np.random.seed(0) (ch_bin/ch_bin.py:22) then ONE
    np.random.permutation(points_to_assign) per executed iteration (algorithm.py:45).  Drawing all
    max_iterations rows up-front consumes the same stream prefix for the iterations that do execute."""
    pts = np.where(np.asarray(initial_bins) == -1)[0]
    if seed is not None:
        np.random.seed(seed)
    return np.stack([np.random.permutation(pts) for _ in range(max_iterations)]).astype(np.int64).reshape(
        max_iterations, len(pts)
    )


def fit_cluster(
    samples: np.ndarray,
    num_clusters: int,
    initial_bins: np.ndarray,
    distance_matrix: Optional[np.ndarray] = None,
    num_neighbors: int = 15,
    max_iterations: int = 10,
    metric: str = "convex",
    perms: Optional[np.ndarray] = None,
    threads: int = 1,
    max_steps: int = -1,
    return_info: bool = False,
):
    """algorithm.py:12-76 (sequential; `threads` > 1 only parallelises the C bins of one step)."""
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    initial_bins = np.ascontiguousarray(initial_bins, dtype=np.int64)
    n, d = samples.shape
    if perms is None:
        perms = draw_permutations(initial_bins, max_iterations, seed=None)
    perms = np.ascontiguousarray(perms, dtype=np.int64)
    U = perms.shape[1]
    # "affine" and "affine-qp" are the same geometric distance (see affine_hull_distance): the sequential loop uses one solver
    met = {"convex": 0, "affine-qp": 1, "affine": 1}.get(metric)
    if met is None:
        raise NotImplementedError(f"Metric {metric} not implemented")
    if distance_matrix is not None:
        distance_matrix = np.ascontiguousarray(distance_matrix, dtype=np.float64)
        dptr = distance_matrix.ctypes.data
    else:
        dptr = None
    labels = np.empty(n, dtype=np.int64)
    iters = ctypes.c_int32(0)
    conv = ctypes.c_int32(0)
    nqp = ctypes.c_int64(0)
    changed = np.zeros(max(max_iterations, 1), dtype=np.int64)
    rc = lib().chb_oracle_fit_cluster(
        samples, n, d, int(num_clusters), initial_bins, dptr, int(num_neighbors), int(max_iterations), met,
        perms, U, int(threads), int(max_steps), labels, ctypes.byref(iters), ctypes.byref(conv), changed,
        ctypes.byref(nqp),
    )
    if rc != 0:
        raise ValueError(f"oracle fit_cluster failed rc={rc}")
    if return_info:
        return labels, dict(iterations=iters.value, converged=bool(conv.value),
                            changed=changed[: iters.value].copy(), qps=nqp.value)
    return labels


def verify_iteration(
    samples: np.ndarray,
    num_clusters: int,
    old_labels: np.ndarray,
    new_labels: np.ndarray,
    perm: np.ndarray,
    num_neighbors: int,
    positions: Optional[np.ndarray] = None,
    metric: str = "convex",
    threads: int = 0,
    return_distances: bool = False,
):
    """Position-parallel check that `new_labels` is what ONE iteration of algorithm.py:46-60 makes of `old_labels` under
    the permutation `perm` (oracle/verify.c): at every checked position p the oracle recomputes
    assign(perm[p] | new labels of positions < p, old labels of positions > p and of the seeds) and compares.  All U
    positions checked (positions=None) and none failing <=> new_labels is the sequential result.  Returns a dict:
    mismatches (count), positions, labels (the oracle's), best / second (its two smallest hull distances per position:
    best == second marks a tie the strict '<' of algorithm.py:57 decides)."""
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    old_labels = np.ascontiguousarray(old_labels, dtype=np.int64)
    new_labels = np.ascontiguousarray(new_labels, dtype=np.int64)
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    n, d = samples.shape
    if positions is None:
        positions = np.arange(len(perm), dtype=np.int64)
    positions = np.ascontiguousarray(positions, dtype=np.int64)
    met = {"convex": 0, "affine-qp": 1, "affine": 1}[metric]
    lab = np.empty(len(positions), dtype=np.int64)
    best = np.empty(len(positions))
    second = np.empty(len(positions))
    dist = np.empty((len(positions), int(num_clusters))) if return_distances else None
    rc = lib().chb_oracle_verify_positions(
        samples, n, d, int(num_clusters), old_labels, new_labels, perm, len(perm), positions, len(positions),
        int(num_neighbors), met, int(threads), lab, best, second, dist.ctypes.data if dist is not None else None,
    )
    if rc < 0:
        raise ValueError(f"oracle verify_positions failed rc={rc}")
    out = dict(mismatches=int(rc), positions=positions, labels=lab, best=best, second=second)
    if dist is not None:
        out["distances"] = dist
    return out


# ---------------------------------------------------------------------------------------------------------
# Coverage normalisation and feature merge
def pairwise_sum(a) -> float:
    """numpy's add.reduce over a contiguous float64 axis, restated (this is what `DataFrame.sum(axis=0)` of
    coverage.py:37 runs per column): below 8 elements a plain running sum; up to 128 elements eight running sums combined
    as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and then the tail; above, halves split at a multiple of 8.  Pure Python loop --
    small inputs only; tests pin it against numpy itself."""
    n = len(a)
    if n < 8:
        res = -0.0
        for v in a:
            res += float(v)
        return res
    if n <= 128:
        r = [float(a[u]) for u in range(8)]
        t = 8
        while t < n - (n % 8):
            for u in range(8):
                r[u] += float(a[t + u])
            t += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while t < n:
            res += float(a[t])
            t += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def normalise_coverages(raw: np.ndarray) -> np.ndarray:
    """coverage.py:35-41: divide each column by its sum, then (more than one sample) each row by its sum.
    Column sums: pairwise (contiguous axis of the frame's block); row sums: left to right over the columns."""
    raw = np.asarray(raw, dtype=np.float64)
    colsum = np.ascontiguousarray(raw.T).sum(axis=1)
    t = raw / colsum
    if raw.shape[1] > 1:
        rs = t[:, 0].copy()
        for j in range(1, raw.shape[1]):
            rs = rs + t[:, j]
        t = t / rs[:, None]
    return t


def merge_features(kmer: np.ndarray, cov_norm: np.ndarray, parent: np.ndarray) -> np.ndarray:
    """cli/features.py:106-109 + cli/clustering.py:53: samples = [k-mer columns | coverage columns of the parent contig]."""
    return np.ascontiguousarray(np.hstack([np.asarray(kmer, dtype=np.float64), np.asarray(cov_norm)[np.asarray(parent)]]))
