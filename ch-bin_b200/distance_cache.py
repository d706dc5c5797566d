"""
On-disk distance matrix and feature side-car: the data formats either side of the hot path.

    create_distance_matrix(arr, operating_dir)   <- /root/reference/ch_bin/core/clustering/distance_matrix.py:12-30
        Same file (operating_dir/"distance_matrix.npy"), same format (NumPy .npy, (n, n) '<f8', C order, written through
        numpy.lib.format.open_memmap exactly as the reference does), same reuse rule and log lines -- but the rows come
        from distance.cu, which reproduces scipy's cdist doubles bit for bit, so a later run of the reference
        (InMemDistMatrix = no, cli/clustering.py:61-63) reuses a cache it could have written itself.
    validate_distance_matrix(path, arr)          spot-check of an existing cache against recomputed rows (the reference
        trusts a stale file blindly: "Assuming memmap shape"; the B200 path never reads distances from disk).
    write_features_sidecar / load_samples        binary copy of the `samples` block of features.csv
        (cli/clustering.py:47-53): parsing a 1M x 160 CSV costs far more than the clustering stage on the GPU.

The clustering stage itself never needs the n x n matrix (distance mode 2 regenerates candidate distances on the tensor
cores every round); this module exists so that the B200 path produces and honours the same artefacts as the reference.
"""
from __future__ import annotations

import logging
import os
import time
from pathlib import Path
from typing import Optional

import numpy as np
from numpy.lib.format import open_memmap

from . import capi

logger = logging.getLogger(__name__)

DISTANCE_MATRIX_NAME = "distance_matrix.npy"  # distance_matrix.py:18
SIDECAR_SUFFIX = ".samples.npy"
META_COLUMNS = ["CONTIG_NAME", "PARENT_NAME", "CLUSTER"]  # cli/clustering.py:53


def _row_context(arr: np.ndarray, device: int) -> "capi.Context":
    """A context in which every point is a query slot, so that chb_get_distance_rows serves any row."""
    ctx = capi.Context(device)
    ctx.set_features(arr)
    ctx.set_labels(np.full(len(arr), -1, dtype=np.int64), 1)
    ctx.set_params(1, "convex")
    ctx.set_distance_mode(0)  # exact scipy-recipe rows (distance.cu)
    ctx.build_distance_matrix(False)
    return ctx


def create_distance_matrix(arr: np.ndarray, operating_dir: Path, device: int = 0, chunk_bytes: int = 1 << 28) -> Path:
    """distance_matrix.py:12-30 with the rows computed on the GPU.  Returns the file name; reuses an existing file."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    n = len(arr)
    operating_dir = Path(operating_dir)
    filename = operating_dir / DISTANCE_MATRIX_NAME
    if filename.exists():
        logger.info("Reusing already existing distance matrix at %s.", filename)
        logger.debug("Assuming memmap shape %s", (n, n))
        return filename
    operating_dir.mkdir(parents=True, exist_ok=True)
    start_time = time.time()
    logger.debug("Started creating distance matrix at %s.", filename)
    result = open_memmap(filename=filename, mode="w+", shape=(n, n))
    rows_per_chunk = max(1, min(n, chunk_bytes // (8 * max(n, 1))))
    with _row_context(arr, device) as ctx:
        for r0 in range(0, n, rows_per_chunk):
            cnt = min(rows_per_chunk, n - r0)
            result[r0:r0 + cnt] = ctx.get_distance_rows(r0, cnt)
    result.flush()
    logger.debug("Ended creating distance matrix. Shape is %s", result.shape)
    logger.debug("Distance matrix calculated in %s s.", time.time() - start_time)
    del result
    return filename


def validate_distance_matrix(filename: Path, arr: np.ndarray, rows: int = 8, device: int = 0, seed: int = 0) -> bool:
    """True iff the file has the (n, n) float64 layout and `rows` randomly chosen rows equal the recomputed ones bit
    for bit.  A cache written for another feature matrix fails this check."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    n = len(arr)
    try:
        mm = np.load(filename, mmap_mode="r")
    except Exception:
        return False
    if mm.shape != (n, n) or mm.dtype != np.float64:
        return False
    pick = np.sort(np.random.default_rng(seed).choice(n, size=min(rows, n), replace=False))
    with _row_context(arr, device) as ctx:
        for r in pick:
            if not np.array_equal(np.asarray(mm[r]), ctx.get_distance_rows(int(r), 1)[0]):
                return False
    return True


def sidecar_path(features_csv: Path) -> Path:
    features_csv = Path(features_csv)
    return features_csv.with_name(features_csv.name + SIDECAR_SUFFIX)


def write_features_sidecar(features_csv: Path) -> Path:
    """Parses features.csv once the way cli/clustering.py:47-53 does and stores the float64 `samples` block next to it."""
    import pandas as pd

    df = pd.read_csv(features_csv)
    samples = df.drop(META_COLUMNS, axis=1).values
    out = sidecar_path(features_csv)
    tmp = out.with_name(out.name + ".tmp")
    with open(tmp, "wb") as f:
        np.save(f, np.ascontiguousarray(samples, dtype=np.float64))
    os.replace(tmp, out)
    return out


def load_samples(features_csv: Path, num_rows: Optional[int] = None) -> Optional[np.ndarray]:
    """The side-car's array if it exists, is not older than the CSV and has `num_rows` rows; otherwise None."""
    sc = sidecar_path(features_csv)
    try:
        if not sc.exists() or sc.stat().st_mtime < Path(features_csv).stat().st_mtime:
            return None
        arr = np.load(sc)
    except Exception:
        return None
    if arr.ndim != 2 or arr.dtype != np.float64 or (num_rows is not None and arr.shape[0] != num_rows):
        return None
    return arr
