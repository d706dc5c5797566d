"""
GPU tests of the rows either side of the hot path: the on-disk distance matrix in the reference's
format, and perform_clustering (cli/clustering.py:19-99) end to end through libchbin_b200.so -- pinned against the files
the reference's own perform_clustering wrote (tests/golden/make_golden.py section 5).
"""
import os

import numpy as np
import pytest
from numpy.lib.format import open_memmap

import chbin_b200
import oracle
from chbin_b200 import distance_cache, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_golden.npz"))


def test_distance_matrix_file_matches_reference_file(G, tmp_path):
    import pandas as pd

    csv = tmp_path / "features.csv"
    csv.write_bytes(G["pc_features_csv"].tobytes())
    X = pd.read_csv(csv).drop(distance_cache.META_COLUMNS, axis=1).values
    fn = distance_cache.create_distance_matrix(X, tmp_path / "op")
    assert fn == tmp_path / "op" / "distance_matrix.npy"  # distance_matrix.py:18
    raw = fn.read_bytes()
    assert raw[:128] == G["pc_npy_header"].tobytes()  # same .npy header as the file the reference wrote
    mm = open_memmap(filename=fn, mode="r", shape=(len(X), len(X)))  # how cli/clustering.py:63 opens it
    assert np.array_equal(mm[[0, 57, 179]], G["pc_dm_rows"])  # bit-identical to the reference's rows
    assert np.array_equal(np.asarray(mm), oracle.create_in_mem_distance_matrix(np.ascontiguousarray(X)))
    # reuse rule (distance_matrix.py:19-22): an existing file is returned untouched
    stamp = fn.stat().st_mtime_ns
    assert distance_cache.create_distance_matrix(X, tmp_path / "op") == fn and fn.stat().st_mtime_ns == stamp
    assert distance_cache.validate_distance_matrix(fn, X)
    X2 = X.copy()
    X2[57, 3] += 1e-9
    assert not distance_cache.validate_distance_matrix(fn, X2, rows=len(X))


def test_distance_matrix_file_chunked_rows(tmp_path):
    X, _, _ = synth.make_contig_features(700, 4, 3, 10, seed=3)
    fn = distance_cache.create_distance_matrix(X, tmp_path, chunk_bytes=700 * 8 * 33)  # 33 rows per chunk, ragged tail
    assert np.array_equal(np.load(fn), oracle.create_in_mem_distance_matrix(X))


@pytest.mark.parametrize("in_mem", [True, False])
@pytest.mark.parametrize("sidecar", [False, True])
def test_perform_clustering_end_to_end_matches_reference_files(G, tmp_path, in_mem, sidecar):
    csv = tmp_path / "features.csv"
    csv.write_bytes(G["pc_features_csv"].tobytes())
    if sidecar:
        distance_cache.write_features_sidecar(csv)
    k, iters = (int(v) for v in G["pc_params"])
    np.random.seed(0)  # ch_bin/ch_bin.py:22
    out = chbin_b200.perform_clustering(None, csv, tmp_path / "op", k, iters, "convex", "b200", in_mem)
    assert out.read_bytes() == G["pc_assignment_csv"].tobytes()
    assert (tmp_path / "op" / "distance_matrix.npy").exists() == (not in_mem)  # cli/clustering.py:57-63
