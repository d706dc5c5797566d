"""
CPU tests of the coverage-normalisation / feature-merge row: the oracle restatement against the
outputs of the reference's own `parse_coverages` (coverage.py:13-43, run unmodified by tests/golden/make_golden.py section 6)
and against the pandas merge of cli/features.py:106-109; plus the host-side name join.
"""
import os

import numpy as np
import pytest

import oracle
from chbin_b200 import features

COV_KEYS = ["real", "7_2", "129_3", "300_10", "1100_20"]


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_golden.npz"))


@pytest.mark.parametrize("key", COV_KEYS)
def test_oracle_normalisation_is_bit_identical_to_parse_coverages(G, key):
    raw, want = G[f"cov_raw_{key}"], G[f"cov_norm_{key}"]
    got = oracle.normalise_coverages(raw)
    assert got.shape == want.shape and np.array_equal(got, want)
    if raw.shape[1] > 1:  # rows of a multi-sample table sum to one (coverage.py:39)
        assert np.allclose(got.sum(axis=1), 1.0, rtol=0, atol=1e-14)
    else:  # a single sample is normalised over the column only (coverage.py:38)
        assert abs(got.sum() - 1.0) < 1e-12


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 15, 16, 100, 127, 128, 129, 136, 137, 255, 256, 257, 519, 1000, 1100, 4099])
def test_pairwise_sum_restatement_matches_numpy(n):
    rng = np.random.default_rng(n)
    a = rng.lognormal(3.0, 2.0, n)
    assert oracle.pairwise_sum(a) == float(np.sum(a))  # bit-for-bit: same association order


def test_row_sum_order_matters_and_is_sequential(G):
    # the pin is meaningful: summing the rows pairwise (as numpy would over a contiguous row) gives other doubles
    raw, want = G["cov_raw_1100_20"], G["cov_norm_1100_20"]
    t = raw / np.ascontiguousarray(raw.T).sum(axis=1)
    other = t / np.ascontiguousarray(t).sum(axis=1)[:, None]
    assert not np.array_equal(other, want)


def test_oracle_merge_matches_pandas_merge(G):
    P, S = (int(v) for v in G["merge_cov_key"])
    cov = G[f"cov_norm_{P}_{S}"]
    got = oracle.merge_features(G["merge_kmer"], cov, G["merge_parent"])
    assert np.array_equal(G["merge_order"], np.arange(len(got)))  # the reference's merge kept the sub-contig order
    assert np.array_equal(got, G["merge_samples"])
    assert got.flags["C_CONTIGUOUS"]


def test_parent_rows_join():
    names = np.array(["c0", "c1", "c2", "c1"], dtype=object)  # a repeated name resolves to its first row
    idx = features.parent_rows(["c2", "c0", "c1", "c2"], names)
    assert idx.dtype == np.int64 and idx.tolist() == [2, 0, 1, 2]
    with pytest.raises(ValueError, match="no row in the coverage file"):
        features.parent_rows(["c0", "missing"], names)
    assert features.parent_rows([], names).shape == (0,)


def test_read_coverages_parses_like_the_reference(tmp_path):
    import pandas as pd

    fn = tmp_path / "cov.tsv"
    fn.write_text("a\t1.5\t2.25\nb\t3.0\t0.125\nc\t10\t20\n")
    names, raw = features.read_coverages(fn)
    assert names.tolist() == ["a", "b", "c"] and raw.dtype == np.float64 and raw.flags["C_CONTIGUOUS"]
    assert np.array_equal(raw, pd.read_csv(fn, sep="\t", header=None).drop(columns=[0]).values)
    csv = tmp_path / "cov.csv"
    csv.write_text("a,1.5\nb,2.5\n")
    assert features.read_coverages(csv, delimiter=",")[1].shape == (2, 1)
    bad = tmp_path / "bad.tsv"
    bad.write_text("a\t1.5\t\nb\t3.0\t1.0\n")
    with pytest.raises(ValueError, match="missing coverage"):
        features.read_coverages(bad)
    only_names = tmp_path / "names.tsv"
    only_names.write_text("a\nb\n")
    with pytest.raises(ValueError, match="no coverage columns"):
        features.read_coverages(only_names)
