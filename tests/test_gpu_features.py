"""
GPU tests of chb_set_features_merged / chb_get_features: coverage normalisation and the
[k-mer | coverage] merge on the device, bit-for-bit against the reference's own parse_coverages output and the pandas merge
(tests/golden/make_golden.py section 6), and against the oracle at sizes where the pairwise tree is deep.
"""
import os

import numpy as np
import pytest

import chbin_b200
import oracle
from chbin_b200 import capi, features, synth

pytestmark = pytest.mark.gpu

COV_KEYS = ["real", "7_2", "129_3", "300_10", "1100_20"]


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_golden.npz"))


@pytest.mark.parametrize("key", COV_KEYS)
def test_device_normalisation_is_bit_identical_to_parse_coverages(G, key):
    got = features.normalise_coverages(G[f"cov_raw_{key}"])
    assert np.array_equal(got, G[f"cov_norm_{key}"])


def test_device_merge_matches_pandas_merge(G):
    P, S = (int(v) for v in G["merge_cov_key"])
    got = features.merged_samples(G["merge_kmer"], G[f"cov_raw_{P}_{S}"], G["merge_parent"])
    assert got.flags["C_CONTIGUOUS"] and np.array_equal(got, G["merge_samples"])


@pytest.mark.parametrize("P,S", [(1, 1), (5, 4), (8, 1), (64, 2), (128, 7), (136, 2), (4099, 33), (250_000, 20), (1_000_003, 3)])
def test_device_normalisation_matches_oracle_at_depth(P, S):
    rng = np.random.default_rng(P + S)
    raw = rng.lognormal(3.0, 1.5, (P, S))
    if P > 100:
        raw[rng.integers(0, P, 10)] *= 1e6  # a few dominant contigs: the association order shows in the last bits
    got = features.normalise_coverages(raw)
    assert np.array_equal(got, oracle.normalise_coverages(raw))


def test_merged_context_then_fit_equals_fit_on_host_merged_samples():
    X, bins, _ = synth.make_contig_features(1500, 6, 3, 20, seed=5)
    dk = 136
    rng = np.random.default_rng(2)
    P = 900
    parent = np.concatenate([np.arange(P), rng.integers(0, P, len(X) - P)])
    rng.shuffle(parent)
    cov_raw = rng.lognormal(3.0, 1.0, (P, 3))
    kmer = np.ascontiguousarray(X[:, :dk])
    want_samples = oracle.merge_features(kmer, oracle.normalise_coverages(cov_raw), parent)
    with capi.Context(0) as ctx:
        got_samples = features.merged_samples(kmer, cov_raw, parent, ctx=ctx)
        assert np.array_equal(got_samples, want_samples)
        # the matrix is resident: cluster straight from it, no second upload
        perms = chbin_b200.draw_permutations(bins, 4)
        ctx.set_labels(bins, 6)
        ctx.set_params(5, "convex")
        ctx.build_distance_matrix(False)
        labels_resident = ctx.fit(perms, 4)[0]
    np.random.seed(0)  # the same permutation stream as draw_permutations above (ch_bin.py:22)
    labels_uploaded = chbin_b200.fit_cluster(want_samples, 6, bins, num_neighbors=5, max_iterations=4)
    assert np.array_equal(labels_resident, labels_uploaded)
    assert np.array_equal(labels_uploaded, oracle.fit_cluster(want_samples, 6, bins, None, 5, 4, perms=perms, threads=4))


def test_parse_coverages_frame_and_files(tmp_path):
    rng = np.random.default_rng(9)
    P, S, n, dk = 400, 5, 700, 12
    vals = np.round(rng.lognormal(3.0, 1.0, (P, S)), 3)
    fn = tmp_path / "abundance.tsv"
    fn.write_text("".join(f"contig_{i}\t" + "\t".join(repr(float(v)) for v in vals[i]) + "\n" for i in range(P)))
    names, raw = features.read_coverages(fn)
    df = features.parse_coverages(fn)
    assert list(df.columns) == ["CONTIG_NAME", 1, 2, 3, 4, 5] and df["CONTIG_NAME"].tolist() == names.tolist()
    assert np.array_equal(df.drop(columns=["CONTIG_NAME"]).values, oracle.normalise_coverages(raw))
    parents = [f"contig_{q}" for q in rng.integers(0, P, n)]
    kmer = rng.dirichlet(np.full(dk, 4.0), n)
    got = features.samples_from_files(parents, kmer, fn)
    want = oracle.merge_features(kmer, oracle.normalise_coverages(raw), features.parent_rows(parents, names))
    assert got.shape == (n, dk + S) and np.array_equal(got, want)


def test_merged_abi_validation():
    raw = np.ones((4, 2))
    with capi.Context(0) as ctx:
        with pytest.raises(ValueError, match="not a coverage row"):
            ctx.set_features_merged(np.ones((3, 2)), raw, np.array([0, 4, 1]))
        with pytest.raises(ValueError, match="not a coverage row"):
            ctx.set_features_merged(np.ones((3, 2)), raw, np.array([0, -1, 1]))
        with pytest.raises(ValueError, match="one row per entry"):
            ctx.set_features_merged(np.ones((2, 2)), raw, np.array([0, 1, 2]))
        with pytest.raises(ValueError):
            ctx._check(ctx._lib.chb_get_features(ctx._h, None))
        out = ctx.set_features_merged(None, raw, np.array([3, 0]), want_coverages=True)  # dk = 0: coverages only
        assert np.array_equal(out, np.full((4, 2), 0.5)) and np.array_equal(ctx.get_features(), np.full((2, 2), 0.5))
