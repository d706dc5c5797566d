"""
Coverage normalisation and feature merge on the device.

Mirrors the two reference steps that turn an abundance file plus seq2vec k-mer profiles into the `samples` matrix of
`fit_cluster`:
  * `parse_coverages`  -- ch_bin/core/features/coverage.py:13-43 (same name, arguments and returned frame),
  * the merge          -- ch_bin/cli/features.py:106-109 followed by the column drop of ch_bin/cli/clustering.py:53.
The arithmetic (column sums in pandas' pairwise order, row sums left to right, the two divisions, the gather by parent
contig) runs in libchbin_b200 (`chb_set_features_merged`); file parsing and the name join stay on the host, in pandas, as
in the reference.  k-mer counting (seq2vec) and single-copy-marker seeding stay external.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

from . import capi


def read_coverages(coverage_file: Path, delimiter: str = "\t") -> Tuple[np.ndarray, np.ndarray]:
    """The raw table of coverage.py:30-32: (contig names (P,), coverages (P, S) float64), parsed exactly as the reference
    parses it (pandas' default float parser), not normalised."""
    import pandas as pd

    df = pd.read_csv(coverage_file, sep=delimiter, header=None)
    names = df[0].to_numpy()
    raw = np.ascontiguousarray(df.drop(columns=[0]).to_numpy(dtype=np.float64))
    if raw.ndim != 2 or raw.shape[1] == 0:
        raise ValueError(f"{coverage_file}: no coverage columns")
    if np.isnan(raw).any():
        raise ValueError(f"{coverage_file}: missing coverage values are not supported on the device path")
    return names, raw


def normalise_coverages(raw: np.ndarray, device: int = 0) -> np.ndarray:
    """coverage.py:35-41 on the device: (P, S) raw coverages -> normalised, the same doubles pandas produces."""
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    with capi.Context(device) as ctx:
        return ctx.set_features_merged(None, raw, np.arange(raw.shape[0], dtype=np.int64), want_coverages=True)


def parse_coverages(coverage_file: Path, delimiter: str = "\t", device: int = 0):
    """Drop-in for coverage.py:13 `parse_coverages`: a frame with CONTIG_NAME and one normalised column per sample."""
    import pandas as pd

    names, raw = read_coverages(coverage_file, delimiter)
    norm = normalise_coverages(raw, device)
    df = pd.DataFrame(norm, columns=list(range(1, raw.shape[1] + 1)))
    df.insert(0, "CONTIG_NAME", names)
    return df


def parent_rows(parent_names: Sequence[str], coverage_names: Sequence[str]) -> np.ndarray:
    """Row of the coverage table each sub-contig inherits: the PARENT_NAME == CONTIG_NAME join of cli/features.py:107.
    The reference's inner merge silently drops sub-contigs without a coverage row; here that is an error, because the
    caller's row order (and `initial_bins`) would no longer match."""
    pos = {}
    for i, name in enumerate(coverage_names):
        pos.setdefault(name, i)
    try:
        return np.fromiter((pos[p] for p in parent_names), dtype=np.int64, count=len(parent_names))
    except KeyError as e:
        raise ValueError(f"sub-contig parent {e.args[0]!r} has no row in the coverage file") from None


def merged_samples(kmer: np.ndarray, cov_raw: np.ndarray, parent: np.ndarray, device: int = 0,
                   ctx: Optional["capi.Context"] = None) -> np.ndarray:
    """`samples` = [k-mer profile | normalised coverage of the parent contig], (n, dk + S) float64, C-contiguous.
    With `ctx` the matrix also stays resident in that context (as after `set_features`)."""
    if ctx is not None:
        ctx.set_features_merged(kmer, cov_raw, parent)
        return ctx.get_features()
    with capi.Context(device) as own:
        own.set_features_merged(kmer, cov_raw, parent)
        return own.get_features()


def samples_from_files(sub_contig_parents: Sequence[str], kmer: np.ndarray, coverage_file: Path, delimiter: str = "\t",
                       device: int = 0) -> np.ndarray:
    """From an abundance file and per-sub-contig k-mer profiles to `samples`, without the features.csv round trip."""
    names, raw = read_coverages(coverage_file, delimiter)
    return merged_samples(kmer, raw, parent_rows(sub_contig_parents, names), device)
