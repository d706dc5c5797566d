#!/usr/bin/env python
"""Times chb_set_features_merged (coverage normalisation + [k-mer | coverage] merge on the device, host buffers in, matrix
resident) next to the pandas expressions of coverage.py:36-39 on the host.  Usage: python tools/features_time.py [n P S dk]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chbin_b200 import capi  # noqa: E402


def main():
    n, P, S, dk = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (1_000_000, 400_000, 20, 136)
    rng = np.random.default_rng(0)
    kmer = rng.random((n, dk))
    raw = rng.lognormal(3.0, 1.0, (P, S))
    parent = rng.integers(0, P, n)
    with capi.Context(0) as ctx:
        for _ in range(2):
            ctx.set_features_merged(kmer, raw, parent)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.set_features_merged(kmer, raw, parent)
        t_dev = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.set_features(np.empty((n, dk + S)))
        t_up = (time.perf_counter() - t0) / 5
    import pandas as pd

    df = pd.DataFrame(raw)
    t0 = time.perf_counter()
    t = df.div(df.sum(axis=0), axis=1)
    t = t.div(t.sum(axis=1), axis=0)
    merged = np.hstack([kmer, t.values[parent]])
    t_host = time.perf_counter() - t0
    print(f"n={n} P={P} S={S} dk={dk}: device merged set-up {t_dev * 1e3:.1f} ms (plain upload of the finished matrix "
          f"{t_up * 1e3:.1f} ms); pandas normalise + numpy merge on the host {t_host * 1e3:.1f} ms; merged bytes {merged.nbytes / 1e6:.0f} MB")


if __name__ == "__main__":
    main()
