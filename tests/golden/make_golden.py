#!/usr/bin/env python
"""
Generates the committed golden fixtures in tests/golden/ by running the REFERENCE'S OWN modules
(/root/reference/ch_bin/core/clustering/*.py, unmodified, imported through oracle/ref_shim.py) in the build container.

The reference ships no tests and no golden outputs, and its two solver dependencies
(quadprog 0.1.8, cvxopt 1.2.6) are not installable here, so ref_shim injects stand-ins for those two packages only:
everything else on the path -- distance_matrix.py (scipy cdist, numpy argpartition), hull_distance.py, solve_qp.py
(including the numba nearest_positive_definite and the quadprog->cvxopt fallback), algorithm.py (the sequential loop and
its np.random.permutation draws) -- executes verbatim.  The fixtures therefore pin:
  * distances and neighbour sets against scipy/numpy themselves (fully pinned), and
  * hull distances / labels against the verbatim reference flow with the restated GI solver ("parity unpinned" against
    quadprog proper, stated in DESIGN.md).

Run:  python tests/golden/make_golden.py        (needs /root/reference; cannot run on the GPU box)
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import chbin_b200  # noqa: E402
from chbin_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

REF = "/root/reference"


def five_genomes_like(seed=0):
    """SURVEY.md 8(c) config #1 substitute: the REAL coverage column of test_data/five-genomes-abundance.abund,
    ContigLengthFilterBp=1000 and SeedContigSplitLengthBp=10000 of config/default.ini applied to the contig lengths in
    the names, the 5 longest contigs as seeds, synthetic 4-mer profiles (the FASTA is missing from the mount)."""
    names, cov = [], []
    with open(os.path.join(REF, "test_data", "five-genomes-abundance.abund")) as f:
        for line in f:
            a, b = line.split("\t")
            names.append(a)
            cov.append(float(b))
    cov = np.array(cov)
    cov_norm = cov / cov.sum()  # coverage.py:36-38: every column by its sum over ALL rows, before any filtering
    length = np.array([int(re.search(r"length_(\d+)", nm).group(1)) for nm in names])
    keep = np.where(length >= 1000)[0]
    order = keep[np.argsort(-length[keep], kind="stable")]
    seeds = order[:5]
    rng = np.random.default_rng(seed)
    centroids = rng.dirichlet(np.full(synth.KMER_DIMS, 8.0), size=5)
    rows, bins, parents = [], [], []
    seed_cov = np.log(cov[seeds])
    for ci, s in enumerate(seeds):
        pieces = int(length[s] // 10000)  # preprocess.py:17-35
        for _ in range(pieces):
            rows.append(np.concatenate([rng.dirichlet(4000.0 * centroids[ci]), [cov_norm[s]]]))
            bins.append(ci)
            parents.append(int(s))
    for i in keep:
        if i in seeds:
            continue
        g = int(np.argmin(np.abs(np.log(cov[i]) - seed_cov)))  # the genome whose seed has the closest coverage
        rows.append(np.concatenate([rng.dirichlet(1500.0 * centroids[g]), [cov_norm[i]]]))
        bins.append(-1)
        parents.append(int(i))
    return np.ascontiguousarray(np.array(rows)), np.array(bins, dtype=np.int64), np.array(parents, dtype=np.int64)


def main():
    ref = ref_shim.load()
    out = {}
    print("real solvers importable:", ref_shim.REAL_SOLVERS)

    # ---- 1. distances and per-bin neighbours: scipy / numpy are the reference here
    X, bins, truth = synth.make_contig_features(160, 4, 2, 8, seed=21)
    X[7] = X[3]  # exact duplicates
    D = ref.distance_matrix.create_in_mem_distance_matrix(X)
    rng = np.random.default_rng(5)
    labels = truth.copy()
    labels[rng.random(len(labels)) < 0.35] = -1
    queries = rng.choice(len(X), 12, replace=False)
    nn_k, nn_sets = 5, []
    for q in queries:
        lab = labels.copy()
        lab[q] = -1
        for c in range(4):
            idx = ref.distance_matrix.find_nearest_from_cluster(c, lab, D[q], nn_k)
            nn_sets.append(np.sort(idx))
    out.update(knn_X=X, knn_D=D, knn_labels=labels, knn_queries=queries, knn_k=np.int64(nn_k),
               knn_sets=np.array([np.pad(s, (0, nn_k - len(s)), constant_values=-1) for s in nn_sets]))

    # ---- 2. hull distances through hull_distance.py -> solve_qp.py -> (numba nearest-PD) -> solver
    Xh, _, _ = synth.make_contig_features(300, 5, 3, 10, seed=22)
    hq, hidx, hm, hd, ha, hs = [], [], [], [], [], []
    KMAXG = 12
    for t in range(160):
        m = int(rng.integers(1, KMAXG + 1))
        q = int(rng.integers(len(Xh)))
        idx = rng.choice(np.setdiff1d(np.arange(len(Xh)), [q]), m, replace=False)
        hq.append(q)
        hm.append(m)
        hidx.append(np.pad(idx, (0, KMAXG - m), constant_values=-1))
        hd.append(ref.hull_distance.convex_hull_distance(Xh[q], Xh[idx], "quadprog"))
        ha.append(ref.hull_distance.affine_hull_distance_qp(Xh[q], Xh[idx], "quadprog") if m >= 2 else np.nan)
        hs.append(ref.hull_distance.affine_hull_distance(Xh[q], Xh[idx]) if m >= 2 else np.nan)  # SVD form, :69-87
    out.update(hull_X=Xh, hull_q=np.array(hq), hull_idx=np.array(hidx), hull_m=np.array(hm), hull_dist=np.array(hd),
               hull_affine_qp=np.array(ha), hull_affine=np.array(hs))

    # ---- 3. fit_cluster, verbatim reference loop, np.random.seed(0) as ch_bin/ch_bin.py:22
    cases = {
        "easy": dict(n=400, C=4, S=1, n_seed=25, k=5, conc=4000.0, iters=10),
        "hard": dict(n=500, C=5, S=1, n_seed=20, k=5, conc=60.0, iters=10),
        "k10": dict(n=450, C=3, S=10, n_seed=30, k=10, conc=300.0, iters=6),
        "smallbins": dict(n=300, C=4, S=3, n_seed=3, k=7, conc=500.0, iters=10),
    }
    for name, c in cases.items():
        Xc, bc, _ = synth.make_contig_features(c["n"], c["C"], c["S"], c["n_seed"], seed=31, concentration=c["conc"])
        lab = ref_shim.fit_cluster_reference(Xc, c["C"], bc, c["k"], c["iters"], seed=0)
        out[f"fit_{name}_X"] = Xc
        out[f"fit_{name}_bins"] = bc
        out[f"fit_{name}_labels"] = lab.astype(np.int64)
        out[f"fit_{name}_params"] = np.array([c["C"], c["k"], c["iters"]], dtype=np.int64)
        print(name, "changed vs seeds:", int(np.sum(lab != bc)))

    # ---- 3b. the "affine" metric (hull_distance.py:69-87, scipy.linalg.orth) through the same verbatim loop
    ca = dict(n=350, C=4, S=2, n_seed=20, k=4, conc=800.0, iters=6)
    Xa, ba, _ = synth.make_contig_features(ca["n"], ca["C"], ca["S"], ca["n_seed"], seed=33, concentration=ca["conc"])
    laba = ref_shim.fit_cluster_reference(Xa, ca["C"], ba, ca["k"], ca["iters"], metric="affine", seed=0)
    out.update(fit_affine_X=Xa, fit_affine_bins=ba, fit_affine_labels=laba.astype(np.int64),
               fit_affine_params=np.array([ca["C"], ca["k"], ca["iters"]], dtype=np.int64))
    print("affine changed vs seeds:", int(np.sum(laba != ba)))

    # ---- 4. five-genomes-like (config #1)
    Xg, bg, parents = five_genomes_like()
    print("five-genomes-like: n =", len(Xg), "U =", int(np.sum(bg == -1)), "seed pieces =", np.bincount(bg[bg >= 0]))
    lab = ref_shim.fit_cluster_reference(Xg, 5, bg, 5, 10, seed=0)
    out.update(fg_X=Xg, fg_bins=bg, fg_parents=parents, fg_labels=lab.astype(np.int64),
               fg_params=np.array([5, 5, 10], dtype=np.int64))

    # ---- 5. the caller: perform_clustering (cli/clustering.py:19-99) steps 01-04, CSV in -> CSV out, both matrix modes
    import tempfile
    import pandas as pd
    from pathlib import Path

    cli = ref_shim.load_cli_clustering()
    Xp, bp, _ = synth.make_contig_features(180, 3, 1, 12, seed=41, concentration=900.0)
    prng = np.random.default_rng(7)
    parents = np.empty(len(Xp), dtype=object)
    # seeds of a bin are pieces of one parent contig; the rest are grouped 1-3 sub-contigs per parent
    for c in range(3):
        parents[bp == c] = f"seed_parent_{c}"
    rest = np.where(bp == -1)[0]
    pid, i = 0, 0
    while i < len(rest):
        g = int(prng.integers(1, 4))
        parents[rest[i:i + g]] = f"contig_{pid}"
        pid += 1
        i += g
    df = pd.DataFrame({"CONTIG_NAME": [f"sub_{j}" for j in range(len(Xp))], "PARENT_NAME": parents, "CLUSTER": bp})
    df = pd.concat([df, pd.DataFrame(Xp, columns=[f"F{j}" for j in range(Xp.shape[1])])], axis=1)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        csv = td / "features.csv"
        df.to_csv(csv, index=False)
        np.random.seed(0)
        out_mem = cli.perform_clustering(td / "no.fasta", csv, td / "mem", 5, 6, "convex", "quadprog", True)
        np.random.seed(0)
        out_disk = cli.perform_clustering(td / "no.fasta", csv, td / "disk", 5, 6, "convex", "quadprog", False)
        mem_txt, disk_txt = open(out_mem, "rb").read(), open(out_disk, "rb").read()
        assert mem_txt == disk_txt
        dm = np.load(td / "disk" / "distance_matrix.npy")
        out.update(pc_features_csv=np.frombuffer(open(csv, "rb").read(), dtype=np.uint8),
                   pc_assignment_csv=np.frombuffer(mem_txt, dtype=np.uint8),
                   pc_npy_header=np.frombuffer(open(td / "disk" / "distance_matrix.npy", "rb").read(128), dtype=np.uint8),
                   pc_dm_rows=dm[[0, 57, 179]].copy(), pc_params=np.array([5, 6], dtype=np.int64))
        print("perform_clustering golden:", len(mem_txt), "bytes of binning-assignment.csv,", dm.shape)

    # ---- 6. coverage normalisation (coverage.py:13-43, imported unmodified) and the feature merge (cli/features.py:106-109)
    sys.path.insert(0, REF)
    from ch_bin.core.features.coverage import parse_coverages  # noqa: E402

    def run_parse(path, sep="\t"):
        raw = pd.read_csv(path, sep=sep, header=None).drop(columns=[0]).to_numpy(dtype=np.float64)
        norm = parse_coverages(Path(path), sep).drop(columns=["CONTIG_NAME"]).to_numpy(dtype=np.float64)
        return np.ascontiguousarray(raw), np.ascontiguousarray(norm)

    raw, norm = run_parse(os.path.join(REF, "test_data", "five-genomes-abundance.abund"))
    out.update(cov_raw_real=raw, cov_norm_real=norm)
    crng = np.random.default_rng(11)
    shapes = [(7, 2), (129, 3), (300, 10), (1100, 20)]
    with tempfile.TemporaryDirectory() as td:
        for P, S in shapes:
            vals = np.round(crng.lognormal(3.0, 1.0, (P, S)), 4)
            fn = os.path.join(td, f"cov_{P}_{S}.tsv")
            with open(fn, "w") as f:
                for i in range(P):
                    f.write(f"contig_{i}\t" + "\t".join(repr(float(v)) for v in vals[i]) + "\n")
            raw, norm = run_parse(fn)
            out[f"cov_raw_{P}_{S}"] = raw
            out[f"cov_norm_{P}_{S}"] = norm
        # the merge: the three pandas lines of cli/features.py:106-109 re-run on small frames, then the column drop of
        # cli/clustering.py:53 -- pins the column order [k-mer | coverage] and the PARENT_NAME join
        P, S, n, dk = 300, 10, 420, 24
        fn = os.path.join(td, f"cov_{P}_{S}.tsv")
        df_cov = parse_coverages(Path(fn))
        par_idx = np.concatenate([np.arange(P), crng.integers(0, P, n - P)])
        crng.shuffle(par_idx)
        kmer = crng.dirichlet(np.full(dk, 8.0), n)
        df_init = pd.DataFrame({"CONTIG_NAME": [f"sub_{j}" for j in range(n)], "PARENT_NAME": [f"contig_{q}" for q in par_idx],
                                "CLUSTER": -1})
        df_kmer = pd.concat([pd.DataFrame({"CONTIG_NAME": [f"sub_{j}" for j in range(n)]}),
                             pd.DataFrame(kmer, columns=[f"K{j}" for j in range(dk)])], axis=1)
        df_merged = pd.merge(df_init, df_kmer)
        df_merged = pd.merge(df_merged, df_cov, left_on="PARENT_NAME", right_on="CONTIG_NAME")
        df_merged = df_merged.rename(columns={"CONTIG_NAME_x": "CONTIG_NAME"})
        df_merged = df_merged.drop("CONTIG_NAME_y", axis=1)
        order = np.array([int(s[4:]) for s in df_merged["CONTIG_NAME"]])  # the merge may reorder rows
        samples = df_merged.drop(["CONTIG_NAME", "PARENT_NAME", "CLUSTER"], axis=1).values
        out.update(merge_kmer=kmer, merge_parent=par_idx.astype(np.int64), merge_order=order.astype(np.int64),
                   merge_samples=np.ascontiguousarray(samples), merge_cov_key=np.array([P, S], dtype=np.int64))
        print("coverage goldens:", [(k, out[k].shape) for k in out if k.startswith("cov_norm")], "merge:", samples.shape,
              "row order kept:", bool((order == np.arange(n)).all()))

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
