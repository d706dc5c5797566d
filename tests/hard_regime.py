"""Stage time where the bin pruning stops working: BASELINE config #2's shape (20k contigs, d = 137, C = 50, k = 5) with the
genomes made to overlap by lowering the generator's Dirichlet concentration (4000 = the headline workload: separable;
60 = the parity tests' "hard" case), and the 100k / k = 10 hard case of tests/test_gpu_parity_at_size.py.  One clustering
stage per measurement = label set-up + distance structure + iterations up to the limit, features resident, CUDA events
on the context's stream, L2 flushed between stages.  The final labels are checked against the position-parallel oracle on
sampled positions of the LAST executed iteration (test infrastructure; not inside the timed region).
usage: python tests/hard_regime.py [max_iterations] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from chbin_b200 import capi, synth
import oracle

max_it = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
CASES = [("20k", 20_000, 50, 1, 5, 3, c) for c in (4000.0, 1000.0, 250.0, 120.0, 60.0)] + [("100k", 100_000, 100, 10, 10, 4, 120.0)]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
threads = os.cpu_count() or 1
for name, n, C, S, k, seed, conc in CASES:
    X, bins, truth = synth.make_contig_features(n, C, S, 50, seed=seed, concentration=conc)
    perms = oracle.draw_permutations(bins, max_it, seed=0)
    U = perms.shape[1]
    ctx = capi.Context(0)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_features(X); ctx.set_params(k, "convex"); ctx.set_distance_mode(2)
    ms = []
    for s in range(steps + 1):
        ctx.enable_timers(s == steps)  # last pass: per-kernel event timers on (not counted in the stage time)
        ctx.reset_timers()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.set_labels(bins, C); ctx.build_distance_matrix(True)
        labels, iters, conv, changed = ctx.fit(perms, max_it)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if 1 <= s < steps:
            ms.append(e0.elapsed_time(e1))
    t = ctx.timers()
    # labels before the last executed iteration, for the oracle's one-step check of that iteration
    ctx.set_labels(bins, C); ctx.build_distance_matrix(True)
    prev = bins
    for it in range(iters):
        lab, nch = ctx.fit_iteration(perms[it])
        if it < iters - 1:
            prev = lab
    ctx.close()
    rng = np.random.default_rng(11)
    pos = np.unique(np.concatenate([np.arange(128), np.arange(U - 128, U), rng.choice(U, 768, replace=False)])).astype(np.int64)
    res = oracle.verify_iteration(X, C, prev, lab, perms[iters - 1], k, positions=pos, threads=threads)
    print("%s conc=%g k=%d: stage %.2f ms (min %.2f) over %d iterations (%s), %d rounds, QPs solved %d = %.2f x the reference count %d, "
          "%.3g reference-equivalent QP/s; with timers: gram %.2f ms, selection %.2f, qp %.2f, commit %.2f; final labels == run-to-run %s; "
          "oracle: %d mismatches on %d positions of iteration %d; labels == generator truth %.4f"
          % (name, conc, k, float(np.mean(ms)), float(np.min(ms)), iters, "converged" if conv else "iteration limit", t["rounds"],
             t["qps_solved"], t["qps_solved"] / max(t["qps_reference"], 1), t["qps_reference"],
             t["qps_reference"] / (float(np.mean(ms)) * 1e-3), t["ms_gram"], t["ms_knn"], t["ms_qp"], t["ms_commit"],
             bool(np.array_equal(lab, labels)), int(res["mismatches"]), len(pos), iters, float(np.mean(labels == truth))), flush=True)
