#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 clustering hot path (BASELINE.json: "point-to-hull QP distances/sec
(k=5, d~137) and clustering-stage wall time").

A STEP is one complete clustering stage on one synthetic contig set: distance structure build + every assignment
iteration until the reference's stop rule fires (cli/clustering.py:56-76 of the reference).  Work is counted in
REFERENCE-EQUIVALENT QPs: sum over executed iterations of U*C (one (query, bin) pair = kNN gather + QP + residual
norm).  Speculative re-solves the GPU path performs on top of that are NOT counted.

  value : QPs/s with the feature matrix already resident in HBM when the timed region starts.
  e2e   : the same metric through the public API chbin_b200.fit_cluster() with HOST (pinned) buffers -- context
          creation, H2D of features/labels, D2H of the final labels all inside the timed region.
  roofline     : the dominant kernel of the timed region, timed live with CUDA events on the launch stream.
  cpu_baseline : the oracle port (oracle/*.c) on the host cores, bounded sample, rank 0 at N=1 only.

`--impl reference` times the reference's CPU algorithm (the oracle port: /root/reference is pure Python + two
absent third-party solvers and cannot travel to the GPU box) on all host threads, on a bounded sample.

N > 1 (torchrun): queries are sharded over ranks, the per-round label exchange is an NCCL all-reduce; the workload is
the same total problem (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_ITERATIONS = 10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="20k")
    ap.add_argument("--n", type=int, default=None, help="override the number of contigs of the workload")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(args):
    from chbin_b200 import synth

    X, bins, truth, cfg = synth.make_config(args.workload, seed=0, n=args.n)
    U = int(np.sum(bins == -1))
    return X, bins, cfg, U


def flops_per_qp(k, d):
    return k * (k + 1) * d + 4 * k * d + 3 * d + 2 * k ** 3  # SURVEY.md 8(d) F(k,d)


def bytes_per_qp(k, d, C):
    return 8 * k * d + 4 * k + 8 + 8 * d / C  # SURVEY.md 8(d) B(k,d)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  The timed region of this
    benchmark is short (tens of milliseconds), so the primary sampler polls NVML in-process every few milliseconds;
    `nvidia-smi --query-gpu=... -lms` is the fallback when the NVML bindings are missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_s=0.004):
        self.rows, self.proc, self.gpu, self.period = [], None, gpu_index, period_s
        self.nv, self.handle, self.thread, self.stop_flag = None, None, None, threading.Event()
        self.sm, self.mx, self.reasons = [], [], set()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            try:
                self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            except Exception:
                pass
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        bits = {}
        for nm, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                         ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            v = getattr(nv, attr, None)
            if v is None:
                v = getattr(nv, attr.replace("ClocksEventReason", "ClocksThrottleReason"), None)
            if v is not None:
                bits[nm] = int(v)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    r = int(get_reasons(self.handle))
                    for nm, b in bits.items():
                        if r & b:
                            self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------------
def cpu_sample(X, bins, cfg, seconds, threads):
    """Oracle port on `threads` host threads over a bounded prefix of iteration 1 (rows computed on the fly)."""
    import oracle

    C, k = cfg["C"], cfg["k"]
    perms = oracle.draw_permutations(bins, 1, seed=0)
    U = perms.shape[1]
    probe = min(U, 256)
    t0 = time.perf_counter()
    oracle.fit_cluster(X, C, bins, None, k, 1, perms=perms, threads=threads, max_steps=probe)
    per_step = (time.perf_counter() - t0) / probe
    steps = int(max(probe, min(U, seconds / max(per_step, 1e-9))))
    t0 = time.perf_counter()
    _, info = oracle.fit_cluster(X, C, bins, None, k, 1, perms=perms, threads=threads, max_steps=steps, return_info=True)
    dt = time.perf_counter() - t0
    return info["qps"] / dt, steps, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, bins, cfg, U = workload(args)
    threads = os.cpu_count() or 1
    per_step_seconds = max(2.0, min(args.cpu_seconds, 150.0 / max(args.steps + args.warmup, 1)))
    vals, steps_used = [], 0
    for i in range(args.warmup + args.steps):
        v, steps_used, dt = cpu_sample(X, bins, cfg, per_step_seconds, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    qps = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = (f"first {steps_used} sequential steps of iteration 1 ({steps_used * cfg['C']} QPs) of the {args.workload} workload, "
              f"distance rows recomputed on the fly, OpenMP over rows/bins")
    line = {
        "impl": "reference", "metric": "point-to-hull QP distances/sec", "value": qps, "unit": "QP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, cfg, U),
        "cpu_baseline": {"value": qps, "unit": "QP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU oracle port (GI restatement); quadprog/cvxopt and the Python reference are not installable offline",
    }
    _emit(line)


def config_dict(args, cfg, U):
    return {"workload": f"synthetic {args.workload}: n={cfg['n']} contigs, d={136 + cfg['S']} (136-dim 4-mer + {cfg['S']} coverage), "
                        f"C={cfg['C']} bins, n_seed={cfg['n_seed']}, k={cfg['k']}, U={U}, max_iterations={MAX_ITERATIONS}",
            "n": cfg["n"], "d": 136 + cfg["S"], "C": cfg["C"], "k": cfg["k"], "U": U,
            "in_mem_dist_matrix": True, "l2_policy": "L2 flushed between steps (256 MB memset outside the timed brackets)",
            "window": args.window}


# ----------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    import chbin_b200
    from chbin_b200 import capi, clustering

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        import torch.distributed as dist

        import datetime

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    X, bins, cfg, U = workload(args)
    n, d = X.shape
    C, k = cfg["C"], cfg["k"]
    perms = clustering.draw_permutations(bins, MAX_ITERATIONS, seed=0)  # the host RNG contract (algorithm.py:45)

    # ---------------- value arm: features resident in HBM, one context reused ----------------
    ctx = capi.Context(local_rank)
    stream = torch.cuda.Stream(dev)  # the library's kernels, torch events and NCCL collectives all run on this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    Xd = torch.from_numpy(X).to(dev)
    ctx.set_features_dev(Xd.data_ptr(), n, d)
    ctx.set_params(k, "convex")
    ctx.set_window(args.window)
    u0, u1 = clustering.owned_slots(U, rank, world)
    engine = comm = None

    def one_stage():
        nonlocal engine, comm
        ctx.set_labels(bins, C, u0, u1)
        ctx.build_distance_matrix(True)
        if world == 1:
            labels, iters, conv, changed = ctx.fit(perms, MAX_ITERATIONS)
            return labels, iters
        if engine is None:
            engine = clustering.GpuEngine(ctx, local_rank, stream)
            comm = clustering.TorchComm()
        iters = 0
        with engine.stream_context():
            for it in range(MAX_ITERATIONS):
                nch, _ = clustering.run_iteration(engine, perms[it], comm)
                iters += 1
                if nch == 0:
                    break
        return ctx.get_labels(), iters

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        labels_w, iters_w = one_stage()
    barrier()
    ctx.reset_timers()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # L2 policy: a 256 MB memset (2x the 126 MB L2) runs between steps, outside the per-step event brackets, so that no
    # step starts with the previous step's feature rows or candidate lists in cache
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    events = []
    barrier()
    total_iters = 0
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        labels_v, iters = one_stage()
        e1.record(stream)
        events.append((e0, e1))
        total_iters += iters
    barrier()
    elapsed_ms = float(sum(a.elapsed_time(b) for a, b in events))
    clocks = sampler.stop() if rank == 0 else None
    tm = ctx.timers()
    t_el = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_el.item())
    qps_ref_total = total_iters * U * C
    value = qps_ref_total / (elapsed_ms * 1e-3)
    fp64_peak = ctx.measure_fp64_tflops() if rank == 0 else 0.0
    ctx.close()  # the end-to-end arm below builds its own context through the public API: release this one's HBM first

    # ---------------- e2e arm: public API, host (pinned) buffers ----------------
    Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
    Xp.copy_(torch.from_numpy(X))
    Xh = Xp.numpy()
    bp = torch.empty((n,), dtype=torch.int64, pin_memory=True)
    bp.copy_(torch.from_numpy(bins))
    bh = bp.numpy()
    e2e_iters = 0
    for i in range(max(1, min(args.warmup, 2))):
        np.random.seed(0)
        chbin_b200.fit_cluster(Xh, C, bh, None, k, MAX_ITERATIONS, device=local_rank, window=args.window)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        np.random.seed(0)
        t0 = time.perf_counter()
        lab_e, info_e = chbin_b200.fit_cluster(Xh, C, bh, None, k, MAX_ITERATIONS, device=local_rank, window=args.window,
                                               return_info=True)
        e2e_s += time.perf_counter() - t0
        e2e_iters += info_e["iterations"]
    barrier()
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    e2e_value = e2e_iters * U * C / e2e_s
    same = bool(np.array_equal(lab_e, labels_v))

    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        ms = {"distance": tm["ms_distance"], "gram": tm["ms_gram"], "knn": tm["ms_knn"], "qp": tm["ms_qp"],
              "commit": tm["ms_commit"]}
        ln = {"distance": tm["launches_distance"], "gram": tm["launches_gram"], "knn": tm["launches_knn"],
              "qp": tm["launches_qp"], "commit": tm["launches_commit"]}
        dom = None  # chosen below among the stages that ran and have a roofline
        traffic_all = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic_all = json.load(open(tp))
            except Exception:
                traffic_all = {}
        nown = u1 - u0
        try:
            tf32_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2.0
            tf32_src = "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)"
        except Exception:
            tf32_peak, tf32_src = 1590.0 / 2.0, "fallback (B200_PROFILING.md bf16 figure / 2)"
        stages = {}
        if ln["gram"]:
            # fused tcgen05 Gram + per-bin selection (distance mode 2): TF32 3-term split, K = 3 * dp8 per 128x128 tile
            dp8 = (d + 7) // 8 * 8
            fl = tm["gram_tiles"] * 2.0 * 128 * 128 * 3 * dp8
            stages["gram"] = {"kernel": "gram_select_kernel", "bound": "tensor", "achieved": fl / (ms["gram"] * 1e-3) / 1e12,
                              "peak": tf32_peak, "unit": "TFLOP/s", "ms_total": ms["gram"], "launches": ln["gram"],
                              "flops_per_launch": fl / ln["gram"], "tiles_128x128": tm["gram_tiles"], "peak_source": tf32_src}
        if ln["knn"]:
            if ln["gram"]:
                # re-rank of the surviving (query, bin) pairs: issue/latency bound set-up work, no bandwidth roofline
                stages["knn"] = {"kernel": "rerank_kernel", "bound": "issue", "achieved": None, "peak": None, "unit": None,
                                 "ms_total": ms["knn"], "launches": ln["knn"]}
            else:
                b = tm["rows_scanned"] * n * 4.0  # kNN scan: one row of FP32 candidate values (4n bytes) per item
                stages["knn"] = {"kernel": "knn_scan_kernel", "bound": "hbm", "achieved": b / (ms["knn"] * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "ms_total": ms["knn"], "launches": ln["knn"], "bytes_per_launch": b / ln["knn"],
                                 "peak_source": hbm_src}
        if ln["qp"]:
            b = tm["qps_solved"] * bytes_per_qp(k, d, C)
            f = tm["qps_solved"] * flops_per_qp(k, d)
            stages["qp"] = {"kernel": "qp_small_kernel" if k <= 5 else ("qp_mid_kernel" if k <= 10 else "qp_kernel"),
                            "bound": "hbm", "achieved": b / (ms["qp"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "ms_total": ms["qp"], "launches": ln["qp"], "bytes_per_launch": b / ln["qp"],
                            "qps_solved": tm["qps_solved"], "qps_per_s": tm["qps_solved"] / (ms["qp"] * 1e-3),
                            "fp64_tflops": f / (ms["qp"] * 1e-3) / 1e12, "fp64_peak_tflops": fp64_peak,
                            "fp64_frac": (f / (ms["qp"] * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                            "peak_source": hbm_src,
                            "note": "algorithmic bytes B(k,d) per QP (SURVEY 8d); the gathered rows are L2-resident at this n"}
        if ln["distance"]:
            # candidate-distance Gram of distance mode 1 / the exact-path fallback rows (gram_tc.cu)
            stages["distance"] = {"kernel": "gram_tc_kernel", "bound": "tensor", "achieved": None, "peak": tf32_peak,
                                  "unit": "TFLOP/s", "ms_total": ms["distance"], "launches": ln["distance"],
                                  "peak_source": tf32_src}
        for s in stages.values():
            s["frac"] = (s["achieved"] / s["peak"]) if (s["peak"] and s["achieved"] is not None) else None
            s["share_of_step"] = s["ms_total"] / elapsed_ms
        # dominant kernel = the longest-running stage with a defined roofline (the re-rank and the set-up kernels are
        # issue / latency bound bookkeeping; their shares are reported in `stages` and `launches_by_stage`)
        dom = max((s for s in stages if stages[s].get("achieved") is not None), key=lambda s: ms[s])
        roof = dict(stages[dom])
        roof["traffic"] = traffic_all.get(roof["kernel"])
        roof["avg_launch_ms"] = roof["ms_total"] / roof["launches"]

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, steps_used, dt = cpu_sample(X, bins, cfg, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": "QP/s", "cores": threads, "kind": "port",
                   "sample": f"oracle port, first {steps_used} sequential steps of iteration 1 ({steps_used * C} QPs, {dt:.1f} s), "
                             f"rows recomputed on the fly"}
        line = {
            "metric": "point-to-hull QP distances/sec", "value": value, "unit": "QP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, cfg, U),
            "clustering_stage_ms": elapsed_ms / args.steps, "iterations_per_step": total_iters / args.steps,
            "qps_reference_per_step": qps_ref_total / args.steps,
            "qps_solved_per_step": tm["qps_solved"] / args.steps, "rounds_per_step": tm["rounds"] / args.steps,
            "roofline": roof, "stages": stages, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": int(n * d * 8 + n * 8 + e2e_iters / args.steps * U * 8),
                    "d2h_bytes_per_step": int(n * 8), "ms_per_step": e2e_s * 1e3 / args.steps,
                    "labels_equal_value_arm": same},
            "gpu_launches": int(sum(ln.values()) + tm["launches_other"]),
            "launches_by_stage": dict(ln, other=tm["launches_other"]),
            "clocks": clocks,
        }
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 meanwhile (NCCL's version banner,
    library chatter) was diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
