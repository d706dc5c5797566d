"""Wall-clock phases of one sharded clustering stage with resident features (the `scale_workloads` measurement of bench.py),
rank 0's view, every phase synchronised so that its cost is attributed to it (the sum exceeds the pipelined stage time).
usage: torchrun --nproc-per-node N tools/stage_marks_multi.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from chbin_b200 import synth, clustering

wl = sys.argv[1] if len(sys.argv) > 1 else "1m"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
X, bins, truth, cfg = synth.make_config(wl, seed=0)
R = bench.StageRunner(torch, dist, rank, world, lr, stream, X, bins, cfg, 0)
ctx = R.ctx
ctx.enable_timers(False)
T = time.perf_counter
def sync():
    torch.cuda.synchronize(dev)
for rep in range(4):
    R.run()  # warm-up / pipelined reference
    dist.barrier(); sync()
    t0 = T(); R.run(); sync(); piped = (T() - t0) * 1e3
    dist.barrier(); sync()
    marks = []
    def mark(l):
        sync(); marks.append((l, T()))
    mark("start")
    ctx.set_labels(R.bins, R.C, R.u0, R.u1); mark("set_labels")
    ctx.build_distance_matrix(True); mark("build_distance_matrix")
    if R.engine is None:
        R.engine = clustering.GpuEngine(ctx, lr, stream); R.comm = clustering.TorchComm()
    ctx.iteration_prefetch(R.perms[0]); mark("prefetch perm 0")
    clustering.exchange_guess(R.engine, R.comm, R.U); mark("exchange_guess")
    with R.engine.stream_context():
        for it in range(10):
            nch, rounds = clustering.run_iteration(R.engine, R.perms[it], R.comm, next_perm=R.perms[it + 1])
            mark(f"iteration {it + 1} ({rounds} rounds, {nch} changed)")
            if nch == 0:
                break
    lab = ctx.get_labels(); mark("get_labels")
    if rank == 0 and rep >= 2:
        print(f"rep {rep}: pipelined stage {piped:.3f} ms; phases synchronised: {(marks[-1][1] - marks[0][1]) * 1e3:.3f} ms, acc {float(np.mean(lab == truth)):.4f}")
        for (l0, a), (l1, b) in zip(marks, marks[1:]):
            print(f"    +{(b - a) * 1e3:7.3f} ms  {l1}")
dist.barrier(); dist.destroy_process_group()
