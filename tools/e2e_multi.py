"""Where the end-to-end time of chbin_b200.fit_cluster goes under torchrun (N ranks): wall-clock marks of the host driver
(CHB_PROFILE_FIT=1), rank 0's view.  usage: torchrun --nproc-per-node N tools/e2e_multi.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CHB_PROFILE_FIT"] = "1"
import numpy as np
import torch
import torch.distributed as dist
import chbin_b200
from chbin_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "20k"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
X, bins, truth, cfg = synth.make_config(wl, seed=0)
Xp = torch.empty(X.shape, dtype=torch.float64, pin_memory=True); Xp.copy_(torch.from_numpy(X)); Xh = Xp.numpy()
for rep in range(5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    np.random.seed(0)
    t0 = time.perf_counter()
    lab, info = chbin_b200.fit_cluster(Xh, cfg["C"], bins, None, cfg["k"], 10, device=lr, return_info=True)
    dt = (time.perf_counter() - t0) * 1e3
    if rank == 0 and rep >= 3:
        print(f"rep {rep}: {dt:.3f} ms, acc {float(np.mean(lab == truth)):.4f}")
        prev = 0.0
        for label, t in info["marks_ms"]:
            print(f"    {t:8.3f} ms  (+{t - prev:6.3f})  {label}")
            prev = t
if world > 1:
    dist.barrier(); dist.destroy_process_group()
