"""Stage time of a workload with resident features (CUDA events on the context's stream, L2 flushed between steps), for
A/B runs of the launch-chain switches: CHB_NO_PDL=1, CHB_NO_SIDE=1.  usage: python tools/stage_time.py [workload] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from chbin_b200 import capi, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "20k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
X, bins, truth, cfg = synth.make_config(wl, seed=0)
pts = np.where(bins == -1)[0]
np.random.seed(0)
perms = np.stack([np.random.permutation(pts) for _ in range(10)]).astype(np.int64)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0)
ctx.set_stream(stream.cuda_stream)
ctx.set_features(X); ctx.set_params(cfg["k"], "convex"); ctx.set_distance_mode(2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ctx.enable_timers(False)
ms = []
for s in range(steps + 2):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.set_labels(bins, cfg["C"]); ctx.build_distance_matrix(True)
    labels, iters, conv, changed = ctx.fit(perms, 10)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if s >= 2:
        ms.append(e0.elapsed_time(e1))
flags = " ".join(k for k in ("CHB_NO_PDL", "CHB_NO_SIDE") if os.environ.get(k))
print(f"{wl} [{flags or 'default'}] stage ms: mean {np.mean(ms):.3f} min {np.min(ms):.3f} max {np.max(ms):.3f}  iterations {iters} acc {float(np.mean(labels == truth)):.4f}")
ctx.close()
