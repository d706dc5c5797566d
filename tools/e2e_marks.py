"""Host wall-clock marks of chbin_b200.fit_cluster (CHB_PROFILE_FIT) on the 20k workload, pinned C-ordered input as in
bench.py's e2e arm: mean over the timed calls.  usage: python tools/e2e_marks.py [reps]   (CHB_NO_PDL=1 for the A/B)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CHB_PROFILE_FIT"] = "1"
import numpy as np
import torch
import chbin_b200
from chbin_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
X, bins, truth, cfg = synth.make_config("20k", seed=0)
Xp = torch.empty(X.shape, dtype=torch.float64, pin_memory=True); Xp.copy_(torch.from_numpy(X)); Xh = Xp.numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
acc, tot = {}, []
for rep in range(reps + 3):
    flush.zero_(); torch.cuda.synchronize()
    np.random.seed(0)
    t0 = time.perf_counter()
    lab, info = chbin_b200.fit_cluster(Xh, cfg["C"], bins, None, cfg["k"], 10, device=0, return_info=True)
    dt = (time.perf_counter() - t0) * 1e3
    if rep >= 3:
        tot.append(dt)
        prev = 0.0
        for label, ms in info["marks_ms"]:
            acc.setdefault(label, []).append(ms - prev)
            prev = ms
print("pdl", "off" if os.environ.get("CHB_NO_PDL") else "on", "total ms: mean %.3f min %.3f max %.3f" % (np.mean(tot), np.min(tot), np.max(tot)),
      "labels ok", bool(np.array_equal(lab, truth)))
for label, v in acc.items():
    print("   %-28s +%.3f ms" % (label, float(np.mean(v))))
