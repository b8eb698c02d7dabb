"""Each weight gradient of the headline step alone through the tcgen05 weight-gradient kernel (GPU time by events behind queued work)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
B, T, H1, H2 = 32, 300, 256, 512
r = lambda *s: (torch.rand(*s, device=dev) * 2 - 1)
dg2, fb, hs2, dg1, x1, hs1, dl = r(B * T, 4 * H2), r(B * T, 6), r(B * T, H2), r(B * T, 4 * H1), r(B * T, 90), r(B * T, H1), r(B * T, 15)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
jobs = {"dW_hh2": (dg2, hs2, torch.empty(4 * H2, H2, device=dev), T, 1), "dW_ih2": (dg2, fb, torch.empty(4 * H2, 6, device=dev), T, 0),
        "dW_hh1": (dg1, hs1, torch.empty(4 * H1, H1, device=dev), T, 1), "dW_ih1": (dg1, x1, torch.empty(4 * H1, 90, device=dev), T, 0),
        "dW_pred": (hs1, dl, torch.empty(H1, 15, device=dev), T, 0)}
def timed(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        for _ in range(12):
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
for name, job in jobs.items():
    print(f"{name:8s} alone: {timed(lambda: ops.wgrad_jobs_run([job])):8.1f} us")
print(f"all five: {timed(lambda: ops.wgrad_jobs_run(list(jobs.values()))):8.1f} us")
print(f"hh2+hh1 : {timed(lambda: ops.wgrad_jobs_run([jobs['dW_hh2'], jobs['dW_hh1']])):8.1f} us")
