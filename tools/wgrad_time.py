"""The tcgen05 weight-gradient kernel against the general path (opn_sgemm: split pre-passes + gemm_tc) on the five weight
gradients of the headline OPNet step [32,300] (H1 = 256, H2 = 512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
B, T, H1, H2 = int(os.environ.get("WB", "32")), 300, 256, 512
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)
dg2, fb, hs2, dg1, x1, hs1, dl = r(B, T, 4 * H2), r(B, T, 6), r(B, T, H2), r(B, T, 4 * H1), r(B, T, 90), r(B, T, H1), r(B, T, 15)
w_ih2, w_hh2, w_ih1, w_hh1 = r(4 * H2, 6), r(4 * H2, H2), r(4 * H1, 90), r(4 * H1, H1)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def run_tc():
    j2, a, b = ops._lstm_wgrad_jobs(dg2, fb, hs2, w_ih2, w_hh2, True, True)
    j1, c, d = ops._lstm_wgrad_jobs(dg1, x1, hs1, w_ih1, w_hh1, True, True)
    e = torch.empty(H1, 15, device=dev)
    ops.wgrad_jobs_run(j2 + j1 + [(hs1.reshape(B * T, H1), dl.reshape(B * T, 15), e, T, 0)])
    return a, b, c, d, e
def run_old():
    os.environ["OPN_WGRAD"] = "sgemm"
    a, b = ops._lstm_weight_grads(dg2, fb, hs2, w_ih2, w_hh2, True, True)
    c, d = ops._lstm_weight_grads(dg1, x1, hs1, w_ih1, w_hh1, True, True)
    e = ops._wtt_weight_grad(hs1, dl).t()
    os.environ["OPN_WGRAD"] = "tc"
    return a, b, c, d, e
new, old = run_tc(), run_old()
torch.cuda.synchronize()
for n, o, name in zip(new, old, ("dW_ih2", "dW_hh2", "dW_ih1", "dW_hh1", "dW_pred^T")):
    print(f"{name:10s} max|new - old| {(n - o).abs().max().item():.3e}  of max {o.abs().max().item():.3e}")
for name, fn in (("tcgen05 weight-gradient kernel", run_tc), ("general path (opn_sgemm)", run_old)):
    for mode in ("fp32", "16bit"):
        ops.set_precision(mode)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            for _ in range(12):      # ~0.6 ms of queued GPU work: the host enqueues fn() while it runs (no launch gaps timed)
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ops.set_precision("fp32")
        print(f"{name:32s} {mode:6s} B={B}: {sorted(ts)[len(ts) // 2]:8.1f} us (five weight gradients, cold L2)")
