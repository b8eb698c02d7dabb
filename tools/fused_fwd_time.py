"""Time the fused OPNet forward kernel alone (and check it against the separate kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
B, T, H1, H2 = 32, int(os.environ.get("TT", "300")), 256, 512
g = torch.Generator().manual_seed(1)
r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1)
boxes = torch.rand(B, T, 15, 6, generator=g).to(dev)
w_ih1, w_hh1, w_pred = (r(4 * H1, 90) / H1 ** 0.5).to(dev), (r(4 * H1, H1) / H1 ** 0.5).to(dev), (r(15, H1) / H1 ** 0.5).to(dev)
w_ih2, w_hh2 = (r(4 * H2, 6) / H2 ** 0.5).to(dev), (r(4 * H2, H2) / H2 ** 0.5).to(dev)
def run():
    with torch.no_grad():
        return ops.opnet_trunk(boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2)
hs2, logits = run()
os.environ["OPN_OPNET_FUSED"] = "0"
hs2_ref, logits_ref = run()
os.environ["OPN_OPNET_FUSED"] = "1"
print(f"fused vs separate kernels: max|d hs2| {(hs2 - hs2_ref).abs().max().item():.2e}  max|d logits| {(logits - logits_ref).abs().max().item():.2e}")
from objectpermanence_b200 import _lib
timer = ops.LaunchTimer(); ops.set_launch_timer(timer)
for _ in range(3): run()
torch.cuda.synchronize()
timer = ops.LaunchTimer(); ops.set_launch_timer(timer)
for _ in range(10): run()
torch.cuda.synchronize()
ms = timer.mean_ms("opnet_fwd_fused")
print(f"fused OPNet forward [B={B},T={T}]: {ms:.4f} ms = {ms * 1e3 / T:.3f} us/frame")
