"""Per-role cycle breakdown of the tcgen05 weight-gradient kernel, CTA 0 (phases build: OPN_B200_LIB=.../libopnet_b200_phases.so)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, H = 32, 300, int(os.environ.get("WH", "512"))
rows = B * T
a = torch.rand(rows, 4 * H, device=dev) * 2 - 1
b = torch.rand(rows, H, device=dev) * 2 - 1
out = torch.empty(4 * H, H, device=dev)
jobs = (_lib.WgradJob * 1)(_lib.WgradJob(a.data_ptr(), b.data_ptr(), out.data_ptr(), 4 * H, H, H, rows, T, 4 * H, H, 1))
n = lib.opn_wgrad_workspace_bytes(1, jobs)
ws = torch.zeros(n, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(lib.opn_wgrad(1, jobs, ws.data_ptr(), n, s))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _lib.check(lib.opn_wgrad(1, jobs, ws.data_ptr(), n, s)); e1.record(); torch.cuda.synchronize()
print(f"dW_hh H={H} rows={rows}: {e0.elapsed_time(e1) * 1e3:.1f} us, launches {_lib.launch_count()}")
w = ws[:4096].view(torch.int64).cpu()
for role, name, labels in ((0, "A converter (group 0)", ["loads", "wait stage free", "convert + tmem st", "epilogue (wait done + stores)"]),
                           (2, "MMA thread", ["wait stage full", "issue + commit"])):
    ph = w[32 + 4 * role: 32 + 4 * role + len(labels)].tolist()
    print(f"{name:22s}: total {sum(ph):9d} clk | " + "  ".join(f"{l} {v:9d}" for l, v in zip(labels, ph)))
spans = w[64:64 + 2 * 224].view(-1, 2)
spans = spans[spans[:, 1] > 0]
t0 = spans[:, 0].min().item()
d = (spans[:, 1] - spans[:, 0]).double() / 1e3
print(f"per-CTA spans: {len(d)} CTAs, {d.min():.1f} .. {d.max():.1f} us (mean {d.mean():.1f}); last end {(spans[:, 1].max().item() - t0) / 1e3:.1f} us after the first start")
print(f"CTA 0: {w[60].item()} clocks in {(spans[0, 1] - spans[0, 0]).item() / 1e3:.1f} us = {w[60].item() / (spans[0, 1] - spans[0, 0]).item():.2f} GHz")
