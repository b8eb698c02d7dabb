"""Per-CTA spans of the five-job weight-gradient launch of the headline step (phases build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, H1, H2 = int(os.environ.get("WB", "32")), 300, 256, 512
rows = B * T
r = lambda *s: (torch.rand(*s, device=dev) * 2 - 1)
dg2, fb, hs2, dg1, x1, hs1, dl = r(rows, 4 * H2), r(rows, 6), r(rows, H2), r(rows, 4 * H1), r(rows, 90), r(rows, H1), r(rows, 15)
outs = [torch.empty(4 * H2, H2, device=dev), torch.empty(4 * H2, 6, device=dev), torch.empty(4 * H1, H1, device=dev),
        torch.empty(4 * H1, 90, device=dev), torch.empty(H1, 15, device=dev)]
spec = [(dg2, hs2, outs[0], 1), (dg2, fb, outs[1], 0), (dg1, hs1, outs[2], 1), (dg1, x1, outs[3], 0), (hs1, dl, outs[4], 0)]
jobs = (_lib.WgradJob * 5)(*[_lib.WgradJob(a.data_ptr(), b.data_ptr(), o.data_ptr(), a.shape[1], b.shape[1], o.shape[1], rows, T, a.shape[1], b.shape[1], sh)
                             for a, b, o, sh in spec])
n = lib.opn_wgrad_workspace_bytes(5, jobs)
ws = torch.zeros(n, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.opn_wgrad(5, jobs, ws.data_ptr(), n, s))
torch.cuda.synchronize()
w = ws[:4096].view(torch.int64).cpu()
spans = w[64:64 + 2 * 224].view(-1, 2)
spans = spans[spans[:, 1] > 0]
t0 = spans[:, 0].min().item()
print(f"{len(spans)} CTAs; last end {(spans[:, 1].max().item() - t0) / 1e3:.1f} us after the first start")
for i in range(0, len(spans), 8):
    print(f"  cta {i:3d}.. :", " ".join(f"{(a - t0) / 1e3:5.1f}-{(b - t0) / 1e3:5.1f}" for a, b in spans[i:i + 8].tolist()))
