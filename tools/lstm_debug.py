"""Debug driver: one LSTM layer forward+backward on the GPU, compared with the oracle."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
from oracle import opnet_oracle as oracle

B, T, I, H = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (2, 8, 6, 512))]
ops.set_debug_sync(True)
g = torch.Generator().manual_seed(0)
rnd = lambda *s: (torch.rand(*s, generator=g, dtype=torch.float64) * 2 - 1)
x, w_ih, w_hh, dh = rnd(B, T, I).float(), (rnd(4 * H, I) / math.sqrt(H)).float(), (rnd(4 * H, H) / math.sqrt(H)).float(), rnd(B, T, H).float()
xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
ref = oracle.lstm_layer(xr, wir, whr); ref.backward(dh.double())
dev = torch.device("cuda:0")
xg, wig, whg = [t.to(dev).requires_grad_(True) for t in (x, w_ih, w_hh)]
out = ops.lstm_layer(xg, wig, whg)
print("fwd err", (out.detach().cpu().double() - ref.detach()).abs().max().item())
try:
    out.backward(dh.to(dev))
    for n, a, b in (("dx", xg.grad, xr.grad), ("dw_ih", wig.grad, wir.grad), ("dw_hh", whg.grad, whr.grad)):
        print(n, "err", (a.cpu().double() - b).abs().max().item(), "ref max", b.abs().max().item())
except Exception as e:
    print("BACKWARD FAILED:", e)
