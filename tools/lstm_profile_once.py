"""Launch each persistent LSTM kernel exactly once (for ncu): fwd/bwd at H=256 then H=512, B=32, T=300."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T = 32, int(os.environ.get("T", "300"))
for H in (256, 512):
    xp = torch.randn(B, T, 4 * H, device=dev) * 0.5
    whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / (H ** 0.5)
    hs = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, 4 * H, device=dev); cells = torch.empty(B, T, H, device=dev)
    dh = torch.randn(B, T, H, device=dev) * 0.01; dg = torch.empty(B, T, 4 * H, device=dev)
    ws = torch.empty(lib.opn_lstm_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.opn_lstm_fwd(B, T, H, xp.data_ptr(), whh.data_ptr(), hs.data_ptr(), gates.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), s))
    _lib.check(lib.opn_lstm_bwd(B, T, H, whh.data_ptr(), gates.data_ptr(), cells.data_ptr(), dh.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel(), s))
    torch.cuda.synchronize()
print("done")
# fused OPNet forward, once
H1, H2 = 256, 512
f32 = dict(device=dev, dtype=torch.float32)
bx = torch.rand(B, T, 15, 6, **f32); xp1 = torch.randn(B, T, 4 * H1, **f32) * 0.5
w = [(torch.rand(*shape, **f32) * 2 - 1) / (h ** 0.5) for shape, h in (((4 * H1, H1), H1), ((15, H1), H1), ((4 * H2, 6), H2), ((4 * H2, H2), H2))]
outs = [torch.empty(*shape, **f32) for shape in ((B, T, H1), (B, T, 4 * H1), (B, T, H1), (B, 15, T), (B, T, 15), (B, T, 6), (B, T, H2), (B, T, 4 * H2), (B, T, H2))]
wsf = torch.empty(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
_lib.check(lib.opn_opnet_fwd(B, T, H1, H2, bx.data_ptr(), xp1.data_ptr(), *[x.data_ptr() for x in w], *[o.data_ptr() for o in outs], wsf.data_ptr(), wsf.numel(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done fused")
# fused OPNet backward, once
probs = torch.softmax(torch.randn(B, T, 15, **f32), -1)
g1 = torch.rand(B, T, 4 * H1, **f32); c1 = torch.randn(B, T, H1, **f32) * 0.5
g2 = torch.rand(B, T, 4 * H2, **f32); c2 = torch.randn(B, T, H2, **f32) * 0.5
dh2 = torch.randn(B, T, H2, **f32) * 0.01
dg1 = torch.empty(B, T, 4 * H1, **f32); dg2 = torch.empty(B, T, 4 * H2, **f32); dlg = torch.empty(B, T, 15, **f32)
wsb = torch.empty(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
_lib.check(lib.opn_opnet_bwd(B, T, H1, H2, bx.data_ptr(), probs.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(), w[3].data_ptr(),
                             g1.data_ptr(), c1.data_ptr(), g2.data_ptr(), c2.data_ptr(), dh2.data_ptr(), dg1.data_ptr(), dg2.data_ptr(), dlg.data_ptr(),
                             wsb.data_ptr(), wsb.numel(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done fused bwd")
