"""Per-phase cycle breakdown of the tensor-core recurrence kernels (needs the OPN_LSTM_PHASES build:
python -m objectpermanence_b200.build --phases; run with OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
T, B = 300, 32
FWD = ["publish+stores", "poll", "split+STS+barrier", "MMAs", "barrier", "pointwise", "-", "-"]
BWD = ["reduce", "cell backward", "barrier", "MMAs+publish", "poll", "-", "-", "-"]
for H in (256, 512):
    xp = torch.randn(B, T, 4 * H, device=dev) * 0.5
    whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / (H ** 0.5)
    hs = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, 4 * H, device=dev); cells = torch.empty(B, T, H, device=dev)
    dh = torch.randn(B, T, H, device=dev) * 0.01; dg = torch.empty(B, T, 4 * H, device=dev)
    ws = torch.zeros(lib.opn_lstm_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for name, labels in (("fwd", FWD), ("bwd", BWD)):
        for _ in range(2):
            if name == "fwd":
                _lib.check(lib.opn_lstm_fwd(B, T, H, xp.data_ptr(), whh.data_ptr(), hs.data_ptr(), gates.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), s))
            else:
                _lib.check(lib.opn_lstm_bwd(B, T, H, whh.data_ptr(), gates.data_ptr(), cells.data_ptr(), dh.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel(), s))
        torch.cuda.synchronize()
        words = ws[:4096].view(torch.int64).cpu()
        for cta, off in ((0, 32), (77, 64)):
            ph = words[off:off + 8].tolist()
            tot = sum(ph)
            print(f"H={H} {name} cta {cta}: total {tot / T:7.0f} clk/step | " + "  ".join(f"{l} {v / T:6.0f}" for l, v in zip(labels, ph) if l != "-"), flush=True)

# fused OPNet forward
B, H1, H2 = 32, 256, 512
f32 = dict(device=dev, dtype=torch.float32)
boxes = torch.rand(B, T, 15, 6, **f32)
xp1 = torch.randn(B, T, 4 * H1, **f32) * 0.5
w_hh1 = (torch.rand(4 * H1, H1, **f32) * 2 - 1) / H1 ** 0.5
w_pred = (torch.rand(15, H1, **f32) * 2 - 1) / H1 ** 0.5
w_ih2 = (torch.rand(4 * H2, 6, **f32) * 2 - 1) / H2 ** 0.5
w_hh2 = (torch.rand(4 * H2, H2, **f32) * 2 - 1) / H2 ** 0.5
hs1 = torch.empty(B, T, H1, **f32); g1 = torch.empty(B, T, 4 * H1, **f32); c1 = torch.empty(B, T, H1, **f32)
hs2 = torch.empty(B, T, H2, **f32); g2 = torch.empty(B, T, 4 * H2, **f32); c2 = torch.empty(B, T, H2, **f32)
lg = torch.empty(B, 15, T, **f32); pr = torch.empty(B, T, 15, **f32); fb = torch.empty(B, T, 6, **f32)
ws = torch.zeros(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(lib.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xp1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(), w_hh2.data_ptr(),
                                 hs1.data_ptr(), g1.data_ptr(), c1.data_ptr(), lg.data_ptr(), pr.data_ptr(), fb.data_ptr(), hs2.data_ptr(), g2.data_ptr(), c2.data_ptr(),
                                 ws.data_ptr(), ws.numel(), s))
torch.cuda.synchronize()
words = ws[:4096].view(torch.int64).cpu()
FUSED = ["publish2+stores", "poll h1", "split+barrier", "MMA1+head+barrier", "pointwise1/head", "poll h2", "split+bar+MMA2+bar", "pointwise2"]
for cta, off in ((0, 32), (77, 64)):
    ph = words[off:off + 8].tolist()
    print(f"fused fwd cta {cta}: total {sum(ph) / T:7.0f} clk/frame | " + "  ".join(f"{l} {v / T:6.0f}" for l, v in zip(FUSED, ph)), flush=True)

# fused OPNet backward
probs = torch.softmax(torch.randn(B, T, 15, **f32), -1)
g1.uniform_(0.05, 0.95); g2.uniform_(0.05, 0.95); c1.normal_(0, 0.5); c2.normal_(0, 0.5)
dh2 = torch.randn(B, T, H2, **f32) * 0.01
dg1 = torch.empty(B, T, 4 * H1, **f32); dg2 = torch.empty(B, T, 4 * H2, **f32); dlg = torch.empty(B, T, 15, **f32)
wsb = torch.zeros(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
for _ in range(2):
    _lib.check(lib.opn_opnet_bwd(B, T, H1, H2, boxes.data_ptr(), probs.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(), w_hh2.data_ptr(),
                                 g1.data_ptr(), c1.data_ptr(), g2.data_ptr(), c2.data_ptr(), dh2.data_ptr(), dg1.data_ptr(), dg2.data_ptr(), dlg.data_ptr(),
                                 wsb.data_ptr(), wsb.numel(), s))
torch.cuda.synchronize()
words = wsb[:4096].view(torch.int64).cpu()
FUSEDB = ["gather2+reduce", "cell bwd 2", "barrier A", "MMA2+publish+dfb share", "B0+pub+gather dfb", "sum shares (B1,B2)", "head|gather1,B3,cell bwd 1", "MMA1+publish"]
for cta, off in ((0, 32), (77, 64)):
    ph = words[off:off + 8].tolist()
    print(f"fused bwd cta {cta}: total {sum(ph) / T:7.0f} clk/frame | " + "  ".join(f"{l} {v / T:6.0f}" for l, v in zip(FUSEDB, ph)), flush=True)
