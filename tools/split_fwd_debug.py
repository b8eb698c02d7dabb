"""Split forward (LSTM2 consumer + LSTM1 / head producer) against the single fused kernel: every output, and the timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, H1, H2 = int(os.environ.get("BB", "32")), int(os.environ.get("TT", "300")), 256, 512
f32 = dict(device=dev, dtype=torch.float32)
g = torch.Generator().manual_seed(3)
r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)
boxes = torch.rand(B, T, 15, 6, generator=g).to(dev)
xproj1 = r(B, T, 4 * H1) * 0.5
w_hh1, w_pred, w_ih2, w_hh2 = r(4 * H1, H1) / H1 ** 0.5, r(15, H1) / H1 ** 0.5, r(4 * H2, 6) / H2 ** 0.5, r(4 * H2, H2) / H2 ** 0.5
names = ["hs1", "gates1", "cells1", "logits", "probs", "fb", "hs2", "gates2", "cells2"]
shapes = [(B, T, H1), (B, T, 4 * H1), (B, T, H1), (B, 15, T), (B, T, 15), (B, T, 6), (B, T, H2), (B, T, 4 * H2), (B, T, H2)]
s = torch.cuda.current_stream().cuda_stream
def run(split):
    os.environ["OPN_OPNET_SPLIT"] = str(int(split))
    outs = [torch.full(sh, float("nan"), **f32) for sh in shapes]
    ws = torch.zeros(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
    def call():
        rc = lib.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xproj1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(),
                               w_hh2.data_ptr(), *[o.data_ptr() for o in outs], ws.data_ptr(), ws.numel(), s)
        assert rc == 0, lib.opn_last_error()
    call(); torch.cuda.synchronize()
    st = ws[:16].view(torch.int32).cpu().tolist()
    if split == 2:
        ph = ws[:4096].view(torch.int64).cpu()[32:40].tolist()
        if sum(ph):
            print("producer CTA 0, clocks per frame: " + "  ".join(f"{n} {v / T:6.0f}" for n, v in zip(
                ["top", "poll+split", "barrier", "MMAs", "barrier", "cells+publish+stash"], ph)) + f"   total {sum(ph) / T:6.0f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): call()
    e1.record(); torch.cuda.synchronize()
    return outs, e0.elapsed_time(e1) / 5, st
ref, t_ref, _ = run(False)
got, t_got, st = run(True)
_, t_prod, _ = run(2)
print(f"single fused kernel {t_ref:.4f} ms; split {t_got:.4f} ms; producer alone {t_prod:.4f} ms; status words after the first split call {st}")
for n, a, b in zip(names, got, ref):
    d = (a - b).abs()
    bad = torch.isnan(d).sum().item()
    print(f"  {n:7s} max|split - fused| = {torch.nan_to_num(d, nan=0.0).max().item():.3e}   NaNs {bad}")
    if n in ("fb", "hs2") and torch.nan_to_num(d, nan=1.0).max().item() > 1e-4:
        dd = torch.nan_to_num(d, nan=1.0)
        idx = (dd.reshape(B, T, -1).amax(-1) > 1e-4).nonzero()
        print("    first bad (video, frame):", idx[:6].tolist(), " count", len(idx))
