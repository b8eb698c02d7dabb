"""Split backward (LSTM2 loop + head / LSTM1 kernel on the idle SMs) against the single fused kernel: every output, and the timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, H1, H2 = int(os.environ.get("BB", "32")), int(os.environ.get("TT", "300")), 256, 512
f32 = dict(device=dev, dtype=torch.float32)
g = torch.Generator().manual_seed(5)
R = lambda *s: torch.rand(*s, generator=g).to(dev)
boxes = R(B, T, 15, 6); probs = torch.softmax(torch.randn(B, T, 15, generator=g), -1).to(dev)
w_hh1 = (R(4 * H1, H1) * 2 - 1) / H1 ** 0.5; w_pred = (R(15, H1) * 2 - 1) / H1 ** 0.5
w_ih2 = (R(4 * H2, 6) * 2 - 1) / H2 ** 0.5; w_hh2 = (R(4 * H2, H2) * 2 - 1) / H2 ** 0.5
g1 = R(B, T, 4 * H1); g1[..., 2 * H1:3 * H1] = g1[..., 2 * H1:3 * H1] * 2 - 1; c1 = torch.randn(B, T, H1, generator=g).to(dev) * 0.5
g2 = R(B, T, 4 * H2); g2[..., 2 * H2:3 * H2] = g2[..., 2 * H2:3 * H2] * 2 - 1; c2 = torch.randn(B, T, H2, generator=g).to(dev) * 0.5
dh2 = torch.randn(B, T, H2, generator=g).to(dev) * 0.01
s = torch.cuda.current_stream().cuda_stream
def run(split):
    os.environ["OPN_OPNET_SPLIT"] = str(int(split))
    dg1, dg2, dl = torch.full((B, T, 4 * H1), float("nan"), **f32), torch.full((B, T, 4 * H2), float("nan"), **f32), torch.full((B, T, 15), float("nan"), **f32)
    ws = torch.zeros(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
    def call():
        rc = lib.opn_opnet_bwd(B, T, H1, H2, boxes.data_ptr(), probs.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(), w_hh2.data_ptr(),
                               g1.data_ptr(), c1.data_ptr(), g2.data_ptr(), c2.data_ptr(), dh2.data_ptr(), dg1.data_ptr(), dg2.data_ptr(), dl.data_ptr(),
                               ws.data_ptr(), ws.numel(), s)
        assert rc == 0, lib.opn_last_error()
    call(); torch.cuda.synchronize()
    st = ws[:16].view(torch.int32).cpu().tolist()
    w64 = ws[:4096].view(torch.int64).cpu()
    ph = (w64[32:40].tolist(), w64[96:104].tolist())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): call()
    e1.record(); torch.cuda.synchronize()
    return (dg1, dg2, dl), e0.elapsed_time(e1) / 5, st, ph
ref, t_ref, _, _ = run(0)
got, t_got, st, ph = run(1)
print(f"single fused kernel {t_ref:.4f} ms; split {t_got:.4f} ms; status words after the first split call {st[:4]}")
_, t_l2, _, _ = run(3)
print(f"LSTM2 loop alone (EXT mode) {t_l2:.4f} ms")
if sum(ph[0]) or sum(ph[1]):
    print("LSTM2 loop CTA 0, clocks per frame:", [round(v / T) for v in ph[0]], "total", round(sum(ph[0]) / T))
    print("head/LSTM1 kernel unit CTA 0, clocks per frame [top, polls, cells+frags, barrier, MMAs+publish+barrier]:", [round(v / T) for v in ph[1]], "total", round(sum(ph[1]) / T))
for n, a, b in zip(["dgates1", "dgates2", "dlogits"], got, ref):
    d = (a - b).abs()
    print(f"  {n:8s} max|split - fused| = {torch.nan_to_num(d, nan=0.0).max().item():.3e} of max {b.abs().max().item():.3e}   NaNs {torch.isnan(d).sum().item()}")
