"""Time (and, with the phases build, break down) the fused OPNet backward kernel alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, T, H1, H2 = 32, int(os.environ.get("TT", "300")), 256, 512
f32 = dict(device=dev, dtype=torch.float32)
boxes = torch.rand(B, T, 15, 6, **f32); probs = torch.softmax(torch.randn(B, T, 15, **f32), -1)
w_hh1 = (torch.rand(4 * H1, H1, **f32) * 2 - 1) / H1 ** 0.5; w_pred = (torch.rand(15, H1, **f32) * 2 - 1) / H1 ** 0.5
w_ih2 = (torch.rand(4 * H2, 6, **f32) * 2 - 1) / H2 ** 0.5; w_hh2 = (torch.rand(4 * H2, H2, **f32) * 2 - 1) / H2 ** 0.5
g1 = torch.rand(B, T, 4 * H1, **f32); c1 = torch.randn(B, T, H1, **f32) * 0.5
g2 = torch.rand(B, T, 4 * H2, **f32); c2 = torch.randn(B, T, H2, **f32) * 0.5
dh2 = torch.randn(B, T, H2, **f32) * 0.01
dg1 = torch.empty(B, T, 4 * H1, **f32); dg2 = torch.empty(B, T, 4 * H2, **f32); dl = torch.empty(B, T, 15, **f32)
ws = torch.zeros(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
call = lambda: _lib.check(lib.opn_opnet_bwd(B, T, H1, H2, boxes.data_ptr(), probs.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(), w_hh2.data_ptr(),
                                            g1.data_ptr(), c1.data_ptr(), g2.data_ptr(), c2.data_ptr(), dh2.data_ptr(), dg1.data_ptr(), dg2.data_ptr(), dl.data_ptr(),
                                            ws.data_ptr(), ws.numel(), s))
call(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"fused OPNet backward [B={B},T={T}]: {ms:.3f} ms = {ms * 1e3 / T:.3f} us/frame")
st = ws[:64].view(torch.int32).cpu().tolist()
print(f"stale sweeps of thread 0 / CTA 0 over {T} frames: LSTM2 inbox {st[8]}, LSTM1 inbox {st[9]}, d frames_boxes tile {st[10]}")
words = ws[:4096].view(torch.int64).cpu()
NAMES = ["gather2+reduce", "cell2", "barrier A", "MMA2+publish+dfb share", "B0+publish+gather dfb", "sum shares (B1,B2)", "head | gather1, B3, cell1", "B4+MMA1+publish"]
for cta, off in ((0, 32), (77, 64)):
    ph = words[off:off + 8].tolist()
    if sum(ph) > 0:
        print(f"cta {cta}: total {sum(ph) / T:7.0f} clk/frame | " + "  ".join(f"{n} {v / T:6.0f}" for n, v in zip(NAMES, ph)), flush=True)
