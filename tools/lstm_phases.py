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
