"""Per-role cycle breakdown of the batch-wide tcgen05 recurrences (phases build: python -m objectpermanence_b200.build --phases;
run with OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
os.environ["OPN_LSTM_TC"] = "1"
lib = _lib.load(); dev = torch.device("cuda:0"); T = 300
s = torch.cuda.current_stream().cuda_stream
ROLES = {0: ("fwd TMA", ["loop", "counters", "proxy fence", "empty waits + TMA issue"]),
         1: ("fwd MMA", ["loop", "acc_empty", "first full", "other fulls", "issue+commit"]),
         2: ("fwd epilogue", ["stash stores + xproj issue", "acc_full wait", "tmem ld", "pointwise", "publish"]),
         3: ("bwd epilogue", ["stash issue", "counter wait", "partial sums", "cell bwd + A tile", "dgates stores + acc_full wait", "tmem -> ring + publish"])}
for H in (512, 256):
    for B in (128, 256):
        xp = torch.randn(B, T, 4 * H, device=dev) * 0.5
        whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / (H ** 0.5)
        hs = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, 4 * H, device=dev); cells = torch.empty(B, T, H, device=dev)
        dh = torch.randn(B, T, H, device=dev) * 0.01; dg = torch.empty(B, T, 4 * H, device=dev)
        ws = torch.zeros(lib.opn_lstm_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
        for name, roles in (("fwd", (0, 1, 2)), ("bwd", (3,))):
            for _ in range(2):
                if name == "fwd":
                    _lib.check(lib.opn_lstm_fwd(B, T, H, xp.data_ptr(), whh.data_ptr(), hs.data_ptr(), gates.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), s))
                else:
                    _lib.check(lib.opn_lstm_bwd(B, T, H, whh.data_ptr(), gates.data_ptr(), cells.data_ptr(), dh.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel(), s))
            torch.cuda.synchronize()
            words = ws[:4096].view(torch.int64).cpu()
            for r in roles:
                label, names = ROLES[r]
                ph = words[32 + 8 * r: 32 + 8 * r + len(names)].tolist()
                print(f"H={H} B={B} {label:13s}: total {sum(ph) / T:7.0f} clk/step | " + "  ".join(f"{n} {v / T:6.0f}" for n, v in zip(names, ph)), flush=True)
