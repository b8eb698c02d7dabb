"""us per step of the batch-wide tcgen05 recurrences against the mma.sync kernels, H = 256 / 512, B = 128 / 256 / 512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib, ops
lib = _lib.load(); dev = torch.device("cuda:0"); T = 300
page = ops.status_page(dev)
s = torch.cuda.current_stream().cuda_stream
for H in (512, 256):
    for B in [int(b) for b in os.environ.get("TC_BATCHES", "32,128,256,512").split(",")]:
        xp = torch.randn(B, T, 4 * H, device=dev) * 0.5
        whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / (H ** 0.5)
        hs = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, 4 * H, device=dev); cells = torch.empty(B, T, H, device=dev)
        dh = torch.randn(B, T, H, device=dev) * 0.01; dg = torch.empty(B, T, 4 * H, device=dev)
        ws = torch.zeros(lib.opn_lstm_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
        for mode in ("0", "1"):
            os.environ["OPN_LSTM_TC"] = mode
            for prec in (0, 1):
                lib.opn_set_precision(prec)
                res = []
                for name in ("fwd", "bwd"):
                    def run():
                        if name == "fwd":
                            _lib.check(lib.opn_lstm_fwd(B, T, H, xp.data_ptr(), whh.data_ptr(), hs.data_ptr(), gates.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), s))
                        else:
                            _lib.check(lib.opn_lstm_bwd(B, T, H, whh.data_ptr(), gates.data_ptr(), cells.data_ptr(), dh.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel(), s))
                    run(); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3): run()
                    e1.record(); torch.cuda.synchronize()
                    res.append(e0.elapsed_time(e1) / 3)
                st = int(page[0].item())
                print(f"H={H} B={B:4d} {'tcgen05' if mode == '1' else 'mma.sync'} {'16-bit ' if prec else 'split  '}: fwd {res[0]:8.3f} ms ({res[0] * 1e3 / T:6.2f} us/step)  bwd {res[1]:8.3f} ms ({res[1] * 1e3 / T:6.2f} us/step)  status={st}", flush=True)
                page.zero_()
                if mode == "0":
                    break       # the mma.sync kernels have no 16-bit mode yet
        lib.opn_set_precision(0)
