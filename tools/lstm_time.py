"""Time the persistent LSTM kernels alone (CUDA events) for a few shapes; prints us/step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
from ctypes import c_uint32
lib = _lib.load(); dev = torch.device("cuda:0")
T = 300
import itertools
HS = [int(x) for x in os.environ.get("HS", "256,512").split(",")]
print("math =", os.environ.get("OPN_LSTM_MATH", "(default)"), " exchange =", os.environ.get("OPN_LSTM_EXCHANGE", "(default)"))
for B, H in itertools.product([int(x) for x in os.environ.get("BS", "32").split(",")], HS):
    xp = torch.randn(B, T, 4 * H, device=dev) * 0.5
    whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / (H ** 0.5)
    hs = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, 4 * H, device=dev); cells = torch.empty(B, T, H, device=dev)
    dh = torch.randn(B, T, H, device=dev) * 0.01; dg = torch.empty(B, T, 4 * H, device=dev)
    ws = torch.empty(lib.opn_lstm_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    fwd = lambda: _lib.check(lib.opn_lstm_fwd(B, T, H, xp.data_ptr(), whh.data_ptr(), hs.data_ptr(), gates.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), s))
    bwd = lambda: _lib.check(lib.opn_lstm_bwd(B, T, H, whh.data_ptr(), gates.data_ptr(), cells.data_ptr(), dh.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel(), s))
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        fwd(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        info = (c_uint32 * 3)(); rc = lib.opn_lstm_status(ws.data_ptr(), info)
        print(f"B={B} H={H} {name}: {e0.elapsed_time(e1) / 3 * 1e3 / T:8.3f} us/step  status={rc} {list(info) if rc else ''}", flush=True)

# fused OPNet forward (LSTM1 + who-to-track + LSTM2 in one persistent kernel)
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
ws = torch.empty(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
fused = lambda: _lib.check(lib.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xp1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(), w_hh2.data_ptr(),
                                             hs1.data_ptr(), g1.data_ptr(), c1.data_ptr(), lg.data_ptr(), pr.data_ptr(), fb.data_ptr(), hs2.data_ptr(), g2.data_ptr(), c2.data_ptr(),
                                             ws.data_ptr(), ws.numel(), s))
fused(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): fused()
e1.record(); torch.cuda.synchronize()
info = (c_uint32 * 3)(); rc = lib.opn_lstm_status(ws.data_ptr(), info)
print(f"B={B} fused OPNet forward: {e0.elapsed_time(e1) / 3 * 1e3 / T:8.3f} us/frame  ({e0.elapsed_time(e1) / 3:.3f} ms)  status={rc} {list(info) if rc else ''}", flush=True)
